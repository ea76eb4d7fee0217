"""``StreamedConvFactor`` (opt-in, ``vivit_b200.set_conv_factor_streaming``): the factor of a convolution weight
emitted in output-channel chunks at every use and never materialised as a whole gives the results of the
materialised factor -- all four Computations and the savefield closures, on the conv fixtures, through the plain-torch
test double (host logic; ``tests/test_parity_gpu.py::test_streamed_conv_factor`` runs it on the kernels)."""
import pytest
import torch
from torch import nn

import tests._torch_kernels as double
import vivit_b200 as vv
from vivit_b200.factors import DenseFactor, StreamedConvFactor


def _problem(one_d=False):
    torch.manual_seed(3)
    if one_d:
        model = nn.Sequential(nn.Conv1d(2, 6, 3, padding=1), nn.ReLU(), nn.Flatten(), nn.Linear(6 * 9, 4)).double()
        x = torch.rand(5, 2, 9, dtype=torch.float64)
    else:
        model = nn.Sequential(
            nn.Conv2d(2, 5, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(5, 7, 3, stride=2, padding=1), nn.Sigmoid(),
            nn.Flatten(), nn.Linear(7 * 2 * 2, 4),
        ).double()
        x = torch.rand(5, 2, 6, 6, dtype=torch.float64)
    return model, x, torch.randint(0, 4, (5,))


def _run(comp, one_d, groups_of):
    model, x, y = _problem(one_d)
    groups = groups_of(model)
    m, lf = vv.extend(model), vv.extend(nn.CrossEntropyLoss())
    with vv.backpack(*comp.get_extensions(), extension_hook=comp.get_extension_hook(groups)):
        lf(m(x), y).backward()
    return [comp.get_result(g) for g in groups]


def _flat(res):
    out = []
    for r in res:
        if torch.is_tensor(r):
            out.append(r)
        else:
            for part in r:
                out += [part] if torch.is_tensor(part) else list(part)
    return out


@pytest.fixture
def streaming(monkeypatch):
    double.install(monkeypatch)
    yield
    vv.set_conv_factor_streaming(None)


@pytest.mark.parametrize("one_d", [False, True])
@pytest.mark.parametrize("chunk_bytes", [1, 2000, 1 << 30])
def test_streamed_conv_factor_gives_the_materialised_results(streaming, one_d, chunk_bytes):
    top = lambda ev: list(range(max(0, ev.numel() - 3), ev.numel()))  # noqa: E731
    damping = lambda evals, evecs, gammas, lambdas: torch.ones_like(evals)  # noqa: E731
    groups_of = lambda model: [{"params": list(model.parameters()), "criterion": top, "damping": damping}]  # noqa: E731
    makers = [
        lambda: vv.EigvalshComputation(),
        lambda: vv.EighComputation(),
        lambda: vv.DirectionalDerivativesComputation(),
        lambda: vv.DirectionalDampedNewtonComputation(),
        lambda: vv.EighComputation(subsampling=[3, 0, 1]),
    ]
    for make in makers:
        vv.set_conv_factor_streaming(None)
        want = _flat(_run(make(), one_d, groups_of))
        vv.set_conv_factor_streaming(chunk_bytes)
        got = _flat(_run(make(), one_d, groups_of))
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert a.shape == b.shape
            if a.dim() >= 2 and a.shape[0] <= 3 and isinstance(make(), vv.EighComputation):  # eigenvectors: sign
                for k in range(a.shape[0]):
                    s = torch.sign((a[k] * b[k]).sum())
                    assert torch.allclose(a[k] * s, b[k], rtol=1e-8, atol=1e-10)
            else:
                assert torch.allclose(a.abs(), b.abs(), rtol=1e-8, atol=1e-10), (a - b).abs().max()


def test_streamed_conv_factor_is_what_the_extension_saves(streaming):
    model, x, y = _problem()
    m, lf = vv.extend(model), vv.extend(nn.CrossEntropyLoss())
    kinds = {}
    for chunk in (None, 1500):
        vv.set_conv_factor_streaming(chunk)
        ext = vv.SqrtGGNExact(lazy=True)
        with vv.backpack(ext):
            lf(m(x), y).backward()
        conv_w = model[0].weight
        factor = getattr(conv_w, ext.savefield)
        kinds[chunk] = factor
        for p in model.parameters():
            p.grad = None
    dense, streamed = kinds[None], kinds[1500]
    assert isinstance(dense, DenseFactor) and isinstance(streamed, StreamedConvFactor)
    chunks = list(streamed.chunks())
    assert len(chunks) > 1 and chunks[0][0] == 0 and chunks[-1][1] == 5  # several chunks over the 5 output channels
    assert all(Vt.numel() * Vt.element_size() <= 1500 or hi - lo == 1 for lo, hi, Vt in chunks)
    assert torch.allclose(streamed.materialize(), dense.materialize(), rtol=1e-12, atol=1e-14)
    assert torch.allclose(streamed.gram_mat(), dense.gram_mat(), rtol=1e-10, atol=1e-12)
    M = torch.rand(3, *conv_w.shape, dtype=torch.float64)
    assert torch.allclose(streamed.vt_mat_prod(M), dense.vt_mat_prod(M), rtol=1e-10, atol=1e-12)
    v = torch.rand(streamed.R, dtype=torch.float64)
    assert torch.allclose(streamed.v_apply(v), dense.v_apply(v), rtol=1e-10, atol=1e-12)
