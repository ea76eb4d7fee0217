"""``StreamedConvFactor`` on the kernels: conv weight factors emitted in output-channel chunks at every use
(``vivit_b200.set_conv_factor_streaming``) against the materialised factors, same inputs, same GPU."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _problem(dtype):
    torch.manual_seed(3)
    model = nn.Sequential(
        nn.Conv2d(3, 6, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(6, 9, 3, stride=2, padding=1), nn.Sigmoid(),
        nn.Flatten(), nn.Linear(9 * 2 * 2, 5),
    ).to(dtype)
    x = torch.rand(8, 3, 8, 8, dtype=dtype)
    return model.to(DEV), x.to(DEV), torch.randint(0, 5, (8,)).to(DEV)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("chunk_bytes", [1, 20000])
def test_streamed_conv_factor(dtype, chunk_bytes):
    import vivit_b200 as vv
    from vivit_b200.factors import StreamedConvFactor

    top = lambda ev: list(range(ev.numel() - 3, ev.numel()))  # noqa: E731
    damping = lambda evals, evecs, gammas, lambdas: torch.ones_like(evals)  # noqa: E731
    tol = 1e-3 if dtype == torch.float32 else 1e-9  # fp32: two solves of Gram matrices that differ in the last bits

    def run(make):
        model, x, y = _problem(dtype)
        groups = [{"params": list(model.parameters()), "criterion": top, "damping": damping}]
        comp = make()
        m, lf = vv.extend(model), vv.extend(nn.CrossEntropyLoss())
        with vv.backpack(*comp.get_extensions(), extension_hook=comp.get_extension_hook(groups)):
            lf(m(x), y).backward()
        return comp.get_result(groups[0])

    def same(a, b, what):
        a, b = a.double(), b.double()
        assert a.shape == b.shape, what
        assert (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1e-30), (what, (a - b).abs().max().item())

    try:
        results = {}
        for chunk in (None, chunk_bytes):
            vv.set_conv_factor_streaming(chunk)
            results[chunk] = (
                run(vv.EighComputation), run(vv.DirectionalDerivativesComputation), run(vv.DirectionalDampedNewtonComputation)
            )
        # the streamed factor is what the extension saves, in several chunks
        vv.set_conv_factor_streaming(chunk_bytes)
        model, x, y = _problem(dtype)
        ext = vv.SqrtGGNExact(lazy=True)
        m, lf = vv.extend(model), vv.extend(nn.CrossEntropyLoss())
        with vv.backpack(ext):
            lf(m(x), y).backward()
        factor = getattr(model[3].weight, ext.savefield)
        assert isinstance(factor, StreamedConvFactor) and (chunk_bytes > 1 or len(list(factor.chunks())) == 9)
    finally:
        vv.set_conv_factor_streaming(None)
    (ev0, vecs0), (g0, l0), steps0 = results[None]
    (ev1, vecs1), (g1, l1), steps1 = results[chunk_bytes]
    same(ev1, ev0, "eigenvalues")
    same(l1, l0, "lambdas")
    same(g1.abs(), g0.abs(), "gammas")
    for a, b in zip(steps1, steps0):
        same(a, b, "Newton step")
    for a, b in zip(vecs1, vecs0):
        for k in range(a.shape[0]):
            sign = torch.sign((a[k].double() * b[k].double()).sum())
            same(a[k] * sign, b[k], "eigenvector")
