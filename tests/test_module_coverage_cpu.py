"""Module map edge cases on CPU through the test double."""

import pytest
import torch
from torch import nn

import tests._torch_kernels as double
from tests.problems import ND_NETS
from tests.test_host_cpu import run_backward


@pytest.fixture(autouse=True)
def torch_kernels(monkeypatch):
    double.install(monkeypatch)


def test_batchnorm_in_training_mode_is_rejected():
    """[BackPACK] has no per-sample GGN factor for a BatchNorm layer that mixes the batch."""
    from vivit_b200 import BatchGrad, SqrtGGNExact

    torch.manual_seed(0)
    model = nn.Sequential(nn.BatchNorm1d(3), nn.Flatten(), nn.Linear(12, 3))
    x, y = torch.rand(4, 3, 4), torch.randint(0, 3, (4,))
    for ext in (SqrtGGNExact(), BatchGrad()):
        with pytest.raises(NotImplementedError):
            run_backward(model, nn.CrossEntropyLoss(), x, y, [ext], None)


def test_unsupported_module_raises():
    """``fail_mode="ERROR"`` of the reference's module map (``secondorder/vivit/__init__.py:83``)."""
    from vivit_b200 import SqrtGGNExact

    model = nn.Sequential(nn.Linear(4, 4), nn.Softplus(), nn.Linear(4, 2))
    x, y = torch.rand(3, 4), torch.randint(0, 2, (3,))
    with pytest.raises(NotImplementedError):
        run_backward(model, nn.CrossEntropyLoss(), x, y, [SqrtGGNExact()], None)


# ---- branched models, Pad, Slicing (vivit_b200.custom_module; reference fixture test/settings.py:160-181) ----


def _sqrt_ggn_matches_autograd(model, loss_fn, x, y):
    """``sum_p V_p V_p^T`` of the materialised factors against the autograd GGN."""
    from oracle.autograd_ggn import AutogradGGN
    from vivit_b200 import SqrtGGNExact

    run_backward(model, loss_fn, x, y, [SqrtGGNExact()], None)
    V = torch.cat([p.sqrt_ggn_exact.flatten(2) for p in model.parameters()], 2).flatten(0, 1)  # [C N, D]
    for p in model.parameters():
        del p.sqrt_ggn_exact
    want = AutogradGGN(model, loss_fn, x, y).ggn()
    assert torch.allclose(V.t() @ V, want, rtol=1e-9, atol=1e-12)


def test_two_branches_with_parameters_and_a_shared_factor():
    """Both branches hold parameters, one of them starts with modules that hand ``S`` through unchanged
    (Identity, Flatten): the sum at the shared input must not write into a buffer another tensor carries."""
    from vivit_b200.custom_module import Parallel

    torch.manual_seed(1)
    model = nn.Sequential(
        nn.Linear(5, 6),
        nn.Tanh(),
        Parallel(
            nn.Sequential(nn.Identity(), nn.Flatten(), nn.Linear(6, 4)),
            nn.Sequential(nn.Linear(6, 4), nn.Sigmoid()),
            nn.Linear(6, 4),
        ),
        Parallel(nn.Identity(), nn.Identity(), nn.Sequential(nn.Flatten(), nn.ReLU())),
        nn.Linear(4, 3),
    ).double()
    x, y = torch.rand(4, 5, dtype=torch.float64), torch.randint(0, 3, (4,))
    _sqrt_ggn_matches_autograd(model, nn.CrossEntropyLoss(), x, y)


def test_identity_between_layers_and_repeated_backward():
    """``nn.Identity`` returns its input object: the engine gives its output an identity of its own.
    A second backward through a fresh forward pass starts from clean tensors."""
    torch.manual_seed(2)
    model = nn.Sequential(nn.Linear(5, 4), nn.Identity(), nn.ReLU(), nn.Identity(), nn.Linear(4, 3)).double()
    x, y = torch.rand(3, 5, dtype=torch.float64), torch.randint(0, 3, (3,))
    for _ in range(2):
        _sqrt_ggn_matches_autograd(model, nn.CrossEntropyLoss(), x, y)


def test_pad_and_slicing_over_feature_maps():
    from vivit_b200.custom_module import Pad, Slicing

    torch.manual_seed(3)
    model = nn.Sequential(
        nn.Conv2d(2, 3, 3, padding=1),
        Pad((1, 0, 2, 1), value=-1.0),
        nn.Sigmoid(),
        Slicing((slice(None), slice(0, 2), slice(1, None, 2))),
        nn.Flatten(),
        nn.Linear(2 * 4 * 6, 3),
    ).double()
    x, y = torch.rand(3, 2, 5, 5, dtype=torch.float64), torch.randint(0, 3, (3,))
    _sqrt_ggn_matches_autograd(model, nn.CrossEntropyLoss(), x, y)


def test_pad_and_slicing_error_contract():
    from vivit_b200 import SqrtGGNExact
    from vivit_b200.custom_module import Pad, Slicing

    x, y = torch.rand(4, 6), torch.randint(0, 2, (4,))
    bad = [
        nn.Sequential(nn.Linear(6, 4), Pad((1, 1), mode="reflect"), nn.Linear(6, 2)),
        nn.Sequential(nn.Linear(6, 4), Pad((1, 1, 0, 0)), nn.Linear(6, 2)),  # would pad the batch axis
    ]
    for model in bad:
        with pytest.raises(NotImplementedError):
            run_backward(model, nn.CrossEntropyLoss(), x, y, [SqrtGGNExact()], None)
    model = nn.Sequential(nn.Linear(6, 4), Slicing((slice(0, 4, 1), slice(0, 2))), nn.Linear(2, 2))
    with pytest.raises(NotImplementedError):
        run_backward(model, nn.CrossEntropyLoss(), x, y, [SqrtGGNExact()], None)


def test_inplace_activation_returns_its_input_object():
    """``ReLU(inplace=True)`` hands back the tensor it was given, like ``Identity``: factors and per-sample
    gradients still match autograd."""
    from oracle.autograd_ggn import AutogradGGN
    from vivit_b200 import BatchGrad

    torch.manual_seed(4)
    model = nn.Sequential(nn.Linear(5, 4), nn.ReLU(inplace=True), nn.Linear(4, 3)).double()
    x, y = torch.rand(3, 5, dtype=torch.float64), torch.randint(0, 3, (3,))
    _sqrt_ggn_matches_autograd(model, nn.CrossEntropyLoss(), x, y)
    run_backward(model, nn.CrossEntropyLoss(), x, y, [BatchGrad()], None)
    got = torch.cat([p.grad_batch.flatten(1) for p in model.parameters()], 1)
    assert torch.allclose(got, AutogradGGN(model, nn.CrossEntropyLoss(), x, y).batch_grad(), rtol=1e-10, atol=1e-13)


def test_scale_module_in_a_weighted_skip_connection():
    """``ScaleModule`` ([BackPACK] maps it and ``Identity`` to the same handler, ``__init__.py:118-119``):
    host code against autograd, oracle against autograd."""
    from oracle import reference_path as ref
    from oracle.autograd_ggn import AutogradGGN
    from vivit_b200.custom_module import Parallel, ScaleModule

    torch.manual_seed(5)
    model = nn.Sequential(
        nn.Linear(5, 4),
        Parallel(ScaleModule(0.5), nn.Sequential(nn.Linear(4, 4), nn.Tanh(), ScaleModule(-2.0)), ScaleModule()),
        nn.Linear(4, 3),
    ).double()
    x, y = torch.rand(4, 5, dtype=torch.float64), torch.randint(0, 3, (4,))
    _sqrt_ggn_matches_autograd(model, nn.CrossEntropyLoss(), x, y)
    gram, _ = ref.gram_sqrt_ggn(model, nn.CrossEntropyLoss(), x, y)
    ggn = AutogradGGN(model, nn.CrossEntropyLoss(), x, y).ggn()
    n = min(gram.shape[0], ggn.shape[0])
    assert torch.allclose(torch.linalg.eigvalsh(gram)[-n:], torch.linalg.eigvalsh(ggn)[-n:], rtol=1e-8, atol=1e-12)


def test_one_dimensional_convolution_and_pooling():
    """``Conv1d`` / ``MaxPool1d`` / ``AvgPool1d`` (module map ``secondorder/vivit/__init__.py:90-101``) run on the
    2-d kernels over feature maps of unit height: factors and per-sample gradients against autograd."""
    from oracle.autograd_ggn import AutogradGGN
    from vivit_b200 import BatchGrad

    torch.manual_seed(6)
    model = nn.Sequential(
        nn.Conv1d(2, 3, 3, stride=2, padding=1),
        nn.ReLU(),
        nn.MaxPool1d(2, stride=1),
        nn.Conv1d(3, 4, 2, dilation=2, bias=False),
        nn.Tanh(),
        nn.AvgPool1d(2, stride=2, padding=1),
        nn.Flatten(),
        nn.Linear(4 * 3, 3),
    ).double()
    x, y = torch.rand(3, 2, 13, dtype=torch.float64), torch.randint(0, 3, (3,))
    assert model(x).shape == (3, 3)
    _sqrt_ggn_matches_autograd(model, nn.CrossEntropyLoss(), x, y)
    run_backward(model, nn.CrossEntropyLoss(), x, y, [BatchGrad()], None)
    got = torch.cat([p.grad_batch.flatten(1) for p in model.parameters()], 1)
    assert torch.allclose(got, AutogradGGN(model, nn.CrossEntropyLoss(), x, y).batch_grad(), rtol=1e-10, atol=1e-13)


def test_one_dimensional_layers_oracle_and_computation_agree():
    """The same 1-d net through ``EigvalshComputation`` (host code) and the oracle restatement."""
    from oracle import reference_path as ref
    from vivit_b200 import EigvalshComputation

    torch.manual_seed(7)
    model = nn.Sequential(
        nn.Conv1d(2, 3, 3, padding=1), nn.Sigmoid(), nn.MaxPool1d(2), nn.Conv1d(3, 2, 2, stride=2),
        nn.AvgPool1d(2, stride=1), nn.Flatten(), nn.Linear(2 * 2, 3),
    ).double()
    x, y = torch.rand(4, 2, 12, dtype=torch.float64), torch.randint(0, 3, (4,))
    for sub in (None, [2, 0]):
        groups = [{"params": list(model.parameters())}]
        comp = EigvalshComputation(subsampling=sub)
        run_backward(model, nn.CrossEntropyLoss(), x, y, [comp.get_extension()], comp.get_extension_hook(groups))
        (want,) = ref.eigvalsh(model, nn.CrossEntropyLoss(), x, y, groups, subsampling=sub)
        assert torch.allclose(comp.get_result(groups[0]), want, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("name", sorted(ND_NETS))
def test_conv3d_and_transposed_convolutions(name):
    """``Conv3d`` / ``ConvTranspose1d/2d/3d`` (module map ``secondorder/vivit/__init__.py:84-101``) are composed from
    the 2-d kernels (``backprop/conv_nd.py``): factors and per-sample gradients against autograd, eigenvalues
    through ``EigvalshComputation`` against the oracle restatement (with and without sub-sampling)."""
    from oracle import reference_path as ref
    from oracle.autograd_ggn import AutogradGGN
    from vivit_b200 import BatchGrad, EigvalshComputation

    torch.manual_seed(8)
    model, in_shape = ND_NETS[name]()
    model = model.double()
    x, y = torch.rand(*in_shape, dtype=torch.float64), torch.randint(0, 3, (in_shape[0],))
    assert model(x).shape == (in_shape[0], 3)
    _sqrt_ggn_matches_autograd(model, nn.CrossEntropyLoss(), x, y)
    run_backward(model, nn.CrossEntropyLoss(), x, y, [BatchGrad()], None)
    got = torch.cat([p.grad_batch.flatten(1) for p in model.parameters()], 1)
    assert torch.allclose(got, AutogradGGN(model, nn.CrossEntropyLoss(), x, y).batch_grad(), rtol=1e-10, atol=1e-13)
    for sub in (None, [2, 0]):
        groups = [{"params": list(model.parameters())}]
        comp = EigvalshComputation(subsampling=sub)
        run_backward(model, nn.CrossEntropyLoss(), x, y, [comp.get_extension()], comp.get_extension_hook(groups))
        (want,) = ref.eigvalsh(model, nn.CrossEntropyLoss(), x, y, groups, subsampling=sub)
        assert torch.allclose(comp.get_result(groups[0]), want, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("name", sorted(ND_NETS))
def test_conv3d_and_transposed_convolutions_sharded(name):
    """Parameter sharding (``process_group=``) of these layers: a rank owns a dim-0 slice of every parameter --
    output channels of a ``Conv3d`` weight, INPUT channels of a transposed-convolution weight -- and the slices
    of two ranks put together are the unsharded factor and per-sample gradients."""
    from vivit_b200 import BatchGrad, SqrtGGNExact

    torch.manual_seed(9)
    model, in_shape = ND_NETS[name]()
    model = model.double()
    x, y = torch.rand(*in_shape, dtype=torch.float64), torch.randint(0, 3, (in_shape[0],))

    def collect(shard):
        exts = [SqrtGGNExact(), BatchGrad()]
        for e in exts:
            e._shard = shard
        run_backward(model, nn.CrossEntropyLoss(), x, y, exts, None)
        out = [(p.sqrt_ggn_exact, p.grad_batch) for p in model.parameters()]
        for p in model.parameters():
            del p.sqrt_ggn_exact, p.grad_batch
        return out

    full, parts = collect(None), [collect((r, 2)) for r in range(2)]
    for i, (V, g) in enumerate(full):
        assert torch.allclose(torch.cat([parts[0][i][0], parts[1][i][0]], 2), V, rtol=1e-12, atol=1e-14)
        assert torch.allclose(torch.cat([parts[0][i][1], parts[1][i][1]], 1), g, rtol=1e-12, atol=1e-14)
