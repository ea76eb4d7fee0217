"""Module map edge cases on CPU through the test double."""

import pytest
import torch
from torch import nn

import tests._torch_kernels as double
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
