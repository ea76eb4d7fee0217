"""Seeded test problems.

The first block restates the reference's fixtures that the hot-path scope
covers (``/root/reference/test/settings.py:28-64``: 2-layer MLPs with
cross-entropy, Conv2d+MaxPool+Linear+Sigmoid, MLP with MSE).  Problems are
seeded with ``torch.manual_seed(0)`` exactly as ``test/problem.py:84-90`` does:
model first, then input, then target.
"""

from dataclasses import dataclass
from typing import Callable, List

import torch
from torch import nn

from vivit_b200.custom_module import Pad, Parallel, Slicing


@dataclass
class Problem:
    name: str
    input_fn: Callable[[], torch.Tensor]
    module_fn: Callable[[], nn.Module]
    loss_fn: Callable[[], nn.Module]
    target_fn: Callable[[], torch.Tensor]
    seed: int = 0

    def make(self, dtype=torch.float64, device="cpu"):
        torch.manual_seed(self.seed)
        model = self.module_fn()
        x = self.input_fn()
        y = self.target_fn()
        loss = self.loss_fn()
        model = model.to(device=device, dtype=dtype)
        x = x.to(device=device, dtype=dtype)
        if y.is_floating_point():
            y = y.to(dtype=dtype)
        y = y.to(device)
        return model, loss.to(device), x, y


def _cls(n, c):
    return lambda: torch.randint(size=(n,), low=0, high=c)


def _eval_mode(model):
    """``initialize_training_false_recursive`` (test/utils.py:81-114): random running statistics and
    affine parameters for every BatchNorm layer, whole model in evaluation mode."""
    for m in model.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.running_mean = torch.rand_like(m.running_mean)
            m.running_var = torch.rand_like(m.running_var)
            m.weight.data = torch.rand_like(m.weight)
            m.bias.data = torch.rand_like(m.bias)
    return model.train(False)


PROBLEMS: List[Problem] = [
    # test/settings.py:29-35
    Problem(
        "mlp-ce-mean",
        lambda: torch.rand(3, 7),
        lambda: nn.Sequential(nn.Linear(7, 6), nn.Linear(6, 5), nn.ReLU()),
        lambda: nn.CrossEntropyLoss(reduction="mean"),
        _cls(3, 5),
    ),
    # test/settings.py:42-54
    Problem(
        "conv-ce-mean",
        lambda: torch.rand(4, 3, 6, 6),
        lambda: nn.Sequential(
            nn.Conv2d(3, 2, kernel_size=3, stride=1, padding=1),
            nn.ReLU(),
            nn.MaxPool2d(kernel_size=3, stride=2),
            nn.Flatten(),
            nn.Linear(8, 5),
            nn.Sigmoid(),
        ),
        lambda: nn.CrossEntropyLoss(reduction="mean"),
        _cls(4, 5),
    ),
    # test/settings.py:56-63
    Problem(
        "mlp-mse-mean",
        lambda: torch.rand(3, 7),
        lambda: nn.Sequential(nn.Linear(7, 6), nn.Sigmoid(), nn.Linear(6, 5), nn.Sigmoid()),
        lambda: nn.MSELoss(reduction="mean"),
        lambda: torch.rand(3, 5),
    ),
    # extra coverage for the benchmark architectures: strided/padded convs,
    # ceil-mode max-pool (tf 'same' pooling of cifar10_3c3d), avg-pool + tanh,
    # several conv layers in a row (all-cnn-c style, global average pooling).
    Problem(
        "3c3d-mini",
        lambda: torch.rand(4, 2, 11, 11),
        lambda: nn.Sequential(
            nn.Conv2d(2, 3, kernel_size=3),
            nn.ReLU(),
            nn.MaxPool2d(kernel_size=3, stride=2, ceil_mode=True),
            nn.Conv2d(3, 4, kernel_size=3, padding=1),
            nn.ReLU(),
            nn.MaxPool2d(kernel_size=3, stride=2, ceil_mode=True),
            nn.Flatten(),
            nn.Linear(16, 6),
            nn.ReLU(),
            nn.Linear(6, 4),
        ),
        lambda: nn.CrossEntropyLoss(reduction="mean"),
        _cls(4, 4),
    ),
    Problem(
        "allcnnc-mini",
        lambda: torch.rand(5, 3, 8, 8),
        lambda: nn.Sequential(
            nn.Conv2d(3, 4, kernel_size=3, padding=1),
            nn.ReLU(),
            nn.Conv2d(4, 4, kernel_size=3, stride=2, padding=1),
            nn.ReLU(),
            nn.Conv2d(4, 6, kernel_size=1),
            nn.Tanh(),
            nn.AvgPool2d(kernel_size=4),
            nn.Flatten(),
        ),
        lambda: nn.CrossEntropyLoss(reduction="mean"),
        _cls(5, 6),
    ),
    # the remaining element-wise / padding modules of the reference's module map
    # (vivit/extensions/secondorder/vivit/__init__.py:95-117): ZeroPad2d, LeakyReLU, ELU, SELU, LogSigmoid
    Problem(
        "pad-acts",
        lambda: torch.rand(4, 2, 5, 5) - 0.3,
        lambda: nn.Sequential(
            nn.ZeroPad2d((1, 0, 2, 1)),
            nn.Conv2d(2, 3, kernel_size=3),
            nn.LeakyReLU(0.1),
            nn.Conv2d(3, 3, kernel_size=2),
            nn.ELU(alpha=0.7),
            nn.Flatten(),
            nn.Linear(45, 6),
            nn.SELU(),
            nn.Linear(6, 4),
            nn.LogSigmoid(),
        ),
        lambda: nn.CrossEntropyLoss(reduction="mean"),
        _cls(4, 4),
    ),
    # nn.Linear with one / two / three additional input dimensions (test/settings.py:66-112): the
    # reference materialises the weight factor of such layers (linear.py:26-27,38-39)
    Problem(
        "one-additional",
        lambda: torch.rand(3, 4, 5),
        lambda: nn.Sequential(nn.Linear(5, 3), nn.Sigmoid(), nn.Linear(3, 2), nn.Sigmoid(), nn.Flatten()),
        lambda: nn.MSELoss(reduction="mean"),
        lambda: torch.rand(3, 4 * 2),
    ),
    Problem(
        "two-additional",
        lambda: torch.rand(3, 4, 2, 5),
        lambda: nn.Sequential(nn.Linear(5, 3), nn.Tanh(), nn.Linear(3, 2), nn.Tanh(), nn.Flatten()),
        lambda: nn.MSELoss(reduction="mean"),
        lambda: torch.rand(3, 4 * 2 * 2),
    ),
    Problem(
        "three-additional",
        lambda: torch.rand(3, 4, 2, 3, 5),
        lambda: nn.Sequential(nn.Linear(5, 3), nn.ReLU(), nn.Linear(3, 2), nn.Sigmoid(), nn.Flatten()),
        lambda: nn.MSELoss(reduction="mean"),
        lambda: torch.rand(3, 4 * 2 * 3 * 2),
    ),
    # BatchNorm in evaluation mode (test/settings.py:118-160)
    Problem(
        "bn1d-mse",
        lambda: torch.rand(2, 3, 4),
        lambda: _eval_mode(nn.Sequential(nn.BatchNorm1d(num_features=3), nn.Flatten(), nn.Linear(12, 3), nn.Sigmoid())),
        lambda: nn.MSELoss(),
        lambda: torch.rand(2, 3),
    ),
    Problem(
        "bn2d-ce",
        lambda: torch.rand(3, 2, 4, 3),
        lambda: _eval_mode(nn.Sequential(nn.BatchNorm2d(num_features=2), nn.Flatten(), nn.Linear(24, 3))),
        lambda: nn.CrossEntropyLoss(),
        _cls(3, 3),
    ),
    Problem(
        "bn3d-ce",
        lambda: torch.rand(3, 3, 4, 1, 2),
        lambda: _eval_mode(nn.Sequential(nn.BatchNorm3d(num_features=3), nn.Flatten(), nn.Linear(24, 3))),
        lambda: nn.CrossEntropyLoss(),
        _cls(3, 3),
    ),
    Problem(
        "linear-bn3d-mse",
        lambda: torch.rand(3, 3, 4, 1, 2),
        lambda: _eval_mode(nn.Sequential(nn.Linear(2, 3), nn.BatchNorm3d(num_features=3), nn.Sigmoid(), nn.Flatten())),
        lambda: nn.MSELoss(),
        lambda: torch.rand(3, 4 * 1 * 3 * 3),
    ),
    # test/settings.py:160-181 -- branched model: skip connection around Linear -> Slicing, behind a Pad
    Problem(
        "branching-linear-slicing-pad",
        lambda: torch.rand(3, 7),
        lambda: nn.Sequential(
            nn.Linear(7, 4),
            nn.ReLU(),
            Pad((1, 1), mode="constant", value=0.5),
            Parallel(
                nn.Identity(),
                nn.Sequential(nn.Linear(6, 8), Slicing((slice(None), slice(0, 6)))),
            ),
            nn.Sigmoid(),
            nn.Linear(6, 4),
        ),
        lambda: nn.CrossEntropyLoss(),
        _cls(3, 4),
    ),
]

# Conv3d and transposed convolutions (module map ``secondorder/vivit/__init__.py:84-101``; no reference fixture
# contains one): ``name -> (model, input shape)``, three classes
def _conv3d_net():
    return nn.Sequential(
        nn.Conv3d(2, 3, (2, 3, 2), stride=(1, 2, 1), padding=(1, 1, 0)),
        nn.Tanh(),
        nn.Conv3d(3, 2, 2, dilation=(2, 1, 1), bias=False),
        nn.Sigmoid(),
        nn.Flatten(),
        nn.Linear(36, 3),
    ), (3, 2, 4, 6, 5)


def _conv_transpose2d_net():
    return nn.Sequential(
        nn.Conv2d(2, 3, 3, stride=2),
        nn.ReLU(),
        nn.ConvTranspose2d(3, 2, 3, stride=2, padding=1, output_padding=1),
        nn.Tanh(),
        nn.ConvTranspose2d(2, 2, (2, 3), dilation=(2, 1), bias=False),
        nn.Sigmoid(),
        nn.Flatten(),
        nn.Linear(2 * 8 * 8, 3),
    ), (3, 2, 7, 7)


def _conv_transpose1d_net():
    return nn.Sequential(
        nn.ConvTranspose1d(2, 3, 3, stride=2, padding=2),
        nn.Tanh(),
        nn.ConvTranspose1d(3, 2, 2, stride=3, output_padding=2, bias=False),
        nn.Flatten(),
        nn.Linear(2 * 22, 3),
    ), (4, 2, 5)


def _conv_transpose3d_net():
    return nn.Sequential(
        nn.ConvTranspose3d(2, 2, (2, 3, 2), stride=(2, 1, 2), padding=(0, 1, 1)),
        nn.Sigmoid(),
        nn.Conv3d(2, 2, 2, stride=2),
        nn.Flatten(),
        nn.Linear(2 * 2 * 1 * 1, 3),
    ), (3, 2, 2, 3, 2)


ND_NETS = {
    "conv3d": _conv3d_net,
    "conv-transpose2d": _conv_transpose2d_net,
    "conv-transpose1d": _conv_transpose1d_net,
    "conv-transpose3d": _conv_transpose3d_net,
}

# test/settings.py:36-41 -- only used at the extension level (the Computations
# require reduction='mean', see eigh.py:33-34)
PROBLEM_SUM = Problem(
    "mlp-ce-sum",
    lambda: torch.rand(3, 7),
    lambda: nn.Sequential(nn.Linear(7, 6), nn.ReLU(), nn.Linear(6, 5)),
    lambda: nn.CrossEntropyLoss(reduction="sum"),
    _cls(3, 5),
)

IDS = [p.name for p in PROBLEMS]


# -- criteria and groupings (test/linalg/settings.py:23-44, test/optim/settings.py:21-141)
def keep_all(evals):
    return list(range(evals.numel()))


def keep_nonzero(evals, min_abs=1e-4):
    return [i for i in range(evals.numel()) if evals[i].abs() >= min_abs]


def make_top_k(k, must_exceed=1e-5):
    def criterion(evals):
        n = len(evals)
        shift = 0 if n <= k else n - k
        return [i + shift for i, ev in enumerate(evals[shift:]) if ev > must_exceed]

    return criterion


def one_group(model, **extra):
    return [{"params": list(model.parameters()), **extra}]


def weights_and_biases(model, **extra):
    named = list(model.named_parameters())
    a = {"params": [p for n, p in named if "bias" in n], **extra}
    b = {"params": [p for n, p in named if "bias" not in n], **extra}
    return [a, b]


GROUPINGS = [one_group, weights_and_biases]
GROUPING_IDS = ["one", "weights_and_biases"]


def constant_damping(value):
    def damping(evals, evecs, gammas, lambdas):
        return value * torch.ones(gammas.shape[1], dtype=gammas.dtype, device=gammas.device)

    return damping
