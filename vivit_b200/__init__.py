"""vivit_b200 -- B200-native low-rank GGN hot path with ViViT's BackPACK-facing API.

    from vivit_b200 import EighComputation, backpack, extend

    model, loss_fn = extend(model), extend(loss_fn)
    comp = EighComputation()
    groups = [{"params": list(model.parameters()), "criterion": lambda ev: [ev.numel() - 1]}]
    with backpack(comp.get_extension(), extension_hook=comp.get_extension_hook(groups)):
        loss_fn(model(X), y).backward()
    evals, evecs = comp.get_result(groups[0])

Every numerical step runs in hand-written sm_100a CUDA kernels reached through the C ABI
of ``include/vivit_b200.h``; there is no CPU or PyTorch fallback.
"""

from vivit_b200.backprop import (
    BatchGrad,
    SqrtGGNExact,
    SqrtGGNMC,
    ViViTGGNExact,
    ViViTGGNMC,
    backpack,
    disable,
    extend,
)
from vivit_b200 import custom_module, extensions, hessianfree
from vivit_b200.factors import set_conv_factor_streaming
from vivit_b200.linalg import EighComputation, EigvalshComputation, SolveQueue
from vivit_b200.optim import DirectionalDampedNewtonComputation, DirectionalDerivativesComputation

__version__ = "0.1.0"

__all__ = [
    "custom_module",
    "extensions",
    "hessianfree",
    "EigvalshComputation",
    "EighComputation",
    "DirectionalDerivativesComputation",
    "DirectionalDampedNewtonComputation",
    "SolveQueue",
    "set_conv_factor_streaming",
    "ViViTGGNExact",
    "ViViTGGNMC",
    "SqrtGGNExact",
    "SqrtGGNMC",
    "BatchGrad",
    "backpack",
    "extend",
    "disable",
]
