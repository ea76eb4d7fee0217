"""Helpers shared by the linalg and optim computations."""

from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
from torch import Tensor
from torch.nn import Module

from vivit_b200 import kernels
from vivit_b200.backprop.extensions import factor_extension


def get_vivit_extension(subsampling, mc_samples, shard=None):
    """``ViViTGGNExact`` / ``ViViTGGNMC`` by ``mc_samples`` (``vivit/linalg/utils.py:11-28``); ``shard`` =
    ``(rank, world)`` restricts every parameter to this rank's dim-0 slice (``vivit_b200.dist``)."""
    ext = factor_extension("vivit", subsampling, mc_samples)
    ext._shard = shard
    return ext


def get_hook_store_batch_size(
    param_groups: List[Dict], destination: Dict[int, int], verbose: bool = False
) -> Callable[[Module], None]:
    """Hook recording ``N = module.input0.shape[0]`` once per backward pass for all
    groups (``vivit/linalg/utils.py:31-64``)."""

    def hook_store_batch_size(module: Module) -> None:
        if destination == {}:
            batch_size = module.input0.shape[0]
            for group in param_groups:
                if verbose:
                    print(f"Group {id(group)}: Store 'batch_size'")
                destination[id(group)] = batch_size

    return hook_store_batch_size


def normalize(tensors: List[Tensor], norm2: Optional[Tensor] = None) -> None:
    """Scale stacked vectors in parameter-list format to unit norm, in place
    (``vivit/linalg/utils.py:67-76``).  ``norm2`` (float64 ``[K]``) holds the squared norms
    already accumulated by the back-transform kernels."""
    if not tensors:
        return
    if norm2 is None:
        K = tensors[0].shape[0]
        norm2 = torch.zeros(K, dtype=torch.float64, device=tensors[0].device)
        for t in tensors:
            norm2 += (t.double() ** 2).flatten(1).sum(1)
    for t in tensors:
        kernels.scale_rows_rsqrt(t, norm2)
