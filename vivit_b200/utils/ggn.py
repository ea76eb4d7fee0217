"""Products with the GGN factor ``V`` stored per parameter (``vivit/utils/ggn.py``), on the back-transform
kernel of the CUDA library (``vvt_backtransform_dense``: the factor is streamed through once)."""

from __future__ import annotations

import math
from typing import Iterable, List, Optional, Union

import torch
from torch import Tensor

from vivit_b200 import kernels


def Vmp(V_t: Tensor, mat: Tensor, start_dim: int) -> Tensor:
    """``V`` applied to stacked Gram-space vectors: ``mat [F, *lead]``, ``V_t [*lead, *p.shape]`` ->
    ``[F, *p.shape]`` (``utils/ggn.py:94-115``)."""
    lead = math.prod(V_t.shape[:start_dim])
    out = kernels.backtransform_dense(
        mat.reshape(mat.shape[0], lead).contiguous(), V_t.reshape(lead, -1).contiguous(), None
    )
    return out.reshape(mat.shape[0], *V_t.shape[start_dim:])


def _get_V_t(param, savefield: str, subsampling: Optional[List[int]] = None) -> Tensor:
    """``V^T`` of a parameter, restricted to the listed samples (axis 1) (``utils/ggn.py:53-70``)."""
    V_t = getattr(param, savefield)
    if subsampling is not None:
        V_t = V_t.index_select(1, torch.as_tensor(subsampling, device=V_t.device))
    return V_t


def V_param_mat_prod(param, mat: Tensor, savefield: str, subsampling: Optional[List[int]] = None) -> Tensor:
    """``V_p @ mat`` for one parameter (``utils/ggn.py:73-91``)."""
    return Vmp(_get_V_t(param, savefield, subsampling=subsampling), mat, 2)


def V_mat_prod(
    mat: Tensor, parameters: Iterable, savefield: str, subsampling: Optional[List[int]] = None, concat: bool = False
) -> Union[List[Tensor], Tensor]:
    """``V @ mat`` for ``mat [F, C, N]``: a list of ``[F, *p.shape]`` tensors, or ``[F, D]`` if ``concat``
    (``utils/ggn.py:11-50``)."""
    assert mat.dim() == 3, f"mat must be [F, C, N]. Got {mat.dim()} dimensions."
    result = [V_param_mat_prod(p, mat, savefield, subsampling=subsampling) for p in parameters]
    if concat:
        return torch.cat([r.flatten(start_dim=1) for r in result], dim=1)
    return result
