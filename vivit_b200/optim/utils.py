"""Extension factory of the optim computations (``vivit/optim/utils.py``)."""

from typing import List, Union

from vivit_b200.backprop.extensions import SqrtGGNExact, SqrtGGNMC


def get_sqrt_ggn_extension(
    subsampling: Union[None, List[int]], mc_samples: int, lazy: bool = False
) -> Union[SqrtGGNExact, SqrtGGNMC]:
    """``SqrtGGNExact`` for ``mc_samples == 0`` else ``SqrtGGNMC`` (``vivit/optim/utils.py:8-25``)."""
    return (
        SqrtGGNExact(subsampling=subsampling, lazy=lazy)
        if mc_samples == 0
        else SqrtGGNMC(subsampling=subsampling, mc_samples=mc_samples, lazy=lazy)
    )
