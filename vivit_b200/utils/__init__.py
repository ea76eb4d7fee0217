"""Small helpers shared by the computations."""

from typing import Optional

from torch import Tensor


def delete_savefield(param: Tensor, savefield: str, verbose: Optional[bool] = False) -> None:
    """Drop a per-parameter buffer (``vivit/utils/__init__.py:8-19``)."""
    if verbose:
        print(f"Param {id(param)}: Delete '{savefield}'")
    delattr(param, savefield)
