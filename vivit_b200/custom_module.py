"""Modules for branched models, padding and slicing.

[BackPACK] ships these as ``backpack.custom_module.{branching,pad,slicing}``; the reference's
last fixture (``test/settings.py:160-181``, a skip connection around ``Linear -> Slicing``
behind a ``Pad``) is built from them.  They are ordinary ``torch.nn`` modules; the factor
back-propagation through them lives in ``vivit_b200.backprop.extensions``:

* ``ScaleModule`` -- the factor is scaled by the same constant;
* ``SumModule``   -- the factor reaches every summand unchanged;
* ``Parallel``    -- container: every branch sees the same input, the outputs are merged
  (summed); the input's factor is the SUM of the branches' factors
  (``ViViTGGN.accumulate_backpropagated_quantities``, ``secondorder/vivit/__init__.py:130-133``);
* ``Pad``         -- constant padding of trailing dimensions: the factor is cropped;
* ``Slicing``     -- ``input[slice_info]``: the factor is embedded in zeros.
"""

from __future__ import annotations

from typing import Sequence, Tuple, Union

import torch.nn.functional as F
from torch import Tensor, nn


class ScaleModule(nn.Module):
    """Multiply the input by a constant ([BackPACK] ``custom_module.scale_module.ScaleModule``)."""

    def __init__(self, weight: float = 1.0):
        super().__init__()
        if not isinstance(weight, (int, float)):
            raise ValueError(f"weight must be a number, got {type(weight)}")
        self.weight = float(weight)

    def forward(self, input: Tensor) -> Tensor:
        return input * self.weight

    def extra_repr(self) -> str:
        return f"weight={self.weight}"


class SumModule(nn.Module):
    """Sum of all positional inputs."""

    def forward(self, *inputs: Tensor) -> Tensor:
        result = inputs[0]
        for other in inputs[1:]:
            result = result + other
        return result


class Parallel(nn.Module):
    """Feed one input to several branches and merge their outputs (``SumModule`` by default)."""

    def __init__(self, *branches: nn.Module, merge_module: nn.Module = None):
        super().__init__()
        for idx, branch in enumerate(branches):
            self.add_module(str(idx), branch)
        self._n_branches = len(branches)
        self.merge_module = SumModule() if merge_module is None else merge_module

    def branches(self):
        return [getattr(self, str(idx)) for idx in range(self._n_branches)]

    def forward(self, input: Tensor) -> Tensor:
        return self.merge_module(*[branch(input) for branch in self.branches()])


class Pad(nn.Module):
    """``torch.nn.functional.pad`` as a module; only ``mode='constant'`` has a factor rule."""

    def __init__(self, pad: Sequence[int], mode: str = "constant", value: float = 0.0):
        super().__init__()
        self.pad = tuple(int(p) for p in pad)
        self.mode = mode
        self.value = value

    def forward(self, input: Tensor) -> Tensor:
        return F.pad(input, self.pad, mode=self.mode, value=self.value)

    def extra_repr(self) -> str:
        return f"pad={self.pad}, mode={self.mode!r}, value={self.value}"


class Slicing(nn.Module):
    """``input[slice_info]``; ``slice_info`` holds one ``slice`` or ``int`` per leading axis (batch first)."""

    def __init__(self, slice_info: Tuple[Union[slice, int], ...]):
        super().__init__()
        self.slice_info = tuple(slice_info)

    def forward(self, input: Tensor) -> Tensor:
        return input[self.slice_info]

    def extra_repr(self) -> str:
        return f"slice_info={self.slice_info}"
