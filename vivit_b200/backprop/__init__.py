"""A small BackPACK-protocol backprop engine.

The reference reaches its hot path through BackPACK's extension protocol
(SURVEY 8b): ``extend(module)`` makes modules remember ``input0``/``output``,
``with backpack(*extensions, extension_hook=hook): loss.backward()`` runs, for
every module in reverse order, each extension's module handler (back-propagate
the extension's quantity, fill the per-parameter ``savefield``) and then
``extension_hook(module)``.  BackPACK is not installable next to current torch
releases here, and its module handlers are exactly the kernels this project
replaces, so the protocol is re-implemented natively:

* ``extend`` registers a forward hook that stores ``input0``/``output`` and puts
  a tensor hook on the module output; the tensor hook fires during
  ``loss.backward()`` when the gradient w.r.t. that output is known, i.e. in
  reverse execution order, before the module's own backward;
* a second-order quantity travels from ``module.output`` to ``module.input0`` as a
  tensor attribute (the same tensor object is the next module's ``output``); a tensor
  consumed by several modules (branched models, ``vivit_b200.custom_module.Parallel``)
  receives the SUM of their quantities -- its own hook fires only after autograd has
  visited all its consumers;
* handlers call ``vivit_b200.kernels`` (the C ABI); nothing falls back to torch
  arithmetic.

Only leaf modules (no children) are handled; containers are transparent, which
matches the reference's hook skipping ``Sequential`` (``vivit/utils/hooks.py:70``).
"""

from __future__ import annotations

import weakref
from typing import Callable, Optional

import torch
from torch import nn

from vivit_b200.backprop.extensions import (
    BatchGrad,
    Extension,
    SqrtGGNExact,
    SqrtGGNMC,
    ViViTGGNExact,
    ViViTGGNMC,
)

__all__ = [
    "backpack", "extend", "disable", "BatchGrad", "SqrtGGNExact", "SqrtGGNMC",
    "ViViTGGNExact", "ViViTGGNMC", "Extension",
]


class _State:
    extensions = ()
    extension_hook: Optional[Callable[[nn.Module], None]] = None
    store_io = True


_state = _State()


class backpack:
    """Context manager activating extensions for the backward pass(es) inside it.

    Mirrors ``backpack.backpack(*exts, extension_hook=None, debug=False)`` [BackPACK].
    """

    def __init__(self, *exts: Extension, extension_hook=None, debug: bool = False, retain_graph: bool = False):
        for ext in exts:
            if not isinstance(ext, Extension):
                raise ValueError(
                    f"Expected instances of a vivit_b200 extension, got {type(ext)}. "
                    "Did you forget to instantiate the class?"
                )
        if extension_hook is not None and not callable(extension_hook):
            raise ValueError("extension_hook must be callable or None")
        self.exts = exts
        self.extension_hook = extension_hook
        self.debug = debug

    def __enter__(self):
        self._old = (_state.extensions, _state.extension_hook)
        _state.extensions, _state.extension_hook = self.exts, self.extension_hook
        for ext in self.exts:
            ext._begin()
        return self

    def __exit__(self, *exc):
        _state.extensions, _state.extension_hook = self._old
        return False


class disable:
    """Context manager: do not store module inputs/outputs in the forward pass."""

    def __enter__(self):
        self._old = _state.store_io
        _state.store_io = False

    def __exit__(self, *exc):
        _state.store_io = self._old
        return False


def _backward_hook(module_ref, grad_output):
    module = module_ref()
    if module is None or not _state.extensions and _state.extension_hook is None:
        return None
    with torch.no_grad():
        for ext in _state.extensions:
            ext._apply(module, grad_output)
        if _state.extension_hook is not None:
            _state.extension_hook(module)
    return None


def _forward_hook(module, inputs, output):
    if not (_state.store_io and torch.is_grad_enabled()):
        return
    if next(module.children(), None) is not None:  # containers are transparent
        return
    for i, inp in enumerate(inputs):
        setattr(module, f"input{i}", inp)
    aliased = isinstance(output, torch.Tensor) and any(output is inp for inp in inputs)
    if aliased:
        # nn.Identity hands its input through: give the output an identity of its own, so that its
        # tensor hook and its factor attribute are not the producer's (hooks fire in registration order)
        output = output.view_as(output)
    module.output = output
    if isinstance(output, torch.Tensor) and output.requires_grad:
        ref = weakref.ref(module)
        output.register_hook(lambda g, ref=ref: _backward_hook(ref, g))
    return output if aliased else None


def extend(module: nn.Module, debug: bool = False, use_converter: bool = False) -> nn.Module:
    """Make ``module`` (recursively) usable inside ``with backpack(...)`` [BackPACK ``extend``]."""
    if use_converter:
        raise NotImplementedError("use_converter is not supported")
    for child in module.children():
        extend(child, debug=debug)
    if not getattr(module, "_vivit_b200_extended", False):
        module._vivit_b200_extended = True
        module.register_forward_hook(_forward_hook)
    return module
