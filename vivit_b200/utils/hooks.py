"""Parameter-group bookkeeping for extension hooks.

Behavioural restatement of ``vivit/utils/hooks.py:11-405`` (``ModuleHook`` ->
``ParameterHook`` -> ``ParameterGroupsHook``) as one flat class:

* called once per module by the backprop engine right after the extensions ran;
* walks ``module.parameters()``, skipping containers (``Sequential``,
  ``hooks.py:70``), parameters without ``requires_grad``, parameters outside
  every group and parameters already seen (``hooks.py:292-307``);
* per parameter: ``param_computation`` -> ``accumulate`` into the group's
  running result (``hooks.py:250-263,332-345``);
* as soon as the last parameter of a group has arrived, ``group_hook`` fires
  (``hooks.py:309-330,265-277``).

Like the reference's, a hook object is single-use: the set of processed
parameters is never reset (``hooks.py:45,73``); mint a new hook per backward pass.
"""

from __future__ import annotations

from typing import Any, Callable, Dict, List

from torch.nn import Module, Parameter, Sequential

_MISSING = object()


class ParameterGroupsHook:
    """Accumulate per-parameter results group-wise, then post-process each group."""

    def __init__(self, param_groups: List[Dict[str, Any]]):
        ids = [id(p) for g in param_groups for p in g["params"]]
        if len(ids) != len(set(ids)):
            raise ValueError("Same parameters occur in different groups")
        self.savefield = None
        self.processed = set()
        self._param_groups = param_groups
        self._group_of = {id(p): g for g in param_groups for p in g["params"]}
        self._pending = {id(g): {id(p) for p in g["params"]} for g in param_groups}
        self._accumulations: Dict[int, Any] = {}
        self._output: Dict[int, Any] = {}
        self._processed_groups = set()

    # -- to be provided (subclass or from_functions) -------------------------
    def param_computation(self, param: Parameter) -> Any:
        raise NotImplementedError

    def accumulate(self, existing: Any, update: Any) -> Any:
        raise NotImplementedError

    def group_hook(self, accumulation: Any, group: Dict[str, Any]) -> Any:
        raise NotImplementedError

    # -- engine-facing -----------------------------------------------------------
    def __call__(self, module: Module) -> None:
        if isinstance(module, Sequential):
            return
        for param in module.parameters():
            if self.should_run_hook(param, module):
                self.run_hook(param, module)

    def should_run_hook(self, param: Parameter, module: Module) -> bool:
        if isinstance(module, Sequential):
            return False
        return (
            id(param) in self._group_of
            and id(param) not in self.processed
            and param.requires_grad
        )

    def run_hook(self, param: Parameter, module: Module) -> None:
        group = self._group_of[id(param)]
        gid = id(group)
        result = self.param_computation(param)
        existing = self._accumulations.get(gid, _MISSING)
        self._accumulations[gid] = (
            result if existing is _MISSING else self.accumulate(existing, result)
        )
        pending = self._pending[gid]
        pending.discard(id(param))
        if not pending:
            self._output[gid] = self.group_hook(self._accumulations.pop(gid), group)
            self._processed_groups.add(gid)
        self.processed.add(id(param))

    def get_output(self, group: Dict[str, Any], pop: bool = True) -> Any:
        if not all(id(p) in self.processed for p in group["params"]):
            raise ValueError("Group contains unprocessed parameters.")
        return self._output.pop(id(group)) if pop else self._output[id(group)]

    @classmethod
    def from_functions(
        cls,
        param_groups: List[Dict[str, Any]],
        param_computation_fn: Callable[["ParameterGroupsHook", Parameter], Any],
        group_hook_fn: Callable[["ParameterGroupsHook", Any, Dict[str, Any]], Any],
        accumulate_fn: Callable[["ParameterGroupsHook", Any, Any], Any],
    ) -> "ParameterGroupsHook":
        """Build a hook from three free functions taking the hook as first argument
        (``hooks.py:361-390``)."""
        # (the functions are kept as plain attributes and called with the hook as first argument: binding them as
        # methods of the instance would tie the hook, the closures and everything they capture -- the computation
        # object with its results, factors with their activations -- into a reference cycle that only the cyclic
        # garbage collector frees, gigabytes and several steps later)
        hook = _FunctionHook(param_groups)
        hook._fns = (param_computation_fn, group_hook_fn, accumulate_fn)
        return hook


class _FunctionHook(ParameterGroupsHook):
    """``ParameterGroupsHook`` whose three operations are free functions taking the hook as first argument."""

    _fns = (None, None, None)

    def param_computation(self, param: Parameter) -> Any:
        return self._fns[0](self, param)

    def group_hook(self, accumulation: Any, group: Dict[str, Any]) -> Any:
        return self._fns[1](self, accumulation, group)

    def accumulate(self, existing: Any, update: Any) -> Any:
        return self._fns[2](self, existing, update)


class ModuleHook:
    """Per-parameter hook with access to the parameter and its module
    (``vivit/utils/hooks.py:11-109``): runs once on every trainable parameter of every
    non-``Sequential`` module and stores the returned value under ``param.<savefield>``."""

    def __init__(self, savefield: str = None):
        self.savefield = savefield
        self.processed = set()

    def module_hook(self, param: Parameter, module: Module) -> Any:
        raise NotImplementedError

    def __call__(self, module: Module) -> None:
        for param in module.parameters():
            if self.should_run_hook(param, module):
                self.run_hook(param, module)

    def should_run_hook(self, param: Parameter, module: Module) -> bool:
        if isinstance(module, Sequential):
            return False
        return id(param) not in self.processed and param.requires_grad

    def run_hook(self, param: Parameter, module: Module) -> None:
        value = self.module_hook(param, module)
        self._save(value, param)
        self.processed.add(id(param))

    def _save(self, value: Any, param: Parameter) -> None:
        should_save = self.savefield is not None
        if value is not None and not should_save:
            raise ValueError(f"Hook has no savefield, but produced output of type {type(value)}.")
        if should_save:
            setattr(param, self.savefield, value)


class ParameterHook(ModuleHook):
    """Per-parameter hook that only needs the parameter (``vivit/utils/hooks.py:112-143``)."""

    def param_hook(self, param: Parameter) -> Any:
        raise NotImplementedError

    def module_hook(self, param: Parameter, module: Module) -> Any:
        return self.param_hook(param)
