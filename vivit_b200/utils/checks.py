"""Validation of parameter groups and sub-sampling lists.

Same error contract as the reference (``vivit/utils/checks.py:6-49``): every
violation is a ``ValueError``.
"""

from typing import Dict, Iterable, List, Optional


def check_key_exists(param_groups: List[Dict], key: str) -> None:
    """Every group must carry ``key`` (``vivit/utils/checks.py:6-18``)."""
    for group in param_groups:
        if key not in group:
            raise ValueError(f"At least one group is not specifying '{key}'.")


def check_unique_params(param_groups: List[Dict]) -> None:
    """A parameter may belong to one group only (``vivit/utils/checks.py:21-35``)."""
    seen = set()
    for group in param_groups:
        for p in group["params"]:
            if id(p) in seen:
                raise ValueError("At least one parameter is in more than one group.")
            seen.add(id(p))


def check_subsampling_unique(subsampling: Optional[Iterable[int]]) -> None:
    """Sub-sampling indices must not repeat (``vivit/utils/checks.py:38-49``)."""
    if subsampling is None:
        return
    idx = list(subsampling)
    if len(set(idx)) != len(idx):
        raise ValueError("Detected repeated index in subsampling.")
