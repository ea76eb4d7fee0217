"""Validation of parameter groups and sub-sampling lists.

Same error contract as the reference (``vivit/utils/checks.py:6-49``): every
violation is a ``ValueError`` with the reference's message.
"""

from collections import Counter
from typing import Dict, Iterable, List, Optional


def _fail_if(condition: bool, message: str) -> None:
    if condition:
        raise ValueError(message)


def check_key_exists(param_groups: List[Dict], key: str) -> None:
    """Every group must carry ``key`` (``vivit/utils/checks.py:6-18``)."""
    missing = [idx for idx, group in enumerate(param_groups) if key not in group]
    _fail_if(bool(missing), f"At least one group is not specifying '{key}'.")


def check_unique_params(param_groups: List[Dict]) -> None:
    """A parameter may belong to one group only (``vivit/utils/checks.py:21-35``)."""
    uses = Counter(id(p) for group in param_groups for p in group["params"])
    _fail_if(any(n > 1 for n in uses.values()), "At least one parameter is in more than one group.")


def check_subsampling_unique(subsampling: Optional[Iterable[int]]) -> None:
    """Sub-sampling indices must not repeat (``vivit/utils/checks.py:38-49``)."""
    uses = Counter(() if subsampling is None else subsampling)
    _fail_if(any(n > 1 for n in uses.values()), "Detected repeated index in subsampling.")
