"""Extension hooks on [BackPACK]-shaped per-parameter quantities (``vivit/extensions/__init__.py``)."""

from vivit_b200.extensions import hooks

__all__ = ["hooks"]
