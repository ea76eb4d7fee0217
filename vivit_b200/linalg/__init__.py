"""GGN spectra during back-propagation."""

from vivit_b200.linalg.eigh import EighComputation
from vivit_b200.linalg.eigvalsh import EigvalshComputation
from vivit_b200.linalg.solve_queue import SolveQueue

__all__ = ["EighComputation", "EigvalshComputation", "SolveQueue"]
