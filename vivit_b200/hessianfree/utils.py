"""Low-rank helper operators (``vivit/hessianfree/utils.py:7-57``)."""

import numpy as np
from scipy.sparse.linalg import LinearOperator


class LowRank(LinearOperator):
    """``sum_i c_i a_i a_i^T`` with the vectors ``a_i`` stored column-wise in ``A [D, K]``."""

    def __init__(self, c: np.ndarray, A: np.ndarray):
        super().__init__(A.dtype, (A.shape[0], A.shape[0]))
        self._A, self._c = A, c

    def _matvec(self, x: np.ndarray) -> np.ndarray:
        x = x.reshape(-1)
        return self._A @ (self._c * (self._A.T @ x))


class Projector(LowRank):
    """Projector onto the span of the orthonormal columns of ``A [D, K]``."""

    def __init__(self, A: np.ndarray):
        super().__init__(np.ones(A.shape[1], dtype=A.dtype), A)
