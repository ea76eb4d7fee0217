"""Matrix-free Hessian / GGN operators and Lanczos spectral densities.

Mirror of ``vivit/hessianfree/__init__.py:21-318`` (``HessianLinearOperator``, ``GGNLinearOperator`` as
SciPy ``LinearOperator``s over a data set), ``lanczos.py:13-270`` and ``utils.py:7-57``.  These are the
matrix-free cross-check of the spectra the Gram path computes: no ``R x R`` matrix is ever formed.

Differences to the reference, all additive:

* the matrix-vector products are plain ``torch.autograd`` double-backward products written here
  (BackPACK's ``hessian_vector_product`` / ``ggn_vector_product_from_plist`` are not available);
* ``matvec_torch`` keeps the vector on the device, and ``lanczos.fast_lanczos`` uses it to run the whole
  three-term recurrence on the GPU (one host read of the coefficients at the end) and to decompose the
  tridiagonal matrix with the library's own eigensolver (``vvt_syevj``) when the operator lives on a GPU.
"""

from vivit_b200.hessianfree.operators import (
    GGNLinearOperator,
    HessianLinearOperator,
    ggn_vector_product,
    hessian_vector_product,
)

__all__ = [
    "HessianLinearOperator",
    "GGNLinearOperator",
    "hessian_vector_product",
    "ggn_vector_product",
]
