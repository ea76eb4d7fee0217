"""Extension factory of the optim computations (``vivit/optim/utils.py``)."""

from vivit_b200.backprop.extensions import factor_extension


def get_sqrt_ggn_extension(subsampling, mc_samples, lazy=False):
    """``SqrtGGNExact`` / ``SqrtGGNMC`` by ``mc_samples`` (``vivit/optim/utils.py:8-25``); ``lazy`` keeps the
    factor structured instead of materialising ``[C, N, *param.shape]``."""
    return factor_extension("sqrt_ggn", subsampling, mc_samples, lazy=lazy)
