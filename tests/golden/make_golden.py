"""Regenerates the fixtures under ``tests/golden/``.  Run in the build container
(``/root/reference`` is not present on the GPU box):

    python tests/golden/make_golden.py

* ``symeig_killer.pt``: the reference's only binary test fixture
  (``/root/reference/test/utils/tensor_causes_symeig_error.pt``), a 128x128 fp32
  numerically rank-1 symmetric matrix on which LAPACK ``syevd`` fails to
  converge (``test/utils/test_stable_symeig.py:25-45``).  Re-saved unchanged.
* ``ground_truth.pt``: autograd ground truth (float64, ``oracle/autograd_ggn.py``)
  for the seeded problems of ``tests/problems.py``: GGN eigenvalues, per-sample
  directional derivatives and damped Newton steps for the parameter groupings
  and sub-samplings the reference's tests use.  The reference itself cannot run
  here (no BackPACK, no ``Tensor.symeig``), so these are the golden vectors of
  the path: the same quantities the reference's tests compare against.
"""

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.autograd_ggn import AutogradGGN  # noqa: E402
from tests.problems import GROUPING_IDS, GROUPINGS, PROBLEMS, constant_damping, make_top_k  # noqa: E402


def main():
    src = "/root/reference/test/utils/tensor_causes_symeig_error.pt"
    if os.path.exists(src):
        torch.save(torch.load(src).clone(), os.path.join(HERE, "symeig_killer.pt"))

    out = {}
    for problem in PROBLEMS:
        model, loss, x, y = problem.make(torch.float64)
        truth = AutogradGGN(model, loss, x, y)
        for gname, grouping in zip(GROUPING_IDS, GROUPINGS):
            for sub in (None, [1, 0]):
                groups = grouping(model, criterion=make_top_k(10), damping=constant_damping(1.0))
                evals, _ = truth.directions(groups, sub)
                gam, _ = truth.gammas(groups, sub, sub)
                lam = truth.lambdas(groups, sub)
                newton = truth.damped_newton(groups, sub, sub)
                all_evals = [
                    torch.linalg.eigvalsh(truth.ggn(sub)[idx][:, idx]) for idx in truth.group_indices(groups)
                ]
                out[(problem.name, gname, "full" if sub is None else "sub10")] = {
                    "evals_top10": evals,
                    "evals_all": all_evals,
                    "gammas_abs": [g.abs() for g in gam],
                    "lambdas": lam,
                    "newton": newton,
                }
    torch.save(out, os.path.join(HERE, "ground_truth.pt"))
    print(f"wrote {len(out)} cases")


if __name__ == "__main__":
    main()
