"""Worker of ``tests/test_solver_variants_gpu.py``: one eigendecomposition with the solver variant selected by the
environment of this process (the switches are read once per process), checked against float64."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from vivit_b200 import kernels

    R, dtype = int(sys.argv[1]), {"f32": torch.float32, "f64": torch.float64}[sys.argv[2]]
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device="cpu").manual_seed(R)
    rank = int(0.8 * R)
    B = torch.randn(R, rank, dtype=torch.float64, generator=gen) * torch.logspace(0, -3, rank, dtype=torch.float64)
    G = (B @ B.t()).to(dtype).to(dev)
    evals, U, info = kernels.syevj(G, True, return_info=True)
    assert info["converged"], info
    want = torch.linalg.eigvalsh(G.double())
    scale = want.abs().max()
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    err = ((evals.double() - want).abs().max() / scale).item()
    assert err <= tol, ("eigenvalues", err)
    Q = U.double()
    orth = (Q.t() @ Q - torch.eye(R, dtype=torch.float64, device=dev)).abs().max().item()
    assert orth <= (5e-4 if dtype == torch.float32 else 1e-11), ("orthogonality", orth)
    resid = ((G.double() @ Q - Q * evals.double()[None]).norm() / G.double().norm()).item()
    assert resid <= (5e-5 if dtype == torch.float32 else 1e-12), ("residual", resid)
    print(f"variant ok: R={R} {sys.argv[2]} sweeps={info['sweeps']} evalerr={err:.2e} orth={orth:.2e} resid={resid:.2e}")


if __name__ == "__main__":
    main()
