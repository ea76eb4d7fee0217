"""Two ranks on two GPUs over NCCL: the parameter-sharded Computations against the unsharded ones, with the
partial-Gram exchange in ``vvt_nccl_allreduce_gram``.  Skipped on a one-GPU box (the world-size-2 ``gloo`` tests in
``test_dist_cpu.py`` cover the host logic everywhere)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_computations_over_nccl():
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dist_gpu_worker.py")
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
         "127.0.0.1", "--master-port", "29631", worker],
        capture_output=True, text=True, timeout=600,
    )
    assert proc.returncode == 0 and "dist gpu worker ok" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-4000:]
