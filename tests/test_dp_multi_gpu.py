"""Data-parallel parity on real GPUs: N ranks (one per visible GPU, up to 8) through the step launcher's NVLink
peer-memory tail must reproduce the single-process run over the same global batches (tools/dp_check.py: fp32 path, bf16
path, bf16 adapter path with learnable temperatures; weights within the stated tolerance of the single-process run and
bit-identical across the ranks).  Needs at least two GPUs in the box - skipped otherwise (the CPU suite covers the
host logic of the N > 1 path with gloo, tests/test_dp_gloo_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_n_ranks_reproduce_the_single_process_run():
    n = min(torch.cuda.device_count(), 8)
    env = dict(os.environ, UML_DP_TIMEOUT_S="10")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DP_CHECK_OK" in out.stdout, out.stdout[-3000:]
