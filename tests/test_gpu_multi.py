"""Multi-GPU global BA: the landmark-sharded solve (one all-reduce of the reduced camera system per LM iteration)
must reproduce the single-GPU solve — same iteration sequence, parameters and final residuals to 1e-8 — on the small cases and
on the full C5 configuration for its 20 iterations. Runs tools/mgpu_check.py under torch.distributed.run with every visible GPU
(2, 4 or 8); skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_solve_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    assert f"MGPU_CHECK PASS world {world}" in p.stdout, tail
