"""N>1 path on real GPUs: world-size-2 NCCL run of the sharded resampler (peer-memory gather and all-gather
exchange) and of the sharded annealed loop, against the oracle.  Skipped on boxes with a single GPU; the host logic is
covered on CPU/gloo by tests/test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_world2_nccl_resample_and_loop():
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "MULTI_GPU_RESULT" in r.stdout and "error" not in r.stdout.split("MULTI_GPU_RESULT")[1], tail
