#!/bin/bash
# round 2, call u: Laplacian branch (csrc/egnn_lap.cu): parity tests + kernel rate
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_laplacian.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2u_pytest_laplacian.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2u_laplacian_rate.txt
import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import pita_oracle as O
from helpers import make_net
from pita_b200 import ops
for n, B in ((13, 4096), (55, 296)):
    net = make_net(n, O.random_egnn_state(seed=1, dtype=torch.float64))
    w = net.packed_weights("cuda")
    x = ops.remove_mean(torch.randn(B, 3 * n, device="cuda") * 2, n)
    ht = torch.full((B,), 3.0, device="cuda"); beta = torch.full((B,), 0.8, device="cuda")
    ops.egnn_energy_laplacian(w, 32, 3, n, ht, x, beta); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.egnn_energy_laplacian(w, 32, 3, n, ht, x, beta); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("n=%d B=%d  %.2f ms  -> %.0f particles/s" % (n, B, ms, B / ms * 1e3))
PY
