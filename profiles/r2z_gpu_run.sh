#!/bin/bash
# round 2, call z: LJ-13 kernel with batched staging loads: parity + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lj.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2z_pytest_lj.txt
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r2z_lj13.txt
import torch
from pita_b200 import ops
n = 13
for B in (65536, 1 << 20, 1 << 22, 1 << 24):
    x = torch.randn(B, 3 * n, device="cuda") * 1.5
    ops.lj_energy_force(x, n); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.lj_energy_force(x, n); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[3]
    tf = 2613 * B / (ms * 1e-3) / 1e12
    gbs = 316 * B / (ms * 1e-3) / 1e9
    print("n=13 B %9d  %.4f ms  %.3e configs/s  %.2f alg TFLOP/s (%.3f of 74.45)  %.0f alg GB/s (%.3f of 6535)" % (B, ms, B / ms * 1e3, tf, tf / 74.45, gbs, gbs / 6535))
PY
