#!/bin/bash
# round 2, call n: LJ-55 blocked kernel (8 warps per CTA, 16 per SM) vs the 5-warp circulant kernel
mkdir -p gpurun_out
for cfg in 7; do
  PITA_LJ_CFG=$cfg timeout 600 python -m pytest tests/test_gpu_lj.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2n_pytest_lj_cfg$cfg.txt
done
for cfg in 0 7; do
PITA_LJ_CFG=$cfg timeout 300 python - <<'PY' 2>&1 | tee -a gpurun_out/r2n_lj_ab.txt
import os, torch
from pita_b200 import ops
n = 55
for B in (262144, 1 << 20, 1 << 22):
    x = torch.randn(B, 3 * n, device="cuda") * 1.5
    ops.lj_energy_force(x, n); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.lj_energy_force(x, n); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    tf = 46860 * B / (ms * 1e-3) / 1e12
    print("cfg %s B %8d  %.3f ms  %.2f alg TFLOP/s  frac of 74.45 = %.3f" % (os.environ.get("PITA_LJ_CFG"), B, ms, tf, tf / 74.45))
PY
done
