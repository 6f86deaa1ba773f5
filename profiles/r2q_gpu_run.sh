#!/bin/bash
# round 2, call q: blocked LJ-55 kernel with vectorised staging: parity + timing; LJ chain micro-benchmark incl. straight-line variants;
# new GPU tests (generate_samples) and the racecheck re-run of the fp32 engine after the __syncwarp fix
mkdir -p gpurun_out
PITA_LJ_CFG=7 timeout 600 python -m pytest tests/test_gpu_lj.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2q_pytest_lj_cfg7.txt
rm -f gpurun_out/r2q_lj_ab.txt
for cfg in 0 7; do
PITA_LJ_CFG=$cfg timeout 300 python - <<'PY' 2>&1 | tee -a gpurun_out/r2q_lj_ab.txt
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
timeout 300 ./profiles/ubench/lj_chain.bin 2>&1 | tail -20 | tee gpurun_out/r2q_ubench_lj_chain_straightline.txt
timeout 900 python -m pytest tests/test_gpu_sde.py -q -m gpu -k "generate_samples" 2>&1 | tail -5 | tee gpurun_out/r2q_pytest_generate_samples.txt
