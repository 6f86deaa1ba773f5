#!/bin/bash
# round 2, call ab: alanine-dipeptide kernels after merging the primal and tangent passes over each weight matrix: parity + times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ad2.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r2ab_pytest_ad2.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2ab_ad2_times.txt
import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
from helpers import golden
from test_gpu_ad2 import _net
from pita_b200 import ops
g = golden("egnn_ad2_n22.npz")
net = _net(g); w = net.packed_weights("cuda")
for B in (148, 1184, 4736):
    x = torch.from_numpy(g["fk_x"]).float().cuda().repeat(B // 6 + 1, 1)[:B].contiguous()
    ht = torch.full((B,), 0.5, device="cuda"); beta = torch.full((B,), 0.75, device="cuda")
    for name, fn in (("forward", lambda: ops.egnn_forward(w, 64, 5, 22, ht, x, beta)),
                     ("energy", lambda: ops.egnn_energy(w, 64, 5, 22, ht, x, beta)),
                     ("score_div", lambda: ops.egnn_score_div(w, 64, 5, 22, ht, x, beta, need_div=True))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print("B=%d %-10s %.3f ms  -> %.1f particles/s" % (B, name, ms, B / ms * 1e3))
PY
