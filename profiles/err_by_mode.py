"""Accuracy of the EGNN score / divergence kernels by arithmetic mode on the hardest parity case (n = 55, strong coordinate
gain, per-particle noise levels) — measured on the GPU against the fp64 oracle (checker only).  Prints one JSON line per mode."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import pita_oracle as O  # noqa: E402
from pita_b200 import ops  # noqa: E402
from pita_b200.egnn_temp_conditioned import pack_state_dict  # noqa: E402

for n, B, gain in ((55, 5, 0.3), (13, 37, 0.3)):
    sdS = O.random_egnn_state(seed=200 + n, dtype=torch.float64, coord_gain=gain)
    sched = O.EDMSchedule(0.05)
    t = torch.linspace(0.05, 0.98, B, dtype=torch.float64)
    ht = sched.h(t)
    x = O.centre(O.md_shaped_coords(B, n, seed=n, dtype=torch.float64) * (1 + ht.sqrt()[:, None] * 0.5), n)
    beta = 1.3
    s_ref = O.model_score(sdS, ht, x, beta, n)
    div_ref = O.exact_divergence(lambda h1, x1: O.model_score(sdS, h1, x1, beta, n), ht, x)
    s32 = O.model_score({k: v.float() for k, v in sdS.items()}, ht.float(), x.float(), beta, n).double()
    wS = pack_state_dict(sdS, 32, 3, "cuda")

    def errs(got, ref):
        got = got.double().cpu()
        el = ((got - ref).abs() / ref.abs().clamp_min(1.0)).max().item()
        return {"elementwise": el, "normwise": ((got - ref).norm() / ref.norm()).item()}

    print(json.dumps({"n": n, "mode": "torch fp32 on the CPU (the reference's arithmetic)", "score": errs(s32, s_ref)}))
    for mode in ("fp32", "3xtf32", "tf32"):
        s, d = ops.egnn_score_div(wS, 32, 3, n, ht.float().cuda(), x.float().cuda(), beta, mode=mode)
        print(json.dumps({"n": n, "mode": mode, "score": errs(s, s_ref), "div": errs(d, div_ref)}), flush=True)
