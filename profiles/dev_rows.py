"""Development driver (not a benchmark): checks / times the row-engine EGNN kernels on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import pita_oracle as O
from pita_b200 import ops
from pita_b200.egnn_temp_conditioned import pack_state_dict

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

for n, B in (() if what == "prof" else ((13, 37), (55, 5))):
    sd = O.random_egnn_state(seed=200 + n, dtype=torch.float64, coord_gain=0.3)
    w = pack_state_dict(sd, 32, 3, "cuda")
    sched = O.EDMSchedule(0.05)
    t = torch.linspace(0.05, 0.98, B, dtype=torch.float64)
    ht = sched.h(t)
    x = O.centre(O.md_shaped_coords(B, n, seed=n, dtype=torch.float64) * (1 + ht.sqrt()[:, None] * 0.5), n)
    beta = torch.full((B,), 1.3, dtype=torch.float64)
    if what in ("fwd", "all"):
        tc = torch.log(ht) / 8
        y = x / torch.sqrt(1 + ht)[:, None]
        ref = O.egnn_velocity(sd, tc, y, beta, n)
        got = ops.egnn_forward(w, 32, 3, n, tc.float().cuda(), y.float().cuda(), beta.float().cuda()).double().cpu()
        err = (got - ref).abs() / ref.abs().clamp(min=1.0)
        print("fwd n=%d max rel err %.3e  (|ref| max %.3e)" % (n, err.max().item(), ref.abs().max().item()), flush=True)
    if what in ("div", "all"):
        s_ref = O.model_score(sd, ht, x, 1.3, n)
        div_ref = O.exact_divergence(lambda h1, x1: O.model_score(sd, h1, x1, 1.3, n), ht, x)
        for mode in ("3xtf32", "tf32"):
            s, d = ops.egnn_score_div(w, 32, 3, n, ht.float().cuda(), x.float().cuda(), 1.3, mode=mode)
            es = ((s.double().cpu() - s_ref).abs() / s_ref.abs().clamp(min=1.0)).max().item()
            ed = ((d.double().cpu() - div_ref).abs() / div_ref.abs().clamp(min=1.0)).max().item()
            print("div n=%d mode=%s score err %.3e div err %.3e" % (n, mode, es, ed), flush=True)
            if ed > 1e-2:
                print("  got", d.cpu().numpy()[:5], "ref", div_ref.numpy()[:5])

for n, B in (() if what == "prof" else ((13, 1 << 18), (55, 1 << 14))):
    sd = O.random_egnn_state(seed=1, dtype=torch.float64, coord_gain=0.3)
    w = pack_state_dict(sd, 32, 3, "cuda")
    x = O.centre(O.md_shaped_coords(B, n, seed=3) * 1.5, n).cuda()
    ht = torch.full((B,), 2.0, device="cuda"); be = torch.full((B,), 1.0, device="cuda")
    if what in ("fwd", "all"):
        ms = timeit(lambda: ops.egnn_forward(w, 32, 3, n, ht, x, be))
        print("fwd  n=%d B=%d: %.2f ms  (%.3f us/particle)" % (n, B, ms, ms * 1e3 / B), flush=True)
    if what in ("div", "all"):
        for mode in ("3xtf32", "tf32"):
            ms = timeit(lambda: ops.egnn_score_div(w, 32, 3, n, ht, x, 1.0, mode=mode))
            print("div  n=%d B=%d mode=%s: %.2f ms  (%.3f us/particle)" % (n, B, mode, ms, ms * 1e3 / B), flush=True)
    if what in ("energy", "all"):
        ms = timeit(lambda: ops.egnn_energy(w, 32, 3, n, ht, x, be))
        print("energy n=%d B=%d: %.2f ms  (%.3f us/particle)" % (n, B, ms, ms * 1e3 / B), flush=True)

if what == "prof":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    mode = sys.argv[3] if len(sys.argv) > 3 else "3xtf32"
    B = 148 * 18 if n == 13 else 148 * 4
    sd = O.random_egnn_state(seed=1, dtype=torch.float64, coord_gain=0.3)
    w = pack_state_dict(sd, 32, 3, "cuda")
    x = O.centre(O.md_shaped_coords(B, n, seed=3) * 1.5, n).cuda()
    ht = torch.full((B,), 2.0, device="cuda"); be = torch.full((B,), 1.0, device="cuda")
    for _ in range(2):
        ops.egnn_forward(w, 32, 3, n, ht, x, be)
        ops.egnn_score_div(w, 32, 3, n, ht, x, 1.0, mode=mode)
        ops.egnn_energy(w, 32, 3, n, ht, x, be)
    torch.cuda.synchronize()
    print("prof ok")
