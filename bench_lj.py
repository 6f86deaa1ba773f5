#!/usr/bin/env python
"""BASELINE.json configs[4]: Lennard-Jones pairwise energy+force micro-benchmark sweep (roofline vs host CPU).

  python bench_lj.py [--n 55] [--batches 1024,...] [--kernel paired|ordered] [--cpu]

One JSON line per batch size: configurations/s, algorithmic TFLOP/s (31 FLOP per unordered pair + 15 per atom),
fraction of the FP32 FMA peak (148 SMs x 128 lanes x 2 x sm clock) and of the HBM copy peak (MEASURED_PEAKS.json).
Inputs live in HBM; every timed launch is preceded by an L2 flush when the working set is smaller than L2.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=55)
    ap.add_argument("--batches", default="1024,16384,262144,1048576,4194304")
    ap.add_argument("--kernel", default="paired", choices=["paired", "ordered"])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (autograd force) on a bounded sample")
    args = ap.parse_args()
    os.environ["PITA_LJ_KERNEL"] = args.kernel
    import torch
    from pita_b200 import ops

    n, D = args.n, 3 * args.n
    flop_cfg = 31 * n * (n - 1) // 2 + 15 * n
    bytes_cfg = 4 * D + 4 * D + 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    hbm_peak = float(peaks.get("hbm_gbs", 6500.0))
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(7)
    for B in [int(b) for b in args.batches.split(",")]:
        # MD-shaped coordinates: simple-cubic lattice sites, spacing 1.1, jitter 0.08, COM removed
        side = int(round(n ** (1 / 3))) + 1
        sites = torch.stack(torch.meshgrid(*[torch.arange(side, dtype=torch.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] * 1.1
        x = sites.to(dev).reshape(1, D).repeat(B, 1) + 0.08 * torch.randn(B, D, device=dev, generator=gen)
        x = ops.remove_mean(x, n)
        for _ in range(3):
            ops.lj_energy_force(x, n)
        ts = []
        for _ in range(args.reps):
            if B * bytes_cfg < (200 << 20):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lp, f = ops.lj_energy_force(x, n)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        tf = flop_cfg * B / (ms * 1e-3) / 1e12
        gbs = bytes_cfg * B / (ms * 1e-3) / 1e9
        line = {"bench": "lj_energy_force", "kernel": args.kernel, "n_atoms": n, "batch": B, "ms": ms, "configs_per_s": B / (ms * 1e-3),
                "alg_tflops": tf, "fp32_peak_tflops": fp32_peak, "frac_fp32_peak": tf / fp32_peak, "alg_gbs": gbs,
                "frac_hbm_peak": gbs / hbm_peak, "finite": bool(torch.isfinite(lp).all() and torch.isfinite(f).all())}
        print(json.dumps(line), flush=True)
    if args.cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pita_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        Bc = 100000 if n == 13 else 10000
        xc = x[:Bc].cpu()
        O.lj_logprob_force(xc, n)
        t0 = time.perf_counter()
        O.lj_logprob_force(xc, n)
        dt = time.perf_counter() - t0
        print(json.dumps({"bench": "lj_energy_force", "kernel": "cpu-oracle (torch autograd, fp32)", "n_atoms": n, "batch": Bc, "ms": dt * 1e3,
                          "configs_per_s": Bc / dt, "cores": threads}))


if __name__ == "__main__":
    main()
