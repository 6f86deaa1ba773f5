#!/usr/bin/env python
"""Benchmark of the annealed Feynman-Kac sampling step (BASELINE.json metric: particle-steps/s).

  python bench.py [--gpus N --steps K --warmup W] [--workload lj13|lj55] [--particles P] [--impl ours|reference]

One "step" = one full debiased FK step over all particles: energy net (U, grad U, dU/dt), score net
(score + exact divergence), fused Euler-Maruyama/FK update with chunk-quantile clamp, and systematic
resampling (all-gather of log-weights + scan + search + peer-memory gather when N > 1).
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference path on the
host cores (the reference is pure Python/PyTorch) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: LJ-13 annealed sampling with FK resampling, 1M particles on 1xB200
    "lj13": dict(n=13, particles=1 << 20, chunk=512, sigma_min=0.05, label="LJ-13 annealed FK sampling, 1M particles/GPU (BASELINE configs[1])"),
    # BASELINE.json configs[2]: LJ-55, 256k-4M particles sharded across 1/2/4/8 B200
    "lj55": dict(n=55, particles=1 << 18, chunk=512, sigma_min=0.05, label="LJ-55 annealed FK sampling, 256k particles/GPU (BASELINE configs[2])"),
}
# measured by ncu on this kernel (bytes per particle per launch; see profiles/README.md); filled from the round's last capture
NCU_DRAM_BYTES_PER_PARTICLE = {13: (50.237704e9 + 5.881976e9) / 37888,   # profiles/r1n_ncu_scorediv13.txt (37 888 particles)
                               55: (496.511952e9 + 13.595801e9) / 4736}   # profiles/r1n_ncu_scorediv55.txt (4 736 particles)
GAMMA = 4.0 / 3.0  # beta_lower / beta for the 4.0 -> 3.0 rung of the temperature ladder (lj13.yaml:44-50)
BETA = 0.75
T_STEP = 0.5       # SDE time at which the timed steps are evaluated


def egnn_macs_forward(n, H=32, L=3):
    """MAC per sample of the reference's dense EGNN forward."""
    E = n * (n - 1)
    return L * (E * ((2 * H + 2) * H + H * H + H + H * H + H) + n * (2 * H * H + H * H))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_problem(wl, device, seed=12345):
    """Random-init nets with the reference constructors' init (seed 12345, lj13.yaml:11), prior start."""
    from pita_b200.egnn_temp_conditioned import EGNN_dynamics
    from pita_b200.energy_net import EnergyNet
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.score_net import ScoreNet
    from pita_b200.sdes import VEReverseSDE
    torch.manual_seed(seed)
    mk = lambda: EGNN_dynamics(n_particles=wl["n"], n_dimension=3, hidden_nf=32, n_layers=3, act_fn=torch.nn.SiLU(),  # noqa: E731
                               recurrent=True, tanh=True, attention=True, condition_time=True, condition_temperature=True, agg="sum")
    net_s = mk()
    import copy
    net_e = copy.deepcopy(net_s)  # energytemp_module.py:99
    sched = ElucidatingNoiseSchedule(wl["sigma_min"], 80.0, 7.0)
    sde = VEReverseSDE(sched, energy_net=EnergyNet(net_e.to(device)), score_net=ScoreNet(net_s.to(device)))
    return sde, sched


def oracle_step_seconds(wl, n_particles, steps, threads):
    """CPU oracle port of the reference path: one debiased FK step (fk_drift + update + resample) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pita_oracle as O
    torch.set_num_threads(threads)
    n = wl["n"]
    sd = O.random_egnn_state(seed=12345, dtype=torch.float32)
    sched = O.EDMSchedule(wl["sigma_min"])
    gen = torch.Generator().manual_seed(0)
    scale = float((sched.h(torch.tensor(T_STEP)) / GAMMA) ** 0.5)
    x = O.mean_free_prior(n_particles, n, scale, gen=gen)
    cfg = O.LoopConfig(n=n, steps=1, chunk=min(wl["chunk"], n_particles), beta=BETA, resampling_interval=1)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        d = O.fk_drift(sd, sd, sched, O.ConstGamma(GAMMA), T_STEP, x, BETA, n)
        g = float(sched.g(torch.tensor(T_STEP)))
        xn = x + d.drift_x * 1e-3 + g * torch.randn_like(x) * np.sqrt(1e-3)
        ids = O.systematic_resample(d.drift_a * 1e-3, 0.5)
        x = O.centre(xn[torch.from_numpy(ids)], n).detach()
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args, wl):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port (pure-Python reference,
    pinned to it by tests/golden), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.cpu_particles or (64 if wl["n"] == 13 else 4)
    ts = oracle_step_seconds(wl, sample, args.warmup + args.steps, threads)[args.warmup:]
    total = sum(ts)
    val = sample * len(ts) / total
    line = {"impl": "reference", "metric": "particle_steps_per_s", "value": val, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "sample": "%d particles/step on the host" % sample},
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                             "sample": "%d particles x %d steps, torch CPU oracle (vmap(jacrev) divergence)" % (sample, len(ts))},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # BASELINE.json's metric is quoted on LJ-55 (configs[2], 256k particles per GPU at the low end of its 256k-4M range, which
    # fits one GPU); --workload lj13 runs configs[1] (LJ-13, 2^20 particles)
    ap.add_argument("--workload", default=os.environ.get("PITA_BENCH_WORKLOAD", "lj55"), choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the workload's)")
    ap.add_argument("--cpu-particles", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused-noise", action="store_true", help="in-kernel Philox noise instead of a materialised randn tensor")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.particles:
        wl["particles"] = args.particles
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch.distributed as dist
    from pita_b200 import ops
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    from pita_b200.sde_integration import WeightedSDEIntegrator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, D, Nl = wl["n"], 3 * wl["n"], wl["particles"]
    N = Nl * world
    sde, sched = build_problem(wl, dev)
    K, W = args.steps, args.warmup
    S_total = 1000  # energytemp.yaml:79 — dt of the production loop
    integ = WeightedSDEIntegrator(sde=sde, num_integration_steps=S_total, start_resampling_step=0, end_resampling_step=S_total,
                                  lightning_module=None, resampling_interval=1, num_negative_time_steps=0, post_mcmc_steps=0,
                                  batch_size=wl["chunk"], fused_noise=args.fused_noise, collect_logweights=False)
    gam = ConstantAnnealingFactorSchedule(GAMMA)
    tgt = LennardJonesEnergy(dimensionality=D, n_particles=n)
    scale = float((sched.h(torch.tensor(T_STEP, dtype=torch.float64)) / GAMMA) ** 0.5)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = ops.remove_mean(torch.randn(Nl, D, device=dev, generator=gen) * scale, n)
    a = torch.zeros(Nl, device=dev)
    torch.cuda.manual_seed(4321 + rank)  # per-rank diffusion noise
    integ.prepare(Nl, D, dev)
    dt, sqrt_dt = 1.0 / S_total, float(torch.tensor(1.0 / S_total).sqrt())
    step0 = int(round((1.0 - T_STEP) * S_total))

    def fk_step(xx, aa, k):
        return integ._fk_step(T_STEP, step0 + k, xx, aa, dt, sqrt_dt, BETA, n, gam, tgt, 1)[:2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: inputs live in HBM; working set (x, grads, scores, noise) > L2 for the default sizes
    for k in range(W):
        x, a = fk_step(x, a, k)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kern_ev = []
    ev[0].record()
    for k in range(K):
        x, a = fk_step(x, a, W + k)
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    clk = clocks.stop()
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = N * K / (ms * 1e-3)

    # ---- dominant kernel (score + exact divergence) timed alone on the launching stream, L2 flushed between launches
    ht = torch.full((Nl,), float(sched.h(torch.tensor(T_STEP, dtype=torch.float64))), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kt = []
    for _ in range(3 if n == 13 else 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sde.score_net.score_and_divergence(ht, x, BETA)
        e1.record()
        torch.cuda.synchronize()
        kt.append(e0.elapsed_time(e1))
    k_ms = sorted(kt)[len(kt) // 2] if len(kt) > 2 else min(kt)
    alg_flops = (3 * n + 1) * 2.0 * egnn_macs_forward(n) * Nl  # (3n+1) forward-equivalents per particle
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved_tf = alg_flops / (k_ms * 1e-3) / 1e12
    # DRAM traffic of this kernel from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum), captured on a smaller
    # launch of the same kernel and scaled per particle (profiles/README.md names the capture); None if never captured
    traffic = NCU_DRAM_BYTES_PER_PARTICLE.get(n)
    roofline = {"kernel": "egnn_score_div_rows_kernel (%s)" % ops.default_div_mode(), "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": None if traffic is None else traffic * Nl, "kernel_ms": k_ms,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s (of fallback)",
                "note": "algorithmic FLOPs = (3n+1) dense EGNN forwards per particle (SURVEY 8d); the kernel executes ~21 (n=13) / ~76 (n=55) "
                        "forward-equivalents (structured tangents) as tcgen05 kind::tf32 MMAs (x3 in 3xtf32 mode) against the bf16 peak"}

    # ---- second half of BASELINE.json's metric: the Lennard-Jones energy+force kernel as a fraction of the FP32 FMA peak
    #      (31 FLOP per unordered pair + 15 per atom), timed alone on this rank's particles, L2 flushed
    lj_t = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.lj_energy_force(x, n)
        e1.record()
        torch.cuda.synchronize()
        lj_t.append(e0.elapsed_time(e1))
    lj_ms = sorted(lj_t)[2]
    fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    lj_tf = (31 * n * (n - 1) // 2 + 15 * n) * Nl / (lj_ms * 1e-3) / 1e12
    lj_kernel = {"kernel": "lj_pairs_kernel", "configs_per_s": Nl / (lj_ms * 1e-3), "ms": lj_ms, "alg_tflops": lj_tf,
                 "fp32_peak_tflops": fp32_peak, "frac_fp32_peak": lj_tf / fp32_peak,
                 "alg_gbs": (8 * D + 4) * Nl / (lj_ms * 1e-3) / 1e9}

    # ---- end to end through the public step with HOST buffers: H2D of (x, a), one FK step, D2H of (x', a')
    hx = torch.empty(Nl, D, pin_memory=True).copy_(x.cpu())
    ha = torch.zeros(Nl, pin_memory=True)
    hx_out, ha_out = torch.empty_like(hx).pin_memory(), torch.empty_like(ha).pin_memory()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    KE = max(1, min(K, 3))
    e0.record()
    for k in range(KE):
        dx = hx.to(dev, non_blocking=True)
        da = ha.to(dev, non_blocking=True)
        dx, da = fk_step(dx, da, W + K + k)
        hx_out.copy_(dx, non_blocking=True)
        ha_out.copy_(da, non_blocking=True)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = N * KE / (float(e2e_ms.item()) * 1e-3)
    # our kernels per step (profiles/r1d_launches_lj13.csv): energy, score_div, sde_fk_step, fk_quantile + resampling (softmax
    # partials / finalize / clip, scan tile sums / offsets / bins, search, change count, gather) = 13; torch's randn / fills not counted
    launches_per_step = 4 + 9

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = args.cpu_particles or (64 if n == 13 else 4)
        ts = oracle_step_seconds(wl, sample, 2, threads)[1:]
        cpu_baseline = {"value": sample * len(ts) / sum(ts), "unit": "particle-steps/s", "cores": threads, "kind": "port",
                        "sample": "%d particles x %d step (after 1 warm-up), torch CPU oracle of the reference path" % (sample, len(ts))}

    if rank == 0:
        line = {"metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["label"], "particles_per_gpu": Nl, "n_atoms": n, "debias_inference": True,
                           "resampling_interval": 1, "chunk": wl["chunk"], "egnn": "hidden 32, 3 layers, random init seed 12345",
                           "noise": "in-kernel philox" if args.fused_noise else "materialised torch.randn",
                           "l2": "inputs (%.0f MB/step working set) larger than L2" % (Nl * D * 4 * 5 / 1e6),
                           "exchange": integ._resampler.exchange},
                "clocks": clk, "e2e": {"value": e2e_val, "unit": "particle-steps/s", "h2d_bytes_per_step": Nl * D * 4 + Nl * 4,
                                       "d2h_bytes_per_step": Nl * D * 4 + Nl * 4},
                "gpu_launches": launches_per_step * K, "roofline": roofline, "lj_kernel": lj_kernel, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
