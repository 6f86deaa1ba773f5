#!/usr/bin/env python
"""Benchmark of the annealed Feynman-Kac sampling step (BASELINE.json metric: particle-steps/s).

  python bench.py [--gpus N --steps K --warmup W] [--workload lj55|lj13|aldp22] [--particles P] [--scaling weak|strong]
                  [--total-particles T] [--impl ours|reference] [--materialised-noise]

One "step" = one full debiased FK step over all particles through the integrator's public step
(`WeightedSDEIntegrator.sharded_step`, the sharded form of the reference's `ddp_batched_euler_maruyama_step`,
sde_integration.py:214-351): energy net (U, grad U, dU/dt), score net (score + exact divergence), fused
Euler-Maruyama / FK update with the chunk-quantile clamp, systematic resampling (all-gather of log-weights + scan +
search + peer-memory gather when N > 1).  The SDE time advances from step to step like in the production loop.
Prints ONE JSON line (rank 0):
  value      device-resident throughput (inputs in HBM), CUDA events, max over ranks;
  e2e        the same metric through `integrate_sde` on HOST buffers: every call copies its particles from pinned host
             memory, integrates two steps, and copies particles and log-weights back;
  roofline   the dominant kernel group (score + exact divergence) against the SURVEY §8(d) algorithmic FLOPs, plus the
             executed-instruction view (tensor pipe / issue slots from the round's ncu capture, profiles/r2_ncu_summary.json);
  hbm_kernels   achieved GB/s of the HBM-bound kernels (fused SDE/FK step, softmax+scan+search, gather) at this N;
  cpu_baseline / torch_gpu_baseline   the oracle port of the reference path on the host cores / eagerly on this GPU.
`--impl reference` times the CPU oracle port of the reference path on all host cores (the reference is pure
Python/PyTorch; pinned to it by tests/golden) on a bounded sample of the same workload.
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
    # SURVEY §8 row a8': the alanine-dipeptide denoiser (EGNN_dynamics_AD2_cat, 22 atoms, hidden 64, 5 layers) in the same loop;
    # the molecular target energy (OpenMM) is out of scope and is not on the pin_energy=False path; chunk = inference_batch_size
    # of configs/experiment/aldp.yaml:27
    "aldp22": dict(n=22, particles=1 << 14, chunk=2048, sigma_min=0.05, hidden=64, layers=5,
                   label="ALDP-22 annealed FK sampling (EGNN_dynamics_AD2_cat 64x5), 16k particles/GPU"),
}
CPU_SAMPLE = {13: (512, 64), 22: (96, 24), 55: (16, 8)}            # (particles per step, inference chunk) of the host legs
GPU_ORACLE_SAMPLE = {13: (4096, 512), 22: (512, 64), 55: (64, 16)}  # same for the eager-torch-on-GPU leg
GAMMA = 4.0 / 3.0  # beta_lower / beta for the 4.0 -> 3.0 rung of the temperature ladder (lj13.yaml:44-50)
BETA = 0.75
T_STEP = 0.5       # SDE time of the first timed step (the grid then advances by dt = 1/1000 per step)
S_TOTAL = 1000     # energytemp.yaml:79 — dt of the production loop


def egnn_macs_forward(n, H=32, L=3):
    """MAC per sample of the reference's dense EGNN forward."""
    E = n * (n - 1)
    return L * (E * ((2 * H + 2) * H + H * H + H + H * H + H) + n * (2 * H * H + H * H))


def make_denoiser(wl):
    if wl["n"] == 22:
        from pita_b200.egnn_dynamics_ad2_cat import EGNN_dynamics_AD2_cat
        return EGNN_dynamics_AD2_cat(n_particles=22, n_dimensions=3, hidden_nf=64, n_layers=5, act_fn=torch.nn.SiLU(), recurrent=True,
                                     attention=True, tanh=True, agg="sum", condition_beta=True)
    from pita_b200.egnn_temp_conditioned import EGNN_dynamics
    return EGNN_dynamics(n_particles=wl["n"], n_dimension=3, hidden_nf=32, n_layers=3, act_fn=torch.nn.SiLU(), recurrent=True,
                         tanh=True, attention=True, condition_time=True, condition_temperature=True, agg="sum")


def executed_tensor_macs(n):
    """Dense 32x32 products the bilinear engine actually issues per particle (csrc/egnn_tri_*.cu): phase A 25 per edge
    (3xTF32 = 3 MMAs each), phase B one TF32 product per (edge, k) plus prologue / S / R items."""
    E = n * (n - 1)
    return 1024 * (25 * E * 3 + E * (n + 4 * 2 + 4 * 3))


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
    from pita_b200.energy_net import EnergyNet
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.score_net import ScoreNet
    from pita_b200.sdes import VEReverseSDE
    torch.manual_seed(seed)
    net_s = make_denoiser(wl)
    import copy
    net_e = copy.deepcopy(net_s)  # energytemp_module.py:99
    sched = ElucidatingNoiseSchedule(wl["sigma_min"], 80.0, 7.0)
    sde = VEReverseSDE(sched, energy_net=EnergyNet(net_e.to(device)), score_net=ScoreNet(net_s.to(device)))
    return sde, sched


def oracle_step_seconds(wl, n_particles, steps, threads, device="cpu", chunk=None):
    """Oracle port of the reference path (plain torch, vmap(jacrev) divergence like utils.py:30-51): one debiased FK step
    (fk_drift per inference chunk + update + resample) on a bounded sample; on the host cores or eagerly on a GPU."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pita_oracle as O
    torch.set_num_threads(threads)
    n = wl["n"]
    chunk = chunk or n_particles
    state = (O.random_egnn_state(5, 64, seed=12345, dtype=torch.float32, in_nf=23) if n == 22
             else O.random_egnn_state(seed=12345, dtype=torch.float32))
    sd = {k: v.to(device) for k, v in state.items()}
    sched = O.EDMSchedule(wl["sigma_min"])
    gen = torch.Generator(device="cpu").manual_seed(0)
    scale = float((sched.h(torch.tensor(T_STEP)) / GAMMA) ** 0.5)
    x = O.mean_free_prior(n_particles, n, scale, gen=gen).to(device)
    g = float(sched.g(torch.tensor(T_STEP)))
    times = []
    for _ in range(steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        dx, da = [], []
        with torch.device(device):  # the oracle creates its per-chunk constants on the default device
            for lo in range(0, n_particles, chunk):  # the reference's inference_batch_size loop (sde_integration.py:312-343)
                d = O.fk_drift(sd, sd, sched, O.ConstGamma(GAMMA), T_STEP, x[lo:lo + chunk], BETA, n)
                dx.append(d.drift_x)
                da.append(d.drift_a)
            drift_x, drift_a = torch.cat(dx), torch.cat(da)
            xn = x + drift_x * 1e-3 + g * torch.randn_like(x) * np.sqrt(1e-3)
        ids = O.systematic_resample((drift_a * 1e-3).cpu(), 0.5)
        x = O.centre(xn[torch.from_numpy(ids).to(device)], n).detach()
        if device != "cpu":
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args, wl):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port (pure-Python reference,
    pinned to it by tests/golden), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample, chunk = CPU_SAMPLE[wl["n"]]  # larger LJ-55 chunks exhaust host memory; chunks run back to back
    sample = args.cpu_particles or sample
    ts = oracle_step_seconds(wl, sample, args.warmup + args.steps, threads, chunk=chunk)[args.warmup:]
    total = sum(ts)
    val = sample * len(ts) / total
    line = {"impl": "reference", "metric": "particle_steps_per_s", "value": val, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "sample": "%d particles/step on the host, inference chunks of %d" % (sample, chunk)},
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                             "sample": "%d particles x %d steps in chunks of %d, torch CPU oracle (vmap(jacrev) divergence)" % (sample, len(ts), chunk)},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def timed(fn, reps, flush):
    """Median CUDA-event time (ms) of fn() on the current stream, L2 flushed before every launch."""
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def exchange_check(integ, dev, world, rank, Nl, D):
    """Driver-visible multi-GPU correctness: rows that encode their own global index go through the sharded resampler
    (all-gather of a, global scan, peer-memory gather); every rank's result must be the rows its own slots' ancestors
    name, the ancestors computed independently by a full single-rank pita_resample_systematic over all N weights
    (bit-exact against the reference's sample_cat_sys in tests/test_gpu_resample.py)."""
    import torch.distributed as dist
    from pita_b200 import ops
    rs = integ._resampler
    N = Nl * world
    gen = torch.Generator(device=dev).manual_seed(99)  # same stream on every rank
    a_full_ref = torch.randn(N, device=dev, generator=gen) * 3.0
    u0 = 0.3141592653589793
    lo = rank * Nl
    x_local = (torch.arange(lo, lo + Nl, device=dev, dtype=torch.float32)[:, None] + torch.zeros(1, D, device=dev)).contiguous()
    a_g = rs.gather_logweights(a_full_ref[lo:lo + Nl].contiguous())
    ok = bool(torch.equal(a_g, a_full_ref))
    buf = rs.particle_buffer()
    if buf is not None:
        buf.copy_(x_local)
        x_local = buf
    x_new, changes = rs.resample(x_local, a_g, u0)
    ids_all, ch = ops.resample_systematic(ops.softmax_clip(a_full_ref), u0, 0, N, count_changes=True)
    ok = ok and bool(torch.equal(x_new[:, 0].long(), ids_all[lo:lo + Nl])) and bool(torch.equal(x_new[:, 0], x_new[:, D - 1]))
    ok = ok and int(changes.item()) == int(ch.item())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-particles", type=int, default=1 << 20, help="--scaling strong: particles over ALL GPUs")
    ap.add_argument("--cpu-particles", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--materialised-noise", action="store_true", help="torch.randn noise tensor instead of in-kernel Philox")
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
    if args.scaling == "strong":
        wl["particles"] = args.total_particles // world
        wl["label"] = "LJ-%d annealed FK sampling, %d particles in total (strong scaling)" % (wl["n"], args.total_particles)
    n, D, Nl = wl["n"], 3 * wl["n"], wl["particles"]
    N = Nl * world
    sde, sched = build_problem(wl, dev)
    K, W = args.steps, args.warmup
    fused = not args.materialised_noise
    integ = WeightedSDEIntegrator(sde=sde, num_integration_steps=S_TOTAL, start_resampling_step=0, end_resampling_step=S_TOTAL,
                                  lightning_module=None, resampling_interval=1, num_negative_time_steps=0, post_mcmc_steps=0,
                                  batch_size=wl["chunk"], fused_noise=fused, collect_logweights=False)
    gam = ConstantAnnealingFactorSchedule(GAMMA)
    if n in (13, 55):
        tgt = LennardJonesEnergy(dimensionality=D, n_particles=n)
    else:  # the molecular target (OpenMM) is out of scope; the pin_energy=False path reads only its shape
        import types
        tgt = types.SimpleNamespace(n_particles=n, n_spatial_dim=3, is_molecule=True)
    Hn, Ln = wl.get("hidden", 32), wl.get("layers", 3)
    scale = float((sched.h(torch.tensor(T_STEP, dtype=torch.float64)) / GAMMA) ** 0.5)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = ops.remove_mean(torch.randn(Nl, D, device=dev, generator=gen) * scale, n)
    a = torch.zeros(Nl, device=dev)
    torch.manual_seed(777)               # same CPU stream on every rank (Philox key of the call, u0)
    torch.cuda.manual_seed(4321 + rank)  # per-rank diffusion noise when it is materialised
    integ.prepare(Nl, D, dev)
    dt = 1.0 / S_TOTAL
    sqrt_dt = float(torch.tensor(dt).sqrt())
    times = torch.linspace(1.0, 0.0, S_TOTAL + 1)[:-1]
    step0 = int(round((1.0 - T_STEP) * S_TOTAL))

    def fk_step(xx, aa, k):
        return integ.sharded_step(float(times[step0 + k]), step0 + k, xx, aa, dt, sqrt_dt, BETA, n, gam, tgt, 1)[:2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    exchange_parity = None
    if world > 1:
        exchange_parity = exchange_check(integ, dev, world, rank, Nl, D)

    # ---- device-resident timing: inputs live in HBM; working set (x, grads, scores) > L2 for the default sizes
    for k in range(W):
        x, a = fk_step(x, a, k)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
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

    # ---- dominant kernel group (score + exact divergence) timed alone on the launching stream, L2 flushed between launches
    ht = torch.full((Nl,), float(sched.h(torch.tensor(T_STEP, dtype=torch.float64))), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    k_ms = timed(lambda: sde.score_net.score_and_divergence(ht, x, BETA), 3 if n == 13 else 2, flush)
    en_ms = timed(lambda: sde.energy_net._terms(ht, x, BETA, True, True), 2, flush)
    alg_flops = (3 * n + 1) * 2.0 * egnn_macs_forward(n, Hn, Ln) * Nl  # (3n+1) forward-equivalents per particle
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved_tf = alg_flops / (k_ms * 1e-3) / 1e12
    ncu = {}
    try:  # written from the round's last `ncu --set full` capture by profiles/ncu_to_json.py
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json"))).get("n%d" % n, {})
    except Exception:  # noqa: BLE001
        pass
    traffic = ncu.get("dram_bytes_per_particle")
    roofline = {"kernel": "score + exact divergence: tri_phase_a_kernel + tri_phase_b_kernel (%s)" % ops.default_div_mode(),
                "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": None if traffic is None else traffic * Nl, "kernel_ms": k_ms,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s (of fallback)",
                "executed_tflops": 2.0 * executed_tensor_macs(n) * Nl / (k_ms * 1e-3) / 1e12,
                "ncu": ncu or None,
                "note": "achieved = SURVEY 8d algorithmic FLOPs ((3n+1) dense EGNN forwards per particle) / time: the convention the "
                        "contract asks for.  The bilinear engine EXECUTES ~%.0fx fewer tensor FLOPs than that (executed_tflops), as "
                        "tcgen05 kind::tf32 MMAs; in practice it is bound by the CUDA-core issue rate of the element-wise work around "
                        "every product (ncu: issue-slot and tensor-pipe utilisation), not by the tensor pipe"
                        % (alg_flops / (2.0 * executed_tensor_macs(n) * Nl))}

    if n == 22:
        roofline.update({"kernel": "score + exact divergence: ad2_score_div_kernel (fp32 FFMA on the CUDA cores)", "executed_tflops": None,
                         "note": "achieved = SURVEY 8d algorithmic FLOPs ((3n+1) dense EGNN forwards per particle) / time, against the "
                                 "tensor peak this GEMM-shaped work should reach; this round the 64x5 network runs as fp32 FFMA matvecs "
                                 "(one CTA per particle, structured tangents: layer 0 touches only the edges of the perturbed atom)"})

    # ---- HBM-bound kernels of the step at this N (algorithmic bytes of SURVEY §8d / achieved GB/s / fraction of the measured peak)
    gu, sc_ = torch.randn_like(x), torch.randn_like(x)
    dv, dh, en = (torch.randn(Nl, device=dev) for _ in range(3))
    sp = dict(g2=2.0, gamma=GAMMA, dgamma_dt=0.0, dh_dt=1.0, dt=dt, sqrt_dt=sqrt_dt, noise_scale=1.4)
    t_sde = timed(lambda: ops.sde_fk_step(x, gu, sc_, None, dv, dh, en, n, seed=1, offset=3, **sp), 5, flush)
    araw = torch.randn(Nl, device=dev)
    t_q = timed(lambda: ops.fk_quantile_accumulate(araw, a, wl["chunk"], 0.9, dt, False), 5, flush)
    wts = ops.softmax_clip(araw)
    t_rs = timed(lambda: ops.resample_systematic(ops.softmax_clip(araw), 0.37), 5, flush)
    ids, _ = ops.resample_systematic(wts, 0.37)
    t_g = timed(lambda: ops.gather_rows([x.data_ptr()], Nl, ids, D), 5, flush)

    def hb(bytes_per_particle, t):
        gbs = bytes_per_particle * Nl / (t * 1e-3) / 1e9
        return {"ms": t, "bytes_per_particle": bytes_per_particle, "gbs": gbs, "frac_hbm_peak": gbs / hbm_peak}

    hbm_kernels = {"peak_gbs": hbm_peak, "sde_fk_step_kernel (in-kernel Philox)": hb(12 * D + 12 + 12, t_sde),
                   "fk_quantile_kernel": hb(12, t_q), "softmax + scan + search": hb(4 + 4 + 4 + 8, t_rs),
                   "gather_rows_kernel": hb(8 + 8 * D, t_g)}

    # ---- second half of BASELINE.json's metric: the Lennard-Jones energy+force kernel as a fraction of the FP32 FMA peak
    #      (31 FLOP per unordered pair + 15 per atom), timed alone on this rank's particles, L2 flushed
    fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    lj_kernel = None
    if n in (13, 55):
        lj_ms = timed(lambda: ops.lj_energy_force(x, n), 5, flush)
        lj_tf = (31 * n * (n - 1) // 2 + 15 * n) * Nl / (lj_ms * 1e-3) / 1e12
        lj_kernel = {"kernel": "lj_pairs_kernel", "configs_per_s": Nl / (lj_ms * 1e-3), "ms": lj_ms, "alg_tflops": lj_tf,
                     "fp32_peak_tflops": fp32_peak, "fp32_peak_source": "derived: 148 SMs x 128 lanes x 2 FLOP x sm_max_mhz (not in MEASURED_PEAKS.json)",
                     "frac_fp32_peak": lj_tf / fp32_peak, "alg_gbs": (8 * D + 4) * Nl / (lj_ms * 1e-3) / 1e9}

    # ---- exchange step alone (N > 1): all-gather of a + global softmax/scan/search + peer gather + barriers + all-reduce
    exchange_ms = None
    if world > 1:
        def one_exchange():
            a_g = integ._resampler.gather_logweights(araw)
            buf = integ._resampler.particle_buffer()
            if buf is not None:
                buf.copy_(x)
            integ._resampler.resample(buf if buf is not None else x, a_g, 0.37)
        one_exchange()
        barrier()
        tx = torch.tensor([timed(one_exchange, 5, flush)], device=dev)
        dist.all_reduce(tx, op=dist.ReduceOp.MAX)
        exchange_ms = float(tx.item())

    # ---- end to end through integrate_sde with HOST buffers: every call copies its particles from pinned host memory,
    #      integrates S_E2E steps (the time grid advances) and copies particles + log-weights back
    S_E2E = 2
    integ_e = WeightedSDEIntegrator(sde=sde, num_integration_steps=S_E2E, start_resampling_step=0, end_resampling_step=S_E2E,
                                    lightning_module=None, resampling_interval=1, num_negative_time_steps=0, post_mcmc_steps=0,
                                    batch_size=wl["chunk"], fused_noise=fused, time_range=S_E2E * dt)
    hx = torch.empty(N, D, pin_memory=True).copy_(torch.randn(N, D) * scale)
    hx_out = torch.empty(N, D).pin_memory()
    hlw = torch.empty(S_E2E, N).pin_memory()
    # one untimed call first: it creates the integrator's resampler (with several ranks: the symmetric-memory rendezvous)
    integ_e.integrate_sde(hx.to(dev, non_blocking=True), tgt, gam, inverse_temperature=BETA)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    KE = max(1, min(K, 2))
    e0.record()
    for k in range(KE):
        dx = hx.to(dev, non_blocking=True)  # the reference's contract: the full [N, D] set goes in, each rank takes its slice
        xo, lw, uniq, _, _ = integ_e.integrate_sde(dx, tgt, gam, inverse_temperature=BETA)
        hx_out.copy_(xo, non_blocking=True)
        hlw.copy_(lw, non_blocking=True)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = N * KE * S_E2E / (float(e2e_ms.item()) * 1e-3)
    # our kernels per step: energy, phase A x batches, phase B x batches, finalize x batches, sde_fk_step, fk_quantile, resampling
    # (softmax partials / finalize / clip, scan tile sums / offsets / bins, search, change count, gather) = 9
    batches = -(-Nl // (148 * 2 * (128 // n)))
    launches_per_step = (1 + 1 + 2 + 9) if n == 22 else (1 + 3 * batches + 2 + 9)

    cpu_baseline = gpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample, chunk = CPU_SAMPLE[n]
        sample = args.cpu_particles or sample
        ts = oracle_step_seconds(wl, sample, 2, threads, chunk=chunk)[1:]
        cpu_baseline = {"value": sample * len(ts) / sum(ts), "unit": "particle-steps/s", "cores": threads, "kind": "port",
                        "sample": "%d particles x %d step (after 1 warm-up) in chunks of %d, torch CPU oracle of the reference path" % (sample, len(ts), chunk)}
    if rank == 0 and not args.no_gpu_baseline:
        try:
            sample, chunk = GPU_ORACLE_SAMPLE[n]
            ts = oracle_step_seconds(wl, sample, 3, os.cpu_count() or 1, device=str(dev), chunk=chunk)[1:]
            gpu_baseline = {"value": sample * len(ts) / sum(ts), "unit": "particle-steps/s", "kind": "oracle port of the reference path, "
                            "plain torch eager on this B200 (vmap(jacrev) divergence, dense [B,n,n] pair tensors)",
                            "sample": "%d particles x %d steps (after 1 warm-up) in inference chunks of %d" % (sample, len(ts), chunk)}
        except Exception as exc:  # noqa: BLE001 — a baseline leg must not take the benchmark down
            gpu_baseline = {"unavailable": repr(exc)[:200]}

    if rank == 0:
        line = {"metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["label"], "particles_per_gpu": Nl, "n_atoms": n, "debias_inference": True,
                           "resampling_interval": 1, "chunk": wl["chunk"], "egnn": "hidden %d, %d layers, random init seed 12345" % (Hn, Ln),
                           "divergence": "fp32" if n == 22 else ops.default_div_mode(),
                           "noise": "in-kernel philox" if fused else "materialised torch.randn",
                           "l2": "inputs (%.0f MB/step working set) larger than L2" % (Nl * D * 4 * 4 / 1e6),
                           "exchange": integ._resampler.exchange},
                "clocks": clk,
                "e2e": {"value": e2e_val, "unit": "particle-steps/s", "h2d_bytes_per_step": N * D * 4 // S_E2E,
                        "d2h_bytes_per_step": (N * D * 4 + S_E2E * N * 4) // S_E2E,
                        "how": "integrate_sde(S=%d) per call: pinned host [N,D] -> device, %d steps, particles + log-weights -> pinned host" % (S_E2E, S_E2E)},
                "gpu_launches": launches_per_step * K, "roofline": roofline, "energy_kernel_ms": en_ms, "hbm_kernels": hbm_kernels,
                "lj_kernel": lj_kernel, "exchange_parity": exchange_parity, "exchange_ms": exchange_ms,
                "cpu_baseline": cpu_baseline, "torch_gpu_baseline": gpu_baseline}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
