"""Tensor-level wrappers over the C-ABI entry points (include/pita_b200.h).

Each function takes/returns CUDA fp32 tensors, allocates outputs with torch (the library never
allocates) and launches on torch's current stream.  These are the only call sites of the native
library; the reference-interface classes in this package are built on them.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

import functools
import os

from . import _native as N


NVTX = os.environ.get("PITA_NVTX", "0") not in ("", "0")


def _on_device(fn):
    """Run the native call with the CUDA runtime's current device set to the device of the first tensor argument: the
    launchers use the runtime's current device, torch's current stream is looked up per tensor device."""

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = next((a.device for a in list(args) + list(kwargs.values()) if torch.is_tensor(a) and a.is_cuda), None)
        if dev is None:
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            if not NVTX:
                return fn(*args, **kwargs)
            torch.cuda.nvtx.range_push("pita." + fn.__name__)  # PITA_NVTX=1: one range per native call (ncu --nvtx / nsys)
            try:
                return fn(*args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()

    return wrapped


def _expand(v, B: int, device) -> torch.Tensor:
    """beta / h(t) arrive as python floats, 0-d tensors or [B] tensors in the reference."""
    if not torch.is_tensor(v):
        return torch.full((B,), float(v), device=device, dtype=torch.float32)
    v = v.detach().to(device=device, dtype=torch.float32)
    if v.dim() == 0 or v.numel() == 1:
        return v.reshape(1).expand(B).contiguous()
    if v.numel() != B:
        raise RuntimeError("expected a scalar or %d values, got %d" % (B, v.numel()))
    return v.reshape(B).contiguous()


# ---------------------------------------------------------------------------------------------
@_on_device
def lj_energy_force(x: torch.Tensor, n: int, temperature: float = 1.0, energy_factor: float = 1.0,
                    oscillator_scale: float = 1.0, need_force: bool = True):
    lib = N.load()
    x = N.as_f32(x)
    B = x.shape[0]
    logp = torch.empty(B, device=x.device, dtype=torch.float32)
    force = torch.empty_like(x) if need_force else None
    N.check(lib.pita_lj_energy_force(N.ptr(x), B, n, temperature, energy_factor, oscillator_scale, N.ptr(logp),
                                     N.ptr(force), N.stream_ptr(x.device)), "pita_lj_energy_force")
    return logp, force


def egnn_pack_floats(hidden: int, layers: int) -> int:
    return int(N.load().pita_egnn_pack_floats(hidden, layers))


@_on_device
def egnn_forward(wpack, hidden, layers, n, tcond, y, beta) -> torch.Tensor:
    lib = N.load()
    y = N.as_f32(y)
    B = y.shape[0]
    tcond, beta = _expand(tcond, B, y.device), _expand(beta, B, y.device)
    vel = torch.empty_like(y)
    N.check(lib.pita_egnn_forward(N.ptr(wpack), hidden, layers, n, N.ptr(tcond), N.ptr(y), N.ptr(beta), B, N.ptr(vel),
                                  N.stream_ptr(y.device)), "pita_egnn_forward")
    return vel


@_on_device
def egnn_energy(wpack, hidden, layers, n, ht, x, beta, need_grad=True, need_dh=True):
    lib = N.load()
    x = N.as_f32(x)
    B = x.shape[0]
    ht, beta = _expand(ht, B, x.device), _expand(beta, B, x.device)
    e = torch.empty(B, device=x.device, dtype=torch.float32)
    g = torch.empty_like(x) if need_grad else None
    dh = torch.empty(B, device=x.device, dtype=torch.float32) if need_dh else None
    N.check(lib.pita_egnn_energy(N.ptr(wpack), hidden, layers, n, N.ptr(ht), N.ptr(x), N.ptr(beta), B, N.ptr(e), N.ptr(g),
                                 N.ptr(dh), N.stream_ptr(x.device)), "pita_egnn_energy")
    return e, g, dh


DIV_MODES = {"fp32": 0, "3xtf32": 1, "tf32": 2, "bilinear": 3}
_div_ws = {}


@_on_device
def egnn_energy_laplacian(wpack, hidden, layers, n, ht, x, beta) -> torch.Tensor:
    """tr(Hess_x E) per particle (compute_laplacian_exact of EnergyNet.forward_energy, utils.py:68-77)."""
    lib = N.load()
    x = N.as_f32(x)
    B = x.shape[0]
    ht, beta = _expand(ht, B, x.device), _expand(beta, B, x.device)
    lap = torch.empty(B, device=x.device, dtype=torch.float32)
    N.check(lib.pita_egnn_energy_laplacian(N.ptr(wpack), hidden, layers, n, N.ptr(ht), N.ptr(x), N.ptr(beta), B, N.ptr(lap),
                                           N.stream_ptr(x.device)), "pita_egnn_energy_laplacian")
    return lap


def default_div_mode() -> str:
    """PITA_DIV_MODE=bilinear|3xtf32|tf32|fp32.  Default: bilinear — the round-2 engine (csrc/egnn_tri_*.cu, one dense product
    per (middle-layer edge, tangent node)); 3xtf32 / tf32 are the round-1 forward-mode kernel, fp32 the CUDA-core one."""
    import os
    return os.environ.get("PITA_DIV_MODE", "bilinear").lower()


@_on_device
def egnn_score_div(wpack, hidden, layers, n, ht, x, beta, need_div=True, mode=None):
    lib = N.load()
    x = N.as_f32(x)
    B = x.shape[0]
    ht, beta = _expand(ht, B, x.device), _expand(beta, B, x.device)
    s = torch.empty_like(x)
    d = torch.empty(B, device=x.device, dtype=torch.float32) if need_div else None
    m = DIV_MODES[(mode or default_div_mode())]
    ws, ws_bytes = None, 0
    if need_div and m != 0:
        ws_bytes = int(lib.pita_egnn_score_div_workspace_bytes(n, m))
        key = (str(x.device), torch.cuda.current_stream(x.device).cuda_stream)
        ws = _div_ws.get(key)
        if ws is None or ws.numel() < ws_bytes:
            ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
            _div_ws[key] = ws
    N.check(lib.pita_egnn_score_div(N.ptr(wpack), hidden, layers, n, N.ptr(ht), N.ptr(x), N.ptr(beta), B, N.ptr(s), N.ptr(d), m,
                                    None if ws is None else ws.data_ptr(), ws_bytes, N.stream_ptr(x.device)),
            "pita_egnn_score_div")
    return s, d


# ---------------------------------------------------------------------------------------------
@_on_device
def sde_fk_step(x, grad_u, score, noise, div, dE_dh, energy, n: int, *, g2, gamma, dgamma_dt, dh_dt, dt, sqrt_dt,
                noise_scale, debias=True, freeze_x=False, remove_mean=True, seed=0, offset=0, want_a_raw=True,
                out: Optional[torch.Tensor] = None):
    lib = N.load()
    x = N.as_f32(x)
    B = x.shape[0]
    x_out = torch.empty_like(x) if out is None else out
    a_raw = torch.empty(B, device=x.device, dtype=torch.float32) if want_a_raw else None
    p = N.SdeParams(g2, gamma, dgamma_dt, dh_dt, dt, sqrt_dt, noise_scale, int(debias), int(freeze_x), int(remove_mean),
                    int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1))
    N.check(lib.pita_sde_fk_step(N.ptr(x), N.ptr(grad_u), N.ptr(score), N.ptr(noise), N.ptr(div), N.ptr(dE_dh), N.ptr(energy),
                                 B, n, ctypes.byref(p), N.ptr(x_out), N.ptr(a_raw), N.stream_ptr(x.device)),
            "pita_sde_fk_step")
    return x_out, a_raw


K_QUANTILE_MAX_CHUNK = 8192  # one CTA sorts a chunk in shared memory (csrc/sde.cu::fk_quantile_kernel)


@_on_device
def fk_quantile_accumulate(a_raw, a, chunk: int, q: float, dt: float, zero_a: bool, want_drift=False):
    lib = N.load()
    a_raw = N.as_f32(a_raw)
    B = a_raw.shape[0]
    if min(int(chunk), B) > K_QUANTILE_MAX_CHUNK:
        # inference_batch_size above the in-SMEM sort (e.g. the reference's default batch_size=None: one quantile over the
        # whole shard, sde_integration.py:312-343 + sdes.py:230): the SAME chunk partition, each chunk's quantile by
        # torch.quantile on the device (the reference's own op; <= 16M elements per chunk, torch's limit)
        pieces = []
        for lo in range(0, B, int(chunk)):
            c = a_raw[lo:lo + int(chunk)]
            pieces.append(torch.clamp(c, max=torch.quantile(c, q)))
        drift = torch.cat(pieces)
        if zero_a:
            a_out = torch.zeros_like(a_raw)
        else:
            a_out = drift if a is None else a + drift * dt  # a is None: the clamped values themselves (include/pita_b200.h)
        return a_out, (drift if want_drift else None)
    a_out = torch.empty_like(a_raw)
    drift = torch.empty_like(a_raw) if want_drift else None
    N.check(lib.pita_fk_quantile_accumulate(N.ptr(a_raw), N.ptr(a), B, int(chunk), q, dt, int(zero_a), N.ptr(a_out),
                                            N.ptr(drift), N.stream_ptr(a_raw.device)), "pita_fk_quantile_accumulate")
    return a_out, drift


# ---------------------------------------------------------------------------------------------
_ws_cache = {}


def _workspace(Ntot: int, device) -> torch.Tensor:
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    nbytes = int(N.load().pita_resample_workspace_bytes(Ntot))
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, device=device, dtype=torch.uint8)
        _ws_cache[key] = ws
    return ws


@_on_device
def softmax_clip(logits: torch.Tensor) -> torch.Tensor:
    lib = N.load()
    logits = N.as_f32(logits)
    w = torch.empty_like(logits)
    ws = _workspace(logits.numel(), logits.device)
    N.check(lib.pita_softmax_clip(N.ptr(logits), logits.numel(), N.ptr(w), ws.data_ptr(), N.stream_ptr(logits.device)),
            "pita_softmax_clip")
    return w


@_on_device
def resample_systematic(w: torch.Tensor, u0: float, slot_lo: int = 0, slot_hi: Optional[int] = None,
                        count_changes: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    lib = N.load()
    w = N.as_f32(w)
    Ntot = w.numel()
    slot_hi = Ntot if slot_hi is None else slot_hi
    ids = torch.empty(slot_hi - slot_lo, device=w.device, dtype=torch.int64)
    ch = torch.zeros(1, device=w.device, dtype=torch.int64) if count_changes else None
    ws = _workspace(Ntot, w.device)
    N.check(lib.pita_resample_systematic(N.ptr(w), Ntot, float(u0), slot_lo, slot_hi, N.ptr(ids, torch.int64),
                                         N.ptr(ch, torch.int64), ws.data_ptr(), N.stream_ptr(w.device)),
            "pita_resample_systematic")
    return ids, ch


@_on_device
def gather_rows(src_ptrs: Sequence[int], rows_per_rank: int, ids: torch.Tensor, row_floats: int,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dst[i] = src_{ids[i] // rows_per_rank}[ids[i] % rows_per_rank]; src_ptrs are raw device pointers (peers allowed)."""
    lib = N.load()
    n_out = ids.numel()
    dst = torch.empty(n_out, row_floats, device=ids.device, dtype=torch.float32) if out is None else out
    arr = (ctypes.c_void_p * len(src_ptrs))(*[ctypes.c_void_p(int(p)) for p in src_ptrs])
    N.check(lib.pita_gather_rows(arr, len(src_ptrs), int(rows_per_rank), N.ptr(ids, torch.int64), n_out, int(row_floats),
                                 N.ptr(dst), N.stream_ptr(ids.device)), "pita_gather_rows")
    return dst


@_on_device
def remove_mean(x: torch.Tensor, n: int) -> torch.Tensor:
    lib = N.load()
    xc = N.as_f32(x).reshape(-1, 3 * n)
    out = torch.empty_like(xc)
    N.check(lib.pita_remove_mean(N.ptr(xc), xc.shape[0], n, N.ptr(out), N.stream_ptr(xc.device)), "pita_remove_mean")
    return out.reshape(x.shape)


# ---------------------------------------------------------------------------------------------
@_on_device
def descent_step(x, force, noise, n: int, dt: float, remove_mean_: bool):
    lib = N.load()
    x = N.as_f32(x)
    out = torch.empty_like(x)
    N.check(lib.pita_descent_step(N.ptr(x), N.ptr(N.as_f32(force)), N.ptr(noise), x.shape[0], n, dt, int(remove_mean_),
                                  N.ptr(out), N.stream_ptr(x.device)), "pita_descent_step")
    return out


@_on_device
def mala_propose(x, force, noise, n: int, dt: float):
    lib = N.load()
    x = N.as_f32(x)
    xp = torch.empty_like(x)
    lq = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    N.check(lib.pita_mala_propose(N.ptr(x), N.ptr(N.as_f32(force)), N.ptr(N.as_f32(noise)), x.shape[0], n, dt, N.ptr(xp),
                                  N.ptr(lq), N.stream_ptr(x.device)), "pita_mala_propose")
    return xp, lq


@_on_device
def mala_accept(x, logp, x_prop, logp_prop, force_prop, log_q_fwd, uniform, n: int, dt: float, remove_mean_: bool):
    """In place on x and logp; returns the accepted mask (float 0/1)."""
    lib = N.load()
    acc = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    N.check(lib.pita_mala_accept(N.ptr(x), N.ptr(logp), N.ptr(x_prop), N.ptr(logp_prop), N.ptr(N.as_f32(force_prop)),
                                 N.ptr(log_q_fwd), N.ptr(N.as_f32(uniform)), x.shape[0], n, dt, int(remove_mean_), N.ptr(acc),
                                 N.stream_ptr(x.device)), "pita_mala_accept")
    return acc
