"""ctypes binding of the C-ABI shared library (include/pita_b200.h).

The library is built in-tree by `pita_b200._native.build()` (plain nvcc, sm_100a only) and loaded with
ctypes; PyTorch only supplies device memory and the current CUDA stream.  There is NO fallback: if the
library is missing or a tensor is not a contiguous fp32 CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libpita_b200.so")
SOURCES = ["capi.cu", "lj.cu", "resample.cu", "sde.cu", "egnn.cu", "egnn_rows.cu", "egnn_tri_a.cu", "egnn_tri_b.cu", "egnn_ad2.cu", "egnn_lap.cu",
           "umma_selftest.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _sources():
    return [os.path.join(_CSRC, s) for s in SOURCES if os.path.exists(os.path.join(_CSRC, s))]


def _headers():
    hs = [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith(".cuh")]
    return hs + [os.path.join(_HERE, "..", "include", "pita_b200.h")]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _sources() + _headers() if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into pita_b200/libpita_b200.so (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    build_dir = os.path.join(_HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    for src in _sources():
        obj = os.path.join(build_dir, os.path.basename(src) + ".o")
        objs.append(obj)
        if (not force) and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode()))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout.decode()))
    return LIB_PATH


class _SdeParams(ctypes.Structure):
    _fields_ = [("g2", ctypes.c_float), ("gamma", ctypes.c_float), ("dgamma_dt", ctypes.c_float),
                ("dh_dt", ctypes.c_float), ("dt", ctypes.c_float), ("sqrt_dt", ctypes.c_float),
                ("noise_scale", ctypes.c_float), ("debias", ctypes.c_int), ("freeze_x", ctypes.c_int),
                ("remove_mean", ctypes.c_int), ("seed", ctypes.c_uint64), ("offset", ctypes.c_uint64)]


SdeParams = _SdeParams
_lib: Optional[ctypes.CDLL] = None

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I = ctypes.c_int
_F = ctypes.c_float
_D = ctypes.c_double

# symbol -> (restype, argtypes); every symbol declared in include/pita_b200.h
SIGNATURES = {
    "pita_abi_version": (_I, []),
    "pita_last_error": (ctypes.c_char_p, []),
    "pita_lj_energy_force": (_I, [_P, _I64, _I, _F, _F, _F, _P, _P, _P]),
    "pita_egnn_pack_floats": (_I64, [_I, _I]),
    "pita_egnn_forward": (_I, [_P, _I, _I, _I, _P, _P, _P, _I64, _P, _P]),
    "pita_egnn_energy": (_I, [_P, _I, _I, _I, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "pita_egnn_energy_laplacian": (_I, [_P, _I, _I, _I, _P, _P, _P, _I64, _P, _P]),
    "pita_egnn_score_div_workspace_bytes": (_I64, [_I, _I]),
    "pita_egnn_tri_workspace_layout": (_I64, [_I, ctypes.POINTER(_I64), _I]),
    "pita_egnn_score_div": (_I, [_P, _I, _I, _I, _P, _P, _P, _I64, _P, _P, _I, _P, _I64, _P]),
    "pita_sde_fk_step": (_I, [_P, _P, _P, _P, _P, _P, _P, _I64, _I, ctypes.POINTER(_SdeParams), _P, _P, _P]),
    "pita_fk_quantile_accumulate": (_I, [_P, _P, _I64, _I, _F, _F, _I, _P, _P, _P]),
    "pita_resample_workspace_bytes": (_I64, [_I64]),
    "pita_softmax_clip": (_I, [_P, _I64, _P, _P, _P]),
    "pita_resample_systematic": (_I, [_P, _I64, _D, _I64, _I64, _P, _P, _P, _P]),
    "pita_gather_rows": (_I, [ctypes.POINTER(_P), _I, _I64, _P, _I64, _I, _P, _P]),
    "pita_remove_mean": (_I, [_P, _I64, _I, _P, _P]),
    "pita_descent_step": (_I, [_P, _P, _P, _I64, _I, _F, _I, _P, _P]),
    "pita_mala_propose": (_I, [_P, _P, _P, _I64, _I, _F, _P, _P, _P]),
    "pita_mala_accept": (_I, [_P, _P, _P, _P, _P, _P, _P, _I64, _I, _F, _I, _P, _P]),
    "pita_umma_selftest": (_I, [_P, _P, _P, _I, _P]),
}


def load() -> ctypes.CDLL:
    """Load the C-ABI library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "pita_b200: %s not found — run `python -c 'import __graft_entry__ as g; g.build()'` (there is no "
            "CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pita_abi_version() != 1:
        raise RuntimeError("pita_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pita_last_error().decode(errors="replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))


def ptr(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("pita_b200 kernels need CUDA tensors (got device %s); there is no CPU path" % t.device)
    if t.dtype != dtype:
        raise RuntimeError("expected dtype %s, got %s" % (dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("expected a contiguous tensor")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def as_f32(t: torch.Tensor) -> torch.Tensor:
    """Detached contiguous fp32 view/copy on the same (CUDA) device."""
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
