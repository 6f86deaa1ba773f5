"""The C-ABI library builds, loads, and exports every symbol include/pita_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from pita_b200 import _native
    return _native.build()


def test_header_symbols_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "pita_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pita_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 19
    lib = ctypes.CDLL(lib_path)
    for nm in sorted(names):
        assert hasattr(lib, nm), "missing export: " + nm


def test_binding_table_covers_header(lib_path):
    from pita_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "pita_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pita_[a-z0-9_]+)\s*\(", hdr))
    assert names == set(_native.SIGNATURES), names ^ set(_native.SIGNATURES)
    lib = _native.load()
    assert lib.pita_abi_version() == 1
    assert lib.pita_egnn_pack_floats(32, 3) == 96 + 3 * 14656
    assert lib.pita_egnn_pack_floats(64, 3) == -1
    assert lib.pita_resample_workspace_bytes(1 << 20) > 4 * (1 << 20)


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing on the host."""
    import torch
    from pita_b200 import ops
    with pytest.raises(RuntimeError):
        ops.lj_energy_force(torch.zeros(4, 39), 13)
    with pytest.raises(RuntimeError):
        ops.softmax_clip(torch.zeros(16))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pita_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            for line in src.splitlines():
                if re.match(r"\s*(import|from)\s", line):
                    assert not re.search(r"oracle|egnn_analytic", line), (fn, line)


def test_argument_validation_without_a_gpu(lib_path):
    """Argument checks return error codes before anything is launched (so they can run here): empty batches are a no-op,
    unsupported network shapes / atom counts and null pointers are refused with a message, for every EGNN entry point
    including the alanine-dipeptide dispatch and the Laplacian."""
    from pita_b200 import _native
    lib = _native.load()
    null = None
    fake = ctypes.c_void_p(0x1000)  # never dereferenced: validation fails (or B == 0 returns) first
    assert lib.pita_egnn_pack_floats(64, 5) == 23 * 64 + 64 + 5 * (14 * 4096 + 10 * 64)
    assert lib.pita_egnn_score_div_workspace_bytes(22, 3) == 0 and lib.pita_egnn_score_div_workspace_bytes(55, 3) > 0
    # empty batch: OK whatever the pointers
    assert lib.pita_egnn_energy_laplacian(null, 32, 3, 13, null, null, null, 0, null, null) == 0
    assert lib.pita_egnn_forward(null, 64, 5, 22, null, null, null, 0, null, null) == 0
    assert lib.pita_lj_energy_force(null, 0, 55, 1.0, 1.0, 1.0, null, null, null) == 0
    # unsupported shapes
    for args in ((48, 3, 13), (32, 4, 13), (32, 3, 14), (64, 5, 21), (64, 4, 22)):
        rc = lib.pita_egnn_energy(fake, args[0], args[1], args[2], fake, fake, fake, 4, fake, null, null, null)
        assert rc < 0, args
        assert lib.pita_last_error()
    assert lib.pita_egnn_energy_laplacian(fake, 64, 5, 22, fake, fake, fake, 4, fake, null) < 0   # LJ networks only
    assert b"score net" in lib.pita_last_error() or b"LJ" in lib.pita_last_error()
    assert lib.pita_lj_energy_force(fake, 4, 22, 1.0, 1.0, 1.0, fake, null, null) < 0            # reference: n in {13, 55}
    # null pointers with a non-empty batch
    assert lib.pita_egnn_score_div(null, 32, 3, 13, fake, fake, fake, 4, fake, null, 0, null, 0, null) < 0
    assert lib.pita_egnn_energy_laplacian(fake, 32, 3, 13, fake, fake, fake, 4, null, null) < 0
    assert lib.pita_egnn_forward(fake, 64, 5, 22, fake, fake, fake, -1, fake, null) < 0            # negative batch
