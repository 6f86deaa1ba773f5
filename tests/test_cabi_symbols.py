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
