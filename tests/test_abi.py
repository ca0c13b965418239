"""The C-ABI library loads and exports every symbol include/insmos_b200.h declares (no compute calls)."""
import os
import re

from insmos_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "insmos_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(insmos_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libinsmos_b200.so does not export %s" % s
        assert s in _lib.PROTOTYPES, "no ctypes prototype for %s" % s
    assert set(_lib.PROTOTYPES) == set(syms)


def test_host_helpers():
    lib = _lib.load()
    assert lib.insmos_hash_capacity(1) == 1024
    assert lib.insmos_hash_capacity(600000) == 2097152           # 2n slots: load <= 0.5 worst case, ~0.2 on LiDAR input
    assert lib.insmos_rulebook_entries_capacity(129, 27, 128) == 2 * 128 * 27
    assert b"sm_100a" in lib.insmos_version()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from insmos_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor required"):
        ops.gather_rows(torch.zeros(4, 3), torch.zeros(4, dtype=torch.int32))
