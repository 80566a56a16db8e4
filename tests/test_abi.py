"""The C-ABI library loads on a CPU-only box and exports every symbol include/evrep.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from frlw_evd_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    with open(_lib.HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(evrep_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, "no ctypes signature for " + name
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_strings(lib):
    assert lib.evrep_version() == 100
    assert lib.evrep_strerror(0) == b"ok"
    assert b"argument" in lib.evrep_strerror(-1)
    assert b"CUDA" in lib.evrep_strerror(-2)


def test_argument_validation_without_gpu(lib):
    # null pointers / bad sizes are rejected before any CUDA call
    null = ctypes.c_void_p(0)
    assert lib.evrep_decode_dat(null, 8, null, null, null, null, null) == -1
    assert lib.evrep_count_finalize(null, 4, 4, null, 1, null) == -1
    assert lib.evrep_taf_bin_scratch_bytes(0, 10) == -1
    assert lib.evrep_taf_bin_scratch_bytes(240, 304) == 16 + 2 * 240 * 304 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.EvrepError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch
    from frlw_evd_b200 import ops
    ev = torch.zeros((4, 4), dtype=torch.float64)
    with pytest.raises(_lib.EvrepError, match="no CPU fallback"):
        ops.count_image_aos64(ev, (8, 8))
