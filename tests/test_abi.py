"""The C-ABI library loads and exports every symbol include/bppgpu.h declares (no compute without a GPU)."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def header_symbols():
    h = (ROOT / "include" / "bppgpu.h").read_text()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(bppgpu_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol(built_lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(built_lib, s), "libbppgpu.so does not export %s" % s


def test_abi_version_and_error_string(built_lib):
    assert built_lib.bppgpu_abi_version() >= 1
    assert isinstance(built_lib.bppgpu_last_error(), bytes)


def test_struct_layouts_match_library(built_lib):
    """the ctypes mirrors have the layout the library was compiled with"""
    from bpp_phyl_b200 import capi
    assert ctypes.sizeof(capi.ModelDesc) == built_lib.bppgpu_sizeof(0) == 64
    assert ctypes.sizeof(capi.Config) == built_lib.bppgpu_sizeof(1) == 72
    assert ctypes.sizeof(capi.Stats) == built_lib.bppgpu_sizeof(2) == 112


def test_no_cpu_fallback(built_lib):
    """Without a usable device every compute entry point must fail loudly with BPPGPU_E_CUDA."""
    import numpy as np
    from bpp_phyl_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(capi.BppGpuError) as ei:
        capi.Engine(4, 4, 10, np.array([0, 0, 0, 2], np.int32), np.array([0, 1], np.int32), 2, np.eye(4))
    assert ei.value.code == capi.E_CUDA
    m = capi.model_desc(4, capi.MODEL_NONSINGULAR | capi.MODEL_DIAGONALIZABLE, V=np.eye(4), Vinv=np.eye(4),
                        ev_re=np.zeros(4))
    with pytest.raises(capi.BppGpuError) as ei:
        capi.pt_batch(m, [0.1])
    assert ei.value.code == capi.E_CUDA


def test_product_does_not_import_the_oracle():
    for p in (ROOT / "bpp_phyl_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp", ".hpp"):
            assert "oracle" not in p.read_text().replace("oracle/", "").replace("the oracle", ""), p
