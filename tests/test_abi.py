"""The C-ABI library loads without a GPU, exports every symbol declared in
include/spartacus_b200.h, matches the ctypes mirror, and refuses to compute
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import golden_io
import oracle_lib
from spartacus_surface_b200 import _abi, radsurf, RadsurfError
from spartacus_surface_b200._lib import load, LIB_PATH
from spartacus_surface_b200.driver.spartacus_surface_driver import run_radsurf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    header = open(os.path.join(ROOT, "include", "spartacus_b200.h")).read()
    declared = set(re.findall(r"\b(ssb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_abi.EXPORTED_SYMBOLS), declared ^ set(_abi.EXPORTED_SYMBOLS)
    lib = load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in lib.ssb200_version()
    assert os.path.exists(LIB_PATH)


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("ns", [1, 2, 3, 4, 8])
def test_legendre_gauss_matches_oracle(ns):
    a, b = _abi.LegendreGauss(), _abi.LegendreGauss()
    assert load().ssb200_legendre_gauss_init(ns, C.byref(a)) == 0
    assert oracle_lib.load(nofma=True).oracle_legendre_gauss_init(ns, C.byref(b)) == 0
    for f in ("mu", "sin_ang", "tan_ang", "weight", "hweight", "vweight"):
        assert np.array_equal(np.array(getattr(a, f)[:ns]), np.array(getattr(b, f)[:ns])), f
    assert a.vadjustment == b.vadjustment and a.vadjustment2 == b.vadjustment2
    assert load().ssb200_legendre_gauss_init(0, C.byref(a)) == _abi.ERR_ARG
    assert load().ssb200_legendre_gauss_init(17, C.byref(a)) == _abi.ERR_ARG


def test_no_cpu_fallback():
    """Without a CUDA device every solve entry fails loudly with SSB200_ERR_NOGPU."""
    lib = load()
    if lib.ssb200_device_count() > 0:
        pytest.skip("a GPU is present")
    r, _ = golden_io.load_case("simple_surfaces.npz", legendre_gauss_init=lib.ssb200_legendre_gauss_init)
    with pytest.raises(RadsurfError, match="no CPU fallback"):
        run_radsurf(r, radsurf)
    assert lib.ssb200_measure_fp64_peak_tflops(0) < 0
