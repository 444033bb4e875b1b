"""Loader of the CPU oracle (oracle/_build/liboracle.so) for tests, smoke() and
bench.py's cpu_baseline / --impl reference legs ONLY.  Never imported by the
product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from spartacus_surface_b200 import _abi
from spartacus_surface_b200.radsurf_interface import marshal, call_radsurf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
ORACLE_NOFMA_SO = os.path.join(ROOT, "oracle", "_build", "liboracle_nofma.so")
# the same restatement with every scalar in _Float128: ground truth of the parity tests
ORACLE_QUAD_SO = os.path.join(ROOT, "oracle", "_build", "liboracle_quad.so")
_libs = {}


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def load(nofma=False, quad=False):
    """nofma=True: the build without FMA contraction (rounding-sensitivity probe);
    quad=True: the extended-precision build (ground truth; ~100x slower)."""
    path = ORACLE_QUAD_SO if quad else (ORACLE_NOFMA_SO if nofma else ORACLE_SO)
    if path not in _libs:
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        P = C.POINTER
        lib.oracle_legendre_gauss_init.argtypes = [C.c_int32, P(_abi.LegendreGauss)]
        lib.oracle_radsurf.argtypes = [P(_abi.Config), P(_abi.CanopyProperties), P(_abi.SwSpectralProperties),
                                       P(_abi.LwSpectralProperties), P(_abi.BoundaryCondsOut), C.c_int32,
                                       C.c_int32, P(_abi.CanopyFlux), P(_abi.CanopyFlux), P(_abi.CanopyFlux),
                                       P(_abi.CanopyFlux), C.c_int32, C.c_int32]
        dp = P(C.c_double)
        lib.oracle_eigen_decomposition_real.argtypes = [C.c_int32, dp, dp, dp]
        lib.oracle_calc_matrices_sw_eig.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_double] + [dp] * 12
        lib.oracle_calc_matrices_lw_eig.argtypes = [C.c_int32, C.c_double] + [dp] * 8
        lib.oracle_schur_invert_sw.argtypes = [C.c_int32, C.c_int32] + [dp] * 8
        lib.oracle_flops_enable.argtypes = [C.c_int]
        lib.oracle_flops_read.restype = C.c_double
        _libs[path] = lib
    return _libs[path]


def legendre_gauss_init(nstream, lg_ref):
    return load().oracle_legendre_gauss_init(nstream, lg_ref)


def make_solver(nthreads=0, nblocksize=16, nofma=False, quad=False):
    """radsurf-compatible callable backed by the oracle."""
    def solver(config, canopy_props, sw, lw, bc_out, istartcol=None, iendcol=None,
               sw_norm_dir=None, sw_norm_diff=None, lw_internal=None, lw_norm=None):
        structs = marshal(config, canopy_props, sw, lw, bc_out, sw_norm_dir, sw_norm_diff, lw_internal, lw_norm)
        rc = call_radsurf(load(nofma, quad).oracle_radsurf, structs, istartcol, iendcol,
                          extra=(C.c_int32(nthreads), C.c_int32(nblocksize)))
        if rc < 0:
            raise RuntimeError(f"oracle_radsurf failed rc={rc}")
        return rc
    return solver


def fcol(a):
    """numpy (r, c) matrix -> column-major flat double* keeper."""
    arr = np.asfortranarray(a, dtype=np.float64)
    return arr, arr.ctypes.data_as(C.POINTER(C.c_double))
