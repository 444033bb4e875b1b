"""Builds/loads tests/_build/libhostcheck.so: the product's solver bodies
(csrc/*.cuh) compiled for the host and driven by a serial loop.  Test
infrastructure only (see tests/hostcheck/hostcheck.cpp)."""
import ctypes as C
import os
import subprocess

from spartacus_surface_b200 import _abi
from spartacus_surface_b200.radsurf_interface import marshal, call_radsurf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
SO = os.path.join(ROOT, "tests", "_build", "libhostcheck.so")
CSRC = os.path.join(ROOT, "spartacus_surface_b200", "csrc")
_lib = None


def build(force=False):
    deps = [SRC, os.path.join(os.path.dirname(SRC), "asymtx_qr.hpp")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    deps.append(os.path.join(ROOT, "include", "spartacus_b200.h"))
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-Wno-maybe-uninitialized", "-o", SO, SRC])


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(SO)
        P = C.POINTER
        lib.hostcheck_radsurf.argtypes = [P(_abi.Config), P(_abi.CanopyProperties), P(_abi.SwSpectralProperties),
                                          P(_abi.LwSpectralProperties), P(_abi.BoundaryCondsOut), C.c_int32,
                                          C.c_int32, P(_abi.CanopyFlux), P(_abi.CanopyFlux), P(_abi.CanopyFlux),
                                          P(_abi.CanopyFlux), C.c_int64, C.c_int32]
        _lib = lib
    return _lib


def make_solver(budget_doubles=0, fast=False, fused=False, records=False, generic_jacobi=False, stage=True):
    """fast: register-resident bodies (split layer / sweeps kernels); fused: the column-resident
    bodies where they exist (1 and 2 streams), register-resident elsewhere; records: register-resident
    layer bodies followed by the record sweeps (1 and 2 streams) - what the device runs by default;
    generic_jacobi: the generic bodies with the symmetrised Jacobi eigen-systems the DEVICE build of the
    generic kernels uses (the default host build keeps the reference-order QR solver: bit-identity test); stage=False: the register-resident bodies read and write the
    caller's arrays in place instead of the level-major staging buffers of the chunk (csrc/ssb_stage.cuh)."""
    def solver(config, canopy_props, sw, lw, bc_out, istartcol=None, iendcol=None,
               sw_norm_dir=None, sw_norm_diff=None, lw_internal=None, lw_norm=None):
        structs = marshal(config, canopy_props, sw, lw, bc_out, sw_norm_dir, sw_norm_diff, lw_internal, lw_norm)
        rc = call_radsurf(load().hostcheck_radsurf, structs, istartcol, iendcol,
                          extra=(C.c_int64(budget_doubles), C.c_int32((4 if generic_jacobi else (3 if records else (2 if fused else (1 if fast else 0)))) | (0 if stage else 8))))
        if rc < 0:
            raise RuntimeError(f"hostcheck_radsurf failed rc={rc}")
        return rc
    return solver
