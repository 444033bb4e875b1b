"""Array helpers shared by the API types.

Arrays are stored so that their MEMORY equals the reference's Fortran layout:
Fortran (nspec, ncol) / (nspec, ntotlay) is a C-contiguous array of shape
(ncol, nspec) / (ntotlay, nspec).  Members may be numpy arrays (host) or torch
CUDA tensors (device-resident variant); `None` means "not allocated".
"""
import ctypes as C

import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def is_torch(a):
    return type(a).__module__.startswith("torch")


def dptr(a):
    if a is None:
        return C.cast(None, _dp)
    if is_torch(a):
        import torch
        assert a.dtype == torch.float64 and a.is_contiguous()
        return C.cast(a.data_ptr(), _dp)
    # float32 members are the single-precision storage variant (radsurf_interface.radsurf_sp): the
    # address is carried in the same struct slot (the _sp structs of the header are layout-identical)
    assert a.dtype in (np.float64, np.float32) and a.flags["C_CONTIGUOUS"], "C-contiguous real array required"
    return a.ctypes.data_as(_dp)


def iptr(a):
    if a is None:
        return C.cast(None, _ip)
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"], "int32 host array required"
    return a.ctypes.data_as(_ip)


def zeros(shape, like=None, device=None):
    if device is not None:
        import torch
        return torch.zeros(shape, dtype=torch.float64, device=device)
    return np.zeros(shape, dtype=np.float64)
