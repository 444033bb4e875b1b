"""radsurf: the public entry of the path (radsurf/radsurf_interface.F90:20-25).

Same argument list and meaning as the reference subroutine; the body is the
B200 library (ssb200_radsurf for host arrays, ssb200_radsurf_device when the
members are torch CUDA tensors).  The reference aborts on error
(utilities/radiation_io.F90:46-54); here errors raise RadsurfError.
"""
import ctypes as C

from ._arrays import is_torch
from ._lib import load, last_error


class RadsurfError(RuntimeError):
    pass


def marshal(config, canopy_props, sw_spectral_props, lw_spectral_props, bc_out,
            sw_norm_dir=None, sw_norm_diff=None, lw_internal=None, lw_norm=None):
    """Build the C structs of include/spartacus_b200.h from the API objects.

    Returned objects own no array memory: the caller's arrays must outlive
    the call (same ownership rule as the Fortran allocatables).
    """
    structs = dict(
        config=config.as_struct(),
        canopy=canopy_props.as_struct(),
        sw=sw_spectral_props.as_struct() if sw_spectral_props is not None else None,
        lw=lw_spectral_props.as_struct() if lw_spectral_props is not None else None,
        bc=bc_out.as_struct(),
        sw_norm_dir=sw_norm_dir.as_struct() if sw_norm_dir is not None else None,
        sw_norm_diff=sw_norm_diff.as_struct() if sw_norm_diff is not None else None,
        lw_internal=lw_internal.as_struct() if lw_internal is not None else None,
        lw_norm=lw_norm.as_struct() if lw_norm is not None else None,
    )
    return structs


def _ref(s):
    return C.byref(s) if s is not None else None


def call_radsurf(fn, structs, istartcol, iendcol, extra=()):
    s = structs
    return fn(_ref(s["config"]), _ref(s["canopy"]), _ref(s["sw"]), _ref(s["lw"]), _ref(s["bc"]),
              int(istartcol or 0), int(iendcol or 0), _ref(s["sw_norm_dir"]), _ref(s["sw_norm_diff"]),
              _ref(s["lw_internal"]), _ref(s["lw_norm"]), *extra)


def radsurf(config, canopy_props, sw_spectral_props, lw_spectral_props, bc_out,
            istartcol=None, iendcol=None, sw_norm_dir=None, sw_norm_diff=None,
            lw_internal=None, lw_norm=None, stream=None):
    """Solve columns istartcol..iendcol (1-based inclusive, default all)."""
    lib = load()
    structs = marshal(config, canopy_props, sw_spectral_props, lw_spectral_props, bc_out,
                      sw_norm_dir, sw_norm_diff, lw_internal, lw_norm)
    device = is_torch(canopy_props.dz) if canopy_props.dz is not None else is_torch(bc_out.sw_albedo
                                                                                    if bc_out.sw_albedo is not None
                                                                                    else bc_out.lw_emissivity)
    if device:
        rc = call_radsurf(lib.ssb200_radsurf_device, structs, istartcol, iendcol,
                          extra=(C.c_void_p(stream or 0), None))
    else:
        rc = call_radsurf(lib.ssb200_radsurf, structs, istartcol, iendcol)
    if rc < 0:
        raise RadsurfError(f"ssb200_radsurf failed (rc={rc}): {last_error()}")
    return rc


def radsurf_fluxes(config, canopy_props, sw_spectral_props, lw_spectral_props, bc_out, istartcol=None, iendcol=None,
                   sw_flux=None, lw_flux=None, top_flux_dn_sw=None, top_flux_dn_direct_sw=None, top_flux_dn_lw=None,
                   ground_temperature=None, roof_temperature=None, wall_temperature=None,
                   clear_air_temperature=None, veg_temperature=None, veg_air_temperature=None):
    """The reference driver's sequence for a block of columns in one call
    (driver/spartacus_surface_driver.F90:206-261): calc_simple_spectrum_lw, radsurf,
    scale of the normalised fluxes by the top-of-canopy fluxes and sum into `sw_flux` / `lw_flux`.
    The four normalised flux objects never leave the device; members of the spectral property
    objects that are None take read_input's defaults or are derived from the temperatures
    (include/spartacus_b200.h: ssb200_radsurf_fluxes).  Host (numpy) arrays."""
    from . import _abi
    lib = load()
    structs = marshal(config, canopy_props, sw_spectral_props, lw_spectral_props, bc_out)
    keep = [np_ for np_ in (top_flux_dn_sw, top_flux_dn_direct_sw, top_flux_dn_lw, ground_temperature,
                            roof_temperature, wall_temperature, clear_air_temperature, veg_temperature,
                            veg_air_temperature)]
    drv = _abi.DriverInputs(*[_abi.dptr(a) for a in keep])
    fsw = sw_flux.as_struct() if sw_flux is not None else None
    flw = lw_flux.as_struct() if lw_flux is not None else None
    s = structs
    rc = lib.ssb200_radsurf_fluxes(_ref(s["config"]), _ref(s["canopy"]), _ref(s["sw"]), _ref(s["lw"]), C.byref(drv),
                                   _ref(s["bc"]), int(istartcol or 0), int(iendcol or 0), _ref(fsw), _ref(flw))
    if rc < 0:
        raise RadsurfError(f"ssb200_radsurf_fluxes failed (rc={rc}): {last_error()}")
    return rc


def _real_members(*objs):
    import numpy as np
    for o in objs:
        if o is None:
            continue
        for v in vars(o).values():
            if isinstance(v, np.ndarray) and v.dtype.kind == "f":
                yield v


def to_single(obj):
    """Copy of an API object with every float64 numpy member converted to float32 (what a
    -DSINGLE_PRECISION build of the reference holds: jprb = real32, utilities/parkind1.F90:45-49)."""
    import copy
    import numpy as np
    out = copy.copy(obj)
    for k, v in vars(obj).items():
        if isinstance(v, np.ndarray) and v.dtype == np.float64:
            setattr(out, k, np.ascontiguousarray(v.astype(np.float32)))
    return out


def radsurf_sp(config, canopy_props, sw_spectral_props, lw_spectral_props, bc_out,
               istartcol=None, iendcol=None, sw_norm_dir=None, sw_norm_diff=None,
               lw_internal=None, lw_norm=None):
    """radsurf for single-precision (float32) host arrays: ssb200_radsurf_sp.  The arrays cross PCIe
    as float32, the solve runs in FP64 on the device, outputs are the FP64 results rounded to float32."""
    import numpy as np
    lib = load()
    objs = (canopy_props, sw_spectral_props, lw_spectral_props, bc_out, sw_norm_dir, sw_norm_diff, lw_internal, lw_norm)
    if not all(v.dtype == np.float32 for v in _real_members(*objs)):
        raise RadsurfError("radsurf_sp: every real array member must be float32")
    structs = marshal(config, canopy_props, sw_spectral_props, lw_spectral_props, bc_out,
                      sw_norm_dir, sw_norm_diff, lw_internal, lw_norm)
    rc = call_radsurf(lib.ssb200_radsurf_sp, structs, istartcol, iendcol)
    if rc < 0:
        raise RadsurfError(f"ssb200_radsurf_sp failed (rc={rc}): {last_error()}")
    return rc
