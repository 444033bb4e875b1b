"""calc_simple_spectrum_lw (radsurf/radsurf_simple_spectrum.F90:20-68): broadband
sigma*T^4 fill of the LW emission / Planck arrays, the input preparation that the
reference driver performs before each radsurf call.  numpy on host arrays; for
device-resident (torch CUDA) members the library's ssb200_calc_simple_spectrum_lw_device
(SURVEY section 8 "next" row 2: only the temperatures have to cross PCIe)."""
import ctypes as C

from ._arrays import is_torch
from .radsurf_lw_spectral_properties import StefanBoltzmann


def calc_simple_spectrum_lw(config, canopy_props, lw_spectral_props, istartcol=None, iendcol=None):
    cp, lw = canopy_props, lw_spectral_props
    if lw.nspec > 1:
        raise ValueError("Simple longwave spectrum only possible with one input spectral interval")
    c1 = 0 if istartcol is None else istartcol - 1
    c2 = cp.ncol if iendcol is None else iendcol
    l1 = int(cp.istartlay[c1]) - 1
    l2 = int(cp.istartlay[c2 - 1]) - 1 + int(cp.nlay[c2 - 1])
    if is_torch(lw.ground_emission):
        from ._lib import load, last_error
        ptr = lambda a: C.c_void_p(a.data_ptr()) if a is not None else None
        urban = lw.roof_emissivity is not None
        s = lw.as_struct()
        rc = load().ssb200_calc_simple_spectrum_lw_device(
            C.byref(s), cp.ncol, cp.ntotlay, c1 + 1, c2, l1 + 1, l2, ptr(cp.ground_temperature),
            ptr(cp.roof_temperature) if urban else None, ptr(cp.wall_temperature) if urban else None,
            ptr(cp.clear_air_temperature), ptr(cp.veg_temperature) if config.do_vegetation else None,
            ptr(cp.veg_air_temperature) if config.do_vegetation else None, None)
        if rc != 0:
            raise RuntimeError(f"ssb200_calc_simple_spectrum_lw_device failed ({rc}): {last_error()}")
        return
    lw.ground_emission[c1:c2, 0] = StefanBoltzmann * lw.ground_emissivity[c1:c2, 0] * cp.ground_temperature[c1:c2] ** 4
    if l2 > l1:
        if lw.roof_emissivity is not None:
            lw.roof_emission[l1:l2, 0] = StefanBoltzmann * lw.roof_emissivity[l1:l2, 0] * cp.roof_temperature[l1:l2] ** 4
            lw.wall_emission[l1:l2, 0] = StefanBoltzmann * lw.wall_emissivity[l1:l2, 0] * cp.wall_temperature[l1:l2] ** 4
        lw.clear_air_planck[l1:l2, 0] = StefanBoltzmann * cp.clear_air_temperature[l1:l2] ** 4
        if config.do_vegetation:
            lw.veg_planck[l1:l2, 0] = StefanBoltzmann * cp.veg_temperature[l1:l2] ** 4
            lw.veg_air_planck[l1:l2, 0] = StefanBoltzmann * cp.veg_air_temperature[l1:l2] ** 4
