"""config_type of the reference (radsurf/radsurf_config.F90:32-113).

Same member names, defaults (:38-95), namelist group (`&radsurf`, :153-161)
and `consolidate` semantics (:250-270: nswinternal=nsw, nlwinternal=nlw and
the four Legendre-Gauss tables).  The quadrature tables are produced by the
product library's host-side ssb200_legendre_gauss_init.
"""
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from .namelist import read_namelist


@dataclass
class config_type:
    do_sw: bool = True
    do_lw: bool = True
    use_sw_direct_albedo: bool = False
    do_vegetation: bool = True
    do_urban: bool = True
    n_vegetation_region_forest: int = 1
    n_vegetation_region_urban: int = 1
    nsw: int = 1
    nlw: int = 1
    n_stream_sw_forest: int = 4
    n_stream_sw_urban: int = 4
    n_stream_lw_forest: int = 4
    n_stream_lw_urban: int = 4
    use_symmetric_vegetation_scale_forest: bool = True
    use_symmetric_vegetation_scale_urban: bool = True
    vegetation_isolation_factor_forest: float = 0.0
    vegetation_isolation_factor_urban: float = 0.0
    # default-real literals in the reference (radsurf_config.F90:76-81): the
    # value stored is the single-precision 1.0e-6 promoted to double
    min_vegetation_fraction: float = float(np.float32(1.0e-6))
    min_building_fraction: float = float(np.float32(1.0e-6))
    do_save_broadband_flux: bool = True
    do_save_spectral_flux: bool = False
    do_save_flux_profile: bool = False
    iverbose: int = 3
    nswinternal: int = 0
    nlwinternal: int = 0
    lg_sw_forest: object = None
    lg_sw_urban: object = None
    lg_lw_forest: object = None
    lg_lw_urban: object = None
    _consolidated: bool = field(default=False, repr=False)

    _NAMELIST_KEYS = (
        "do_sw", "do_lw", "use_sw_direct_albedo", "do_vegetation", "do_urban", "nsw", "nlw",
        "n_stream_sw_forest", "n_stream_sw_urban", "n_stream_lw_forest", "n_stream_lw_urban",
        "iverbose", "do_save_spectral_flux", "do_save_broadband_flux", "do_save_flux_profile",
        "n_vegetation_region_forest", "n_vegetation_region_urban",
        "use_symmetric_vegetation_scale_forest", "use_symmetric_vegetation_scale_urban",
        "vegetation_isolation_factor_forest", "vegetation_isolation_factor_urban",
        "min_vegetation_fraction", "min_building_fraction")

    def read(self, file_name):
        """read_config_from_namelist (radsurf_config.F90:125-247)."""
        group = read_namelist(file_name).get("radsurf", {})
        for key, val in group.items():
            if key not in self._NAMELIST_KEYS:
                raise ValueError(f"unknown &radsurf namelist entry '{key}'")
            cur = getattr(self, key)
            if isinstance(cur, bool):
                val = bool(val)
            elif isinstance(cur, int):
                val = int(val)
            else:
                val = float(val)
            setattr(self, key, val)
        return self

    def consolidate(self, legendre_gauss_init=None):
        """consolidate_config (radsurf_config.F90:250-270).

        `legendre_gauss_init(nstream, LegendreGauss*) -> int` defaults to the
        product library's ssb200_legendre_gauss_init; tests may pass the
        oracle's to cross-check the tables.
        """
        import ctypes as C
        if legendre_gauss_init is None:
            from ._lib import load
            legendre_gauss_init = load().ssb200_legendre_gauss_init
        self.nswinternal = self.nsw
        self.nlwinternal = self.nlw
        for name, ns in (("lg_sw_forest", self.n_stream_sw_forest), ("lg_sw_urban", self.n_stream_sw_urban),
                         ("lg_lw_forest", self.n_stream_lw_forest), ("lg_lw_urban", self.n_stream_lw_urban)):
            lg = _abi.LegendreGauss()
            rc = legendre_gauss_init(int(ns), C.byref(lg))
            if rc != 0:
                raise ValueError(f"legendre_gauss init failed for nstream={ns} (rc={rc})")
            setattr(self, name, lg)
        self._consolidated = True
        return self

    def as_struct(self):
        if not self._consolidated:
            raise RuntimeError("config_type.consolidate() must be called before radsurf")
        c = _abi.Config()
        for key in ("do_sw", "do_lw", "use_sw_direct_albedo", "do_vegetation", "do_urban",
                    "n_vegetation_region_forest", "n_vegetation_region_urban",
                    "use_symmetric_vegetation_scale_forest", "use_symmetric_vegetation_scale_urban",
                    "iverbose"):
            setattr(c, key, int(getattr(self, key)))
        c.nsw = int(self.nswinternal)
        c.nlw = int(self.nlwinternal)
        c.vegetation_isolation_factor_forest = self.vegetation_isolation_factor_forest
        c.vegetation_isolation_factor_urban = self.vegetation_isolation_factor_urban
        c.min_vegetation_fraction = self.min_vegetation_fraction
        c.min_building_fraction = self.min_building_fraction
        c.lg_sw_forest, c.lg_sw_urban = self.lg_sw_forest, self.lg_sw_urban
        c.lg_lw_forest, c.lg_lw_urban = self.lg_lw_forest, self.lg_lw_urban
        return c
