"""driver_config_type (driver/spartacus_surface_config.F90:21-165)."""
import math
from dataclasses import dataclass

from ..namelist import read_namelist


@dataclass
class driver_config_type:
    do_parallel: bool = True
    nblocksize: int = 16
    nrepeat: int = 1
    istartcol: int = 1
    iendcol: int = 0
    iverbose: int = 3
    do_conservation_check: bool = False
    cos_sza_override: float = -1.0
    ground_sw_albedo: float = -1.0
    roof_sw_albedo: float = -1.0
    wall_sw_albedo: float = -1.0
    ground_lw_emissivity: float = -1.0
    roof_lw_emissivity: float = -1.0
    wall_lw_emissivity: float = -1.0
    vegetation_fraction: float = -1.0
    vegetation_extinction: float = -1.0
    vegetation_extinction_scaling: float = -1.0
    vegetation_fsd: float = -1.0
    vegetation_sw_ssa: float = -1.0
    vegetation_lw_ssa: float = -1.0
    top_flux_dn_sw: float = -1.0
    top_flux_dn_direct_sw: float = -1.0
    top_flux_dn_lw: float = -1.0
    isurfacetype: int = -1

    def read(self, file_name, **overrides):
        """read_config_from_namelist (:76-165); `overrides` plays the role of
        the test suites' change_namelist.sh edits."""
        group = dict(read_namelist(file_name).get("radsurf_driver", {}))
        group.update(overrides)
        solar_zenith_angle = -100.0
        for key, val in group.items():
            if key == "cos_solar_zenith_angle":
                self.cos_sza_override = float(val)
            elif key == "solar_zenith_angle":
                solar_zenith_angle = float(val)
            elif key == "vegetation_lw_ssa":
                # not in the reference namelist (:100-106): cannot be set there
                raise ValueError("vegetation_lw_ssa is not a member of &radsurf_driver")
            elif hasattr(self, key):
                cur = getattr(self, key)
                setattr(self, key, type(cur)(val))
            else:
                raise ValueError(f"unknown &radsurf_driver namelist entry '{key}'")
        if self.cos_sza_override == -1.0 and 0.0 <= solar_zenith_angle <= 180.0:
            self.cos_sza_override = math.cos(solar_zenith_angle * math.pi / 180.0)
        return self
