"""B200-native SPARTACUS-Surface solver path.

Python mirror of the reference's embedding API (the six derived types and the
`radsurf` entry, doc/spartacus_surface_documentation.tex:986-1222) on top of
the C-ABI library csrc/libspartacus_b200.so.
"""
from .radsurf_config import config_type
from .radsurf_canopy_properties import (canopy_properties_type, ITileFlat, ITileForest, ITileUrban,
                                        ITileVegetatedUrban, ITileSimpleUrban, ITileInfiniteStreet)
from .radsurf_sw_spectral_properties import sw_spectral_properties_type
from .radsurf_lw_spectral_properties import lw_spectral_properties_type
from .radsurf_canopy_flux import canopy_flux_type
from .radsurf_boundary_conds_out import boundary_conds_out_type
from .radsurf_interface import radsurf, RadsurfError
from .radsurf_simple_spectrum import calc_simple_spectrum_lw

__all__ = [
    "config_type", "canopy_properties_type", "sw_spectral_properties_type",
    "lw_spectral_properties_type", "canopy_flux_type", "boundary_conds_out_type", "radsurf",
    "RadsurfError", "calc_simple_spectrum_lw", "ITileFlat", "ITileForest", "ITileUrban",
    "ITileVegetatedUrban", "ITileSimpleUrban", "ITileInfiniteStreet",
]
