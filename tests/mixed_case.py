"""Synthetic edge-case input shared by the CPU (host check) and GPU parity tests: ragged
numbers of layers, flat, forest, urban and vegetated-urban tiles in one call, night-time columns, several spectral
intervals with different optical properties, a direct ground albedo."""
import numpy as np

from spartacus_surface_b200 import config_type
from spartacus_surface_b200.radsurf_canopy_properties import (ITileFlat, ITileForest, ITileUrban,
                                                              ITileVegetatedUrban, ITileSimpleUrban,
                                                              ITileInfiniteStreet)
from spartacus_surface_b200.synthetic import make_synthetic

NLAY_MAX = 9


def mixed_config(streams=2, nsw=2, nlw=3):
    return config_type(do_sw=True, do_lw=True, nsw=nsw, nlw=nlw, use_sw_direct_albedo=True,
                       n_vegetation_region_urban=2, n_vegetation_region_forest=2,
                       n_stream_sw_urban=streams, n_stream_lw_urban=streams, n_stream_sw_forest=streams,
                       n_stream_lw_forest=streams)


def make_mixed(cfg, ncol=300, seed=7):
    """(canopy_props, sw, lw) cut out of the regular synthetic canopy."""
    rng = np.random.default_rng(seed)
    cp, sw, lw = make_synthetic(cfg, ncol, NLAY_MAX)
    # the single-layer urban models sit AFTER multi-layer columns (istartlay(j) != j): their
    # per-column members must land in their own column
    rep = np.array([ITileVegetatedUrban, ITileForest, ITileUrban, ITileFlat, ITileVegetatedUrban,
                    ITileForest, ITileSimpleUrban, ITileInfiniteStreet], dtype=np.int32)[np.arange(ncol) % 8]
    nlay = rng.integers(1, NLAY_MAX + 1, size=ncol).astype(np.int32)
    nlay[rep == ITileFlat] = 0
    nlay[(rep == ITileSimpleUrban) | (rep == ITileInfiniteStreet)] = 1
    # pack the ragged layers: keep the lowest nlay[j] layers of every column
    keep = (np.arange(NLAY_MAX)[None, :] < nlay[:, None]).reshape(-1)
    for obj in (cp, sw, lw):
        for k, v in list(vars(obj).items()):
            if isinstance(v, np.ndarray) and v.dtype == np.float64 and v.shape[0] == ncol * NLAY_MAX:
                setattr(obj, k, np.ascontiguousarray(v[keep]))
    cp.set_layers(nlay)
    cp.i_representation = rep
    cp.cos_sza = cp.cos_sza.copy()
    cp.cos_sza[::7] = 0.0      # night
    cp.cos_sza[3::11] = -0.2
    # spectral intervals differ
    for arr, lo, hi in ((sw.air_ext, 0.5, 50.0), (sw.veg_ssa, 0.6, 1.2), (sw.ground_albedo, 0.5, 1.5),
                        (sw.wall_albedo, 0.7, 1.3), (lw.air_ext, 1.0, 300.0), (lw.veg_ssa, 0.5, 2.0),
                        (lw.ground_emissivity, 0.95, 1.0), (lw.wall_emissivity, 0.9, 1.0)):
        n = arr.shape[1]
        arr *= np.linspace(lo, hi, n)[None, :]
    sw.ground_albedo_dir = np.ascontiguousarray(0.8 * sw.ground_albedo)
    sw.wall_specular_frac = np.ascontiguousarray(sw.wall_specular_frac + 0.25)
    return cp, sw, lw
