"""Synthetic vegetated-urban canopy profiles of the BASELINE shape
(SURVEY.md §8d "Value distributions"): every column has `nlay` layers,
i_representation = VegetatedUrban, monotonically decreasing building fraction,
vegetation in the lowest L_v layers only (clear-only sub-block aloft); in a quarter of
the columns the buildings end 1..4 layers below the canopy top (no-building branch).

Values are a pure function of (seed, global column index, field) through a
counter-based hash, so a column's inputs do not depend on how columns are
sharded over ranks or on the device that generates them.
"""
import numpy as np

from .radsurf_canopy_properties import canopy_properties_type, ITileVegetatedUrban
from .radsurf_sw_spectral_properties import sw_spectral_properties_type
from .radsurf_lw_spectral_properties import lw_spectral_properties_type, StefanBoltzmann

SEED = 20190605


def _uniform(xp, seed, field, col, lay=None):
    """u in [0,1) from a splitmix64-style hash of (seed, field, column[, layer])."""
    if xp is np:
        x = col.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        x = x + np.uint64((seed * 1000003 + field * 7919) & 0xFFFFFFFFFFFF)
        if lay is not None:
            x = x + lay.astype(np.uint64) * np.uint64(0xD1B54A32D192ED03)
        with np.errstate(over="ignore"):
            x ^= x >> np.uint64(30)
            x *= np.uint64(0xBF58476D1CE4E5B9)
            x ^= x >> np.uint64(27)
            x *= np.uint64(0x94D049BB133111EB)
            x ^= x >> np.uint64(31)
        return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    import torch
    # torch has no uint64 arithmetic: emulate with int64 wrap-around and logical shifts
    def lsr(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)
    def c(v):  # python int -> signed 64-bit
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v
    x = col.to(torch.int64) * c(0x9E3779B97F4A7C15)
    x = x + c((seed * 1000003 + field * 7919) & 0xFFFFFFFFFFFF)
    if lay is not None:
        x = x + lay.to(torch.int64) * c(0xD1B54A32D192ED03)
    x = x ^ lsr(x, 30)
    x = x * c(0xBF58476D1CE4E5B9)
    x = x ^ lsr(x, 27)
    x = x * c(0x94D049BB133111EB)
    x = x ^ lsr(x, 31)
    return lsr(x, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def make_synthetic(config, ncol, nlay=16, col_offset=0, device=None, seed=SEED, with_temperatures=False):
    """Return (canopy_props, sw_spectral_props, lw_spectral_props) for `ncol`
    columns starting at global column `col_offset` (with_temperatures=True: plus the dict of
    temperatures the longwave emission / Planck members were made from, the inputs of
    radsurf_fluxes / calc_simple_spectrum_lw).  device=None: numpy host
    arrays; otherwise torch tensors on that device (nlay/istartlay/i_representation
    stay numpy int32 host arrays, as the C ABI wants them)."""
    if device is None:
        xp = np
        col = np.arange(col_offset, col_offset + ncol, dtype=np.int64)
        lay = np.arange(nlay, dtype=np.int64)
        full = lambda shape, v: np.full(shape, v, dtype=np.float64)
        where, minimum, maximum, floor = np.where, np.minimum, np.maximum, np.floor
        contiguous = np.ascontiguousarray
    else:
        import torch
        xp = torch
        col = torch.arange(col_offset, col_offset + ncol, dtype=torch.int64, device=device)
        lay = torch.arange(nlay, dtype=torch.int64, device=device)
        full = lambda shape, v: torch.full(shape, v, dtype=torch.float64, device=device)
        where, minimum, maximum, floor = torch.where, torch.minimum, torch.maximum, torch.floor
        contiguous = lambda t: t.contiguous()
    u = lambda field: _uniform(xp, seed, field, col)                       # (ncol,)
    ul = lambda field: _uniform(xp, seed, field, col[:, None], lay[None, :])  # (ncol, nlay)
    lfrac = (lay.astype(np.float64) if xp is np else lay.to(xp.float64)) / float(nlay)

    cp = canopy_properties_type()
    cp.set_layers(np.full(ncol, nlay, dtype=np.int32))
    cp.i_representation = np.full(ncol, ITileVegetatedUrban, dtype=np.int32)
    flat = lambda a: contiguous(a.reshape(ncol * nlay))
    spec = lambda a, n: contiguous(a.reshape(-1, 1).repeat(n, 1) if xp is np else a.reshape(-1, 1).repeat(1, n))

    cp.cos_sza = contiguous(0.05 + 0.95 * u(1))
    cp.dz = flat(2.0 + 2.0 * ul(2))
    b0, p = 0.25 + 0.25 * u(3), 1.0 + u(4)
    # b0 (1-x)(1-(p-1)x): monotone decreasing like b0 (1-x)^p but free of pow(), whose last bit
    # differs between numpy and torch/CUDA
    bf = b0[:, None] * ((1.0 - lfrac[None, :]) * (1.0 - (p[:, None] - 1.0) * lfrac[None, :]))
    # SURVEY 8(d) wants the top layers of some columns to fall below min_building_fraction (the
    # no-building branch of the urban solvers); neither its power law nor the polynomial above gets
    # there in 16 layers, so in a quarter of the columns the buildings end 1..4 layers below the top
    lb = where(u(23) < 0.25, float(nlay) - 1.0 - floor(4.0 * u(24)), float(nlay) + 0.0 * u(24))
    bf = where(lfrac[None, :] * float(nlay) < lb[:, None], bf, 0.0 * bf)
    cp.building_fraction = flat(bf)
    cp.building_scale = flat((20.0 + 20.0 * u(5))[:, None] + 0.0 * bf)
    lv = 4.0 + floor(9.0 * u(6))
    layf = lfrac[None, :] * float(nlay)
    vf = where(layf < lv[:, None], (0.05 + 0.25 * u(7))[:, None] * (1.0 - bf), 0.0 * bf)
    cp.veg_fraction = flat(vf)
    cp.veg_scale = flat((5.0 + 15.0 * u(8))[:, None] + 0.0 * bf)
    cp.veg_ext = flat((0.1 + 0.4 * u(9))[:, None] + 0.0 * bf)
    cp.veg_fsd = flat((0.5 + 0.5 * u(10))[:, None] + 0.0 * bf)
    one = 1.0 + 0.0 * bf
    cp.veg_contact_fraction = flat(minimum(one, vf / maximum(1.0e-6 * one, 1.0 - bf)))

    nsw, nlw = config.nsw, config.nlw
    sw = sw_spectral_properties_type(nsw)
    ntot = ncol * nlay
    sw.air_ext = full((ntot, nsw), 1.0e-5)
    sw.air_ssa = full((ntot, nsw), 0.999)
    sw.veg_ssa = spec(flat((0.1 + 0.7 * u(11))[:, None] + 0.0 * bf), nsw)
    sw.ground_albedo = spec(0.05 + 0.4 * u(12), nsw)
    sw.roof_albedo = spec(flat(0.05 + 0.4 * ul(13)), nsw)
    sw.wall_albedo = spec(flat(0.05 + 0.4 * ul(14)), nsw)
    sw.wall_specular_frac = full((ntot, nsw), 0.0)
    sw.roof_albedo_dir = sw.roof_albedo.copy() if xp is np else sw.roof_albedo.clone()

    lw = lw_spectral_properties_type(nlw)
    p4 = lambda t: (t * t) * (t * t)
    lw.air_ext = full((ntot, nlw), 1.0e-5)
    lw.air_ssa = full((ntot, nlw), 0.0)
    lw.veg_ssa = spec(flat((0.01 + 0.04 * u(15))[:, None] + 0.0 * bf), nlw)
    t_ground = 283.15 + 10.0 * (2.0 * u(16) - 1.0)
    t_roof = 283.15 + 10.0 * (2.0 * ul(17) - 1.0)
    t_wall = 283.15 + 10.0 * (2.0 * ul(18) - 1.0)
    t_air = 278.15 + 5.0 * (2.0 * ul(19) - 1.0)
    lw.ground_emissivity = spec(0.85 + 0.14 * u(20), nlw)
    lw.roof_emissivity = spec(flat(0.85 + 0.14 * ul(21)), nlw)
    lw.wall_emissivity = spec(flat(0.85 + 0.14 * ul(22)), nlw)
    lw.ground_emission = contiguous(StefanBoltzmann * lw.ground_emissivity * p4(spec(t_ground, nlw)))
    lw.roof_emission = contiguous(StefanBoltzmann * lw.roof_emissivity * p4(spec(flat(t_roof), nlw)))
    lw.wall_emission = contiguous(StefanBoltzmann * lw.wall_emissivity * p4(spec(flat(t_wall), nlw)))
    planck = contiguous(StefanBoltzmann * p4(spec(flat(t_air), nlw)))
    lw.clear_air_planck = planck
    lw.veg_planck = planck.copy() if xp is np else planck.clone()
    lw.veg_air_planck = planck.copy() if xp is np else planck.clone()
    if with_temperatures:
        temps = dict(ground_temperature=contiguous(t_ground), roof_temperature=flat(t_roof),
                     wall_temperature=flat(t_wall), clear_air_temperature=flat(t_air))
        temps["veg_temperature"] = temps["veg_air_temperature"] = temps["clear_air_temperature"]
        return cp, sw, lw, temps
    return cp, sw, lw


def to_host(obj):
    """Copy of an API object with every torch member moved to numpy."""
    import copy
    out = copy.copy(obj)
    for k, v in vars(obj).items():
        if type(v).__module__.startswith("torch"):
            setattr(out, k, np.ascontiguousarray(v.cpu().numpy()))
    return out
