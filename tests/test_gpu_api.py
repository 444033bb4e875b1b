"""GPU tests of the rest of the C ABI: device-resident entry, option switches
(kernel family, pipelined host entry, scratch budget), column sub-ranges,
the fused scale / sum / check follow-on steps, error codes."""
import ctypes as C

import numpy as np
import pytest

import golden_io
import oracle_lib
import parity
from spartacus_surface_b200 import (radsurf, RadsurfError, config_type, canopy_flux_type,
                                    boundary_conds_out_type)
from spartacus_surface_b200._lib import load
from spartacus_surface_b200.radsurf_canopy_flux import ALL_FIELDS
from spartacus_surface_b200.synthetic import make_synthetic

pytestmark = pytest.mark.gpu

NCOL, NLAY = 3000, 9


def _cfg(streams=2):
    return config_type(n_vegetation_region_urban=2, n_vegetation_region_forest=2, n_stream_sw_urban=streams,
                       n_stream_lw_urban=streams, n_stream_sw_forest=streams, n_stream_lw_forest=streams).consolidate()


def _outputs(cfg, ncol, ntot, device=None, profile=True):
    bc = boundary_conds_out_type().allocate(ncol, 1, 1, device=device)
    fl = [canopy_flux_type().allocate(cfg, ncol, ntot, 1, use_direct=d, do_save_flux_profile=profile, device=device)
          for d in (True, True, False, False)]
    return bc, fl


def _as_dict(fl, bc):
    names = ("sw_norm_dir", "sw_norm_diff", "lw_internal", "lw_norm")
    get = lambda a: a.cpu().numpy() if hasattr(a, "cpu") else a
    out = {n: {k: get(getattr(f, k)) for k in ALL_FIELDS if getattr(f, k) is not None} for n, f in zip(names, fl)}
    out["bc"] = {k: get(getattr(bc, k)) for k in golden_io.BC_FIELDS}
    return out


def _solve_host(cfg, cp, sw, lw, c1=None, c2=None):
    bc, fl = _outputs(cfg, cp.ncol, cp.ntotlay)
    assert radsurf(cfg, cp, sw, lw, bc, c1, c2, *fl) == 0
    return _as_dict(fl, bc)


def _same(a, b):
    for n, f in a.items():
        for k, v in f.items():
            assert np.array_equal(v, b[n][k]), (n, k, float(np.abs(v - b[n][k]).max()))


def copy_obj(obj):
    import copy
    return copy.copy(obj)


def _to_device(obj):
    import copy
    import torch
    out = copy.copy(obj)
    for k, v in vars(obj).items():
        if isinstance(v, np.ndarray) and v.dtype == np.float64:
            setattr(out, k, torch.from_numpy(v).cuda())
    return out


def test_device_entry_equals_host_entry():
    """Same inputs through ssb200_radsurf (host arrays) and ssb200_radsurf_device (HBM-resident)."""
    import torch
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, NCOL, NLAY)
    host = _solve_host(cfg, cp, sw, lw)
    dcp, dsw, dlw = _to_device(cp), _to_device(sw), _to_device(lw)
    bc, fl = _outputs(cfg, NCOL, dcp.ntotlay, device="cuda:0")
    assert radsurf(cfg, dcp, dsw, dlw, bc, None, None, *fl) == 0
    torch.cuda.synchronize()
    _same(host, _as_dict(fl, bc))


def test_kernel_families_agree_and_match_oracle():
    lib = load()
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 512, NLAY)
    res = {}
    for fast in (0, 1):  # 0: generic kernels (test-only path), 1: register-resident (the product path)
        lib.ssb200_set_option(b"fast_kernels", fast)
        res[fast] = _solve_host(cfg, cp, sw, lw)
    lib.ssb200_set_option(b"fast_kernels", 1)
    lib.ssb200_set_option(b"fused_kernels", 1)  # column-resident experiment (off by default)
    res[2] = _solve_host(cfg, cp, sw, lw)
    lib.ssb200_set_option(b"fused_kernels", 0)
    ora = []
    for kw in ({}, {"nofma": True}, {"quad": True}):
        bc, fl = _outputs(cfg, cp.ncol, cp.ntotlay)
        oracle_lib.make_solver(**kw)(cfg, cp, sw, lw, bc, None, None, *fl)
        ora.append(_as_dict(fl, bc))
    # err(gpu, truth) <= max(1e-9, 2 err(reference_fp64, truth)), tests/parity.py
    for fast in (0, 1, 2):
        ok, worst, lines = parity.check(res[fast], ora[2], ora[0], ora[1], factor=4.0 if fast == 0 else parity.REF_FACTOR)
        assert ok, (fast, lines[:5])


def test_pipeline_and_chunking_do_not_change_results():
    lib = load()
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 70000, 4)  # large enough for the pipelined path (several blocks)
    base = _solve_host(cfg, cp, sw, lw)
    # the pipelined and the chunked runs launch the same kernels on other column blocks: same bits
    lib.ssb200_set_option(b"pipeline", 0)
    try:
        _same(base, _solve_host(cfg, cp, sw, lw))
    finally:
        lib.ssb200_set_option(b"pipeline", 1)
    lib.ssb200_set_option(b"sort_columns", 1)  # columns in segment order instead of input order
    try:
        _same(base, _solve_host(cfg, cp, sw, lw))
    finally:
        lib.ssb200_set_option(b"sort_columns", 0)
    lib.ssb200_set_option(b"scratch_budget_bytes", 64 << 20)  # forces many chunks
    try:
        _same(base, _solve_host(cfg, cp, sw, lw))
    finally:
        lib.ssb200_set_option(b"scratch_budget_bytes", 0)


def test_column_range_untouched_outside():
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 200, NLAY)
    full = _solve_host(cfg, cp, sw, lw)
    bc, fl = _outputs(cfg, cp.ncol, cp.ntotlay)
    for f in fl:
        f.fill(7.0)
    for k in golden_io.BC_FIELDS:
        getattr(bc, k)[...] = 7.0
    assert radsurf(cfg, cp, sw, lw, bc, 51, 120, *fl) == 0
    part = _as_dict(fl, bc)
    for n, fields in part.items():
        for k, v in fields.items():
            rows = cp.ncol if v.shape[0] == cp.ncol else cp.ntotlay
            lo, hi = (50, 120) if rows == cp.ncol else (50 * NLAY, 120 * NLAY)
            assert np.array_equal(v[lo:hi], full[n][k][lo:hi]), (n, k)
            assert np.all(v[:lo] == 7.0) and np.all(v[hi:] == 7.0), (n, k)


def test_scale_sum_check_on_device():
    import torch
    lib = load()
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 400, NLAY)
    bc, fl = _outputs(cfg, cp.ncol, cp.ntotlay)
    assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl) == 0
    rng = np.random.default_rng(1)
    top_dir = rng.uniform(100, 800, size=(cp.ncol, 1))
    top_dif = rng.uniform(10, 200, size=(cp.ncol, 1))
    # host (numpy) versions = the reference's scale / sum (radsurf_canopy_flux.F90:212-282,399-460)
    import copy
    h_dir, h_dif = copy.deepcopy(fl[0]), copy.deepcopy(fl[1])
    h_dir.scale(cp.nlay, top_dir)
    h_dif.scale(cp.nlay, top_dif)
    h_sum = canopy_flux_type().allocate(cfg, cp.ncol, cp.ntotlay, 1, use_direct=True)
    h_sum.sum(h_dir, h_dif)
    # device versions through the C ABI
    def to_dev(f):
        g = canopy_flux_type().allocate(cfg, cp.ncol, cp.ntotlay, 1, use_direct=True, device="cuda:0")
        for k in ALL_FIELDS:
            a = getattr(f, k)
            if a is not None:
                getattr(g, k).copy_(torch.from_numpy(a))
        return g
    d_dir, d_dif = to_dev(fl[0]), to_dev(fl[1])
    d_dir.scale(cp.nlay, torch.from_numpy(top_dir).cuda())
    d_dif.scale(cp.nlay, torch.from_numpy(top_dif).cuda())
    d_sum = canopy_flux_type().allocate(cfg, cp.ncol, cp.ntotlay, 1, use_direct=True, device="cuda:0")
    d_sum.sum(d_dir, d_dif)
    torch.cuda.synchronize()
    for k in ALL_FIELDS:
        a = getattr(h_sum, k)
        if a is not None:
            assert np.array_equal(a, getattr(d_sum, k).cpu().numpy()), k
    # energy budget residual per column (radsurf_canopy_flux.F90:534-535)
    res = torch.zeros(cp.ncol, dtype=torch.float64, device="cuda:0")
    s, c = d_sum.as_struct(), cp.as_struct()
    assert lib.ssb200_canopy_flux_check_device(C.byref(s), C.byref(c), C.c_void_p(res.data_ptr()), None) == 0
    torch.cuda.synchronize()
    tab = h_sum.check(cp, iverbose=0)
    assert np.allclose(res.cpu().numpy(), tab[:, 7], rtol=0, atol=1e-9)
    assert np.abs(tab[:, 7]).max() < 1e-9 * np.abs(tab[:, 6]).max()


def test_error_codes():
    lib = load()
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 8, 1)
    cp.i_representation[:] = 4  # simple urban with one layer is fine ...
    bc, fl = _outputs(cfg, cp.ncol, cp.ntotlay)
    assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl) == 0
    cp2, sw2, lw2 = make_synthetic(cfg, 8, 2)
    cp2.i_representation[:] = 5   # ... with two layers it is the reference's abort (radsurf_interface.F90:281-284)
    bc2, fl2 = _outputs(cfg, cp2.ncol, cp2.ntotlay)
    with pytest.raises(RadsurfError, match="more than one layer"):
        radsurf(cfg, cp2, sw2, lw2, bc2, None, None, *fl2)
    bad = _cfg()
    bad.nswinternal = 3  # spectral resolution mismatch
    with pytest.raises(RadsurfError, match="spectral resolution"):
        radsurf(bad, cp, sw, lw, bc, None, None, *fl)
    assert lib.ssb200_measure_fp64_peak_tflops(1 << 12) > 1.0


def test_full_size_properties():
    """BASELINE size (1,048,576 columns x 16 layers, device-resident): shortwave energy is
    conserved column by column, nothing is non-finite, and a column's result does not depend
    on the columns solved with it (a window of the full solve equals the same columns solved
    alone, bit for bit)."""
    import torch
    lib = load()
    cfg = _cfg()
    ncol, nlay = 1 << 20, 16
    cp, sw, lw = make_synthetic(cfg, ncol, nlay, device="cuda:0")
    bc, fl = _outputs(cfg, ncol, cp.ntotlay, device="cuda:0", profile=False)
    assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl) == 0
    torch.cuda.synchronize()
    for name, f in zip(("sw_norm_dir", "sw_norm_diff", "lw_internal", "lw_norm"), fl):
        for k in ALL_FIELDS:
            a = getattr(f, k)
            if a is not None:
                assert bool(torch.isfinite(a).all()), (name, k)
    for f in fl[:2]:
        res = torch.zeros(ncol, dtype=torch.float64, device="cuda:0")
        s, c = f.as_struct(), cp.as_struct()
        assert lib.ssb200_canopy_flux_check_device(C.byref(s), C.byref(c), C.c_void_p(res.data_ptr()), None) == 0
        torch.cuda.synchronize()
        assert float((res.abs() / f.top_dn[:, 0].abs().clamp_min(1e-30)).max()) < 1e-12
    # the same columns alone: generated with their global column offset, solved as a 4096-column call
    off, n = 777_216, 4096
    cp2, sw2, lw2 = make_synthetic(cfg, n, nlay, col_offset=off, device="cuda:0")
    bc2, fl2 = _outputs(cfg, n, cp2.ntotlay, device="cuda:0", profile=False)
    assert radsurf(cfg, cp2, sw2, lw2, bc2, None, None, *fl2) == 0
    torch.cuda.synchronize()
    for f, g in zip(fl, fl2):
        for k in ALL_FIELDS:
            a, b = getattr(f, k), getattr(g, k)
            if a is None:
                continue
            rows = slice(off, off + n) if a.shape[0] == ncol else slice(off * nlay, (off + n) * nlay)
            assert torch.equal(a[rows], b), k
    lib.ssb200_release()


def test_pruned_interface_state_is_invisible():
    """Without flux profiles (the benchmark configuration) the sweeps store and load only the part of
    the interface state a layer that solves a sub-block of its regions reads
    (ssb_sweep_blocks.cuh: interface_store_pruned): every member other than the profiles equals the
    run with profiles (full interface state) bit for bit, 2 and 4 streams."""
    lib = load()
    for streams in (2, 4):
        cfg = _cfg(streams)
        cp, sw, lw = make_synthetic(cfg, 4096 if streams == 2 else 512, 16)
        outs = []
        for profile in (True, False):
            bc, fl = _outputs(cfg, cp.ncol, cp.ntotlay, profile=profile)
            assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl) == 0
            outs.append(_as_dict(fl, bc))
        assert "flux_dn_layer_top" in outs[0]["lw_norm"] and "flux_dn_layer_top" not in outs[1]["lw_norm"]
        for n, f in outs[1].items():
            for k, v in f.items():
                assert np.array_equal(v, outs[0][n][k]), (streams, n, k, float(np.abs(v - outs[0][n][k]).max()))
    lib.ssb200_release()


def test_pruned_interface_state_mixed_tiles():
    """The same bit-identity on the mixed case (forest and urban tiles with and without vegetation, ragged
    layers, night columns, several spectral intervals; tests/mixed_case.py), 2 and 4 streams."""
    from mixed_case import mixed_config, make_mixed
    for streams in (2, 4):
        cfg = mixed_config(streams).consolidate()
        cp, sw, lw = make_mixed(cfg, ncol=1500)
        outs = []
        for profile in (True, False):
            bc = boundary_conds_out_type().allocate(cp.ncol, cfg.nsw, cfg.nlw)
            fl = [canopy_flux_type().allocate(cfg, cp.ncol, cp.ntotlay, n, use_direct=d, do_save_flux_profile=profile)
                  for n, d in ((cfg.nsw, True), (cfg.nsw, True), (cfg.nlw, False), (cfg.nlw, False))]
            for f in fl:
                f.fill(7.0)
            assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl) == 0
            outs.append(_as_dict(fl, bc))
        for n, f in outs[1].items():
            for k, v in f.items():
                assert np.array_equal(v, outs[0][n][k]), (streams, n, k, float(np.abs(v - outs[0][n][k]).max()))
    load().ssb200_release()


def test_argument_validation_instead_of_device_faults():
    """Unknown tile codes and members a present tile type needs but that are not allocated are argument
    errors with a message, not device faults (round-1 advisor findings); the context stays usable."""
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 64, 4)
    bc, fl = _outputs(cfg, cp.ncol, cp.ntotlay)
    bad = copy_obj(cp)
    bad.i_representation = cp.i_representation.copy()
    bad.i_representation[3] = 9
    with pytest.raises(RadsurfError, match="unknown i_representation"):
        radsurf(cfg, bad, sw, lw, bc, None, None, *fl)
    for obj, member in ((cp, "building_scale"), (cp, "veg_fsd"), (sw, "wall_albedo"), (lw, "ground_emission")):
        broken = copy_obj(obj)
        setattr(broken, member, None)
        args = [broken if o is obj else o for o in (cp, sw, lw)]
        with pytest.raises(RadsurfError, match=member):
            radsurf(cfg, *args, bc, None, None, *fl)
    assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl) == 0  # the same context still solves


def test_calls_from_two_streams_are_ordered():
    """ssb200_radsurf_device shares one context (scratch, plan and status buffers): a call on another
    stream than the previous one waits for it on the device, so back-to-back calls from two streams give
    the results of the same calls made one after the other (round-1 advisor finding)."""
    import torch
    cfg = _cfg()
    probs = []
    for off in (0, 50_000):
        cp, sw, lw = make_synthetic(cfg, 20_000, 16, col_offset=off, device="cuda:0")
        probs.append((cp, sw, lw, _outputs(cfg, cp.ncol, cp.ntotlay, device="cuda:0", profile=False),
                      _outputs(cfg, cp.ncol, cp.ntotlay, device="cuda:0", profile=False)))
    torch.cuda.synchronize()
    for cp, sw, lw, (bc, fl), _ in probs:  # one after the other, default stream
        assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl) == 0
        torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(3):  # back to back, alternating streams, no host synchronisation in between
        for (cp, sw, lw, _, (bc, fl)), st in zip(probs, streams):
            assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl, stream=st.cuda_stream) == 0
    torch.cuda.synchronize()
    for cp, sw, lw, (bc, fl), (bc2, fl2) in probs:
        _same(_as_dict(fl, bc), _as_dict(fl2, bc2))
    load().ssb200_release()


def test_simple_spectrum_lw_on_device():
    """calc_simple_spectrum_lw through the device entry equals the host (numpy) version."""
    import copy
    import torch
    from spartacus_surface_b200.radsurf_simple_spectrum import calc_simple_spectrum_lw
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 500, NLAY)
    rng = np.random.default_rng(3)
    cp.ground_temperature = rng.uniform(260, 310, cp.ncol)
    for k in ("roof_temperature", "wall_temperature", "clear_air_temperature", "veg_temperature",
              "veg_air_temperature"):
        setattr(cp, k, rng.uniform(260, 310, cp.ntotlay))
    fields = ("ground_emission", "roof_emission", "wall_emission", "clear_air_planck", "veg_planck", "veg_air_planck")
    dcp, dlw = _to_device(cp), _to_device(lw)
    for k in fields:
        getattr(lw, k)[...] = -1.0
        getattr(dlw, k).fill_(-1.0)
    calc_simple_spectrum_lw(cfg, cp, lw, 11, 420)
    calc_simple_spectrum_lw(cfg, dcp, dlw, 11, 420)
    torch.cuda.synchronize()
    for k in fields:
        h, d = getattr(lw, k), getattr(dlw, k).cpu().numpy()
        assert np.array_equal(h == -1.0, d == -1.0), k          # same range written
        assert np.allclose(h, d, rtol=1e-15, atol=0), k         # ** 4 (pow) against (t*t)*(t*t)


def test_empty_and_flat_only_inputs():
    """No columns at all, and columns that own no layers (Flat tiles only, ntotlay = 0)."""
    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 16, 2)
    # empty column range
    bc, fl = _outputs(cfg, cp.ncol, cp.ntotlay)
    for f in fl:
        f.fill(7.0)
    assert radsurf(cfg, cp, sw, lw, bc, 5, 4, *fl) == 0
    assert all(np.all(getattr(f, k) == 7.0) for f in fl for k in ALL_FIELDS if getattr(f, k) is not None)
    # Flat tiles only: per-layer arrays are empty
    ncol = 5
    fcp, fsw, flw = make_synthetic(cfg, ncol, 1)
    fcp.set_layers(np.zeros(ncol, dtype=np.int32))
    fcp.i_representation = np.zeros(ncol, dtype=np.int32)
    for obj in (fcp, fsw, flw):
        for k, v in list(vars(obj).items()):
            if isinstance(v, np.ndarray) and v.dtype == np.float64 and v.shape[0] == ncol * 1 and k not in (
                    "cos_sza", "ground_albedo", "ground_albedo_dir", "ground_emissivity", "ground_emission"):
                setattr(obj, k, np.ascontiguousarray(v[:0]))
    fbc, ffl = _outputs(cfg, ncol, 0)
    assert radsurf(cfg, fcp, fsw, flw, fbc, None, None, *ffl) == 0
    obc, ofl = _outputs(cfg, ncol, 0)
    oracle_lib.make_solver()(cfg, fcp, fsw, flw, obc, None, None, *ofl)
    for f, g in zip(ffl, ofl):
        for k in ALL_FIELDS:
            a = getattr(f, k)
            if a is not None:
                assert np.allclose(a, getattr(g, k), rtol=1e-14, atol=1e-300), k
    for k in golden_io.BC_FIELDS:
        assert np.allclose(getattr(fbc, k), getattr(obc, k), rtol=1e-14), k


def test_radsurf_fluxes_equals_radsurf_scale_sum():
    """ssb200_radsurf_fluxes (the driver's calc_simple_spectrum_lw + radsurf + scale + sum in one call,
    normalised flux objects kept on the device, read_input's defaults and the LW emission evaluated on
    the device for members passed as None) against the same steps done one by one: bit-equal."""
    import copy
    from spartacus_surface_b200.radsurf_interface import radsurf_fluxes
    from spartacus_surface_b200.radsurf_lw_spectral_properties import StefanBoltzmann
    cfg = _cfg()
    ncol = 70000  # large enough for the pipelined path (several blocks)
    cp, sw, lw = make_synthetic(cfg, ncol, 4)
    rng = np.random.default_rng(3)
    top_sw = rng.uniform(200.0, 900.0, size=(ncol, 1))
    top_dir = top_sw * rng.uniform(0.2, 0.9, size=(ncol, 1))
    top_lw = rng.uniform(250.0, 400.0, size=(ncol, 1))
    t_ground, t_roof, t_wall, t_air = (rng.uniform(270.0, 300.0, size=s) for s in
                                       ((ncol,), (cp.ntotlay,), (cp.ntotlay,), (cp.ntotlay,)))
    p4 = lambda t: (t * t) * (t * t)
    # step by step, every input explicit (what the reference driver holds in memory)
    lw.ground_emission = (StefanBoltzmann * lw.ground_emissivity[:, 0] * p4(t_ground))[:, None].copy()
    lw.roof_emission = (StefanBoltzmann * lw.roof_emissivity[:, 0] * p4(t_roof))[:, None].copy()
    lw.wall_emission = (StefanBoltzmann * lw.wall_emissivity[:, 0] * p4(t_wall))[:, None].copy()
    for k in ("clear_air_planck", "veg_planck", "veg_air_planck"):
        setattr(lw, k, (StefanBoltzmann * p4(t_air))[:, None].copy())
    bc, fl = _outputs(cfg, ncol, cp.ntotlay)
    assert radsurf(cfg, cp, sw, lw, bc, None, None, *fl) == 0
    fl[0].scale(cp.nlay, top_dir)
    fl[1].scale(cp.nlay, top_sw - top_dir)
    fl[3].scale(cp.nlay, top_lw)
    sw_ref = canopy_flux_type().allocate(cfg, ncol, cp.ntotlay, 1, use_direct=True)
    lw_ref = canopy_flux_type().allocate(cfg, ncol, cp.ntotlay, 1, use_direct=False)
    sw_ref.sum(fl[0], fl[1])
    lw_ref.sum(fl[2], fl[3])
    # one call; everything the driver would default or derive is left to the library
    cp2, sw2, lw2 = copy.copy(cp), copy.copy(sw), copy.copy(lw)
    cp2.veg_contact_fraction = None
    sw2.air_ext = sw2.air_ssa = sw2.wall_specular_frac = sw2.roof_albedo_dir = None
    lw2.air_ext = lw2.air_ssa = None
    for k in ("ground_emission", "roof_emission", "wall_emission", "clear_air_planck", "veg_planck", "veg_air_planck"):
        setattr(lw2, k, None)
    bc2 = boundary_conds_out_type().allocate(ncol, 1, 1)
    sw_flux = canopy_flux_type().allocate(cfg, ncol, cp.ntotlay, 1, use_direct=True)
    lw_flux = canopy_flux_type().allocate(cfg, ncol, cp.ntotlay, 1, use_direct=False)
    assert radsurf_fluxes(cfg, cp2, sw2, lw2, bc2, None, None, sw_flux, lw_flux, top_flux_dn_sw=top_sw,
                          top_flux_dn_direct_sw=top_dir, top_flux_dn_lw=top_lw, ground_temperature=t_ground,
                          roof_temperature=t_roof, wall_temperature=t_wall, clear_air_temperature=t_air,
                          veg_temperature=t_air, veg_air_temperature=t_air) == 0
    for ref, got, name in ((sw_ref, sw_flux, "sw"), (lw_ref, lw_flux, "lw")):
        for k in ALL_FIELDS:
            a, b = getattr(ref, k), getattr(got, k)
            if a is not None:
                assert np.array_equal(a, b), (name, k, float(np.abs(a - b).max()))
    for k in golden_io.BC_FIELDS:
        assert np.array_equal(getattr(bc, k), getattr(bc2, k)), k


def test_single_precision_storage_entry():
    """ssb200_radsurf_sp (float32 arrays of a -DSINGLE_PRECISION build): the arrays
    cross PCIe as float32, the solve runs in FP64, and every output is the FP64 result rounded to nearest
    float32 - i.e. bit-equal to ssb200_radsurf on the widened inputs followed by a cast.  Checked on the
    synthetic canopy (pipelined path: several blocks) and on the mixed-tile case (ragged layers, night
    columns, several spectral intervals, untouched entries)."""
    from mixed_case import mixed_config, make_mixed
    from spartacus_surface_b200.radsurf_interface import radsurf_sp, to_single
    from spartacus_surface_b200.radsurf_canopy_flux import ALL_FIELDS

    def both(cfg, cp, sw, lw, nsw, nlw):
        cp32, sw32, lw32 = to_single(cp), to_single(sw), to_single(lw)
        widen = lambda o: type("W", (), {})
        def wide(o):
            import copy
            w = copy.copy(o)
            for k, v in vars(o).items():
                if isinstance(v, np.ndarray) and v.dtype == np.float32:
                    setattr(w, k, np.ascontiguousarray(v.astype(np.float64)))
            return w
        def outs(single):
            bc = boundary_conds_out_type().allocate(cp.ncol, nsw, nlw)
            fl = [canopy_flux_type().allocate(cfg, cp.ncol, cp.ntotlay, n, use_direct=d)
                  for n, d in ((nsw, True), (nsw, True), (nlw, False), (nlw, False))]
            for o in [bc] + fl:
                for k, v in vars(o).items():
                    if isinstance(v, np.ndarray) and v.dtype == np.float64:
                        v[...] = 7.0
            return (to_single(bc), [to_single(f) for f in fl]) if single else (bc, fl)
        bc32, fl32 = outs(True)
        assert radsurf_sp(cfg, cp32, sw32, lw32, bc32, None, None, *fl32) == 0
        bc64, fl64 = outs(False)
        assert radsurf(cfg, wide(cp32), wide(sw32), wide(lw32), bc64, None, None, *fl64) == 0
        n = 0
        for a, b in zip(fl32 + [bc32], fl64 + [bc64]):
            for k, v in vars(a).items():
                if isinstance(v, np.ndarray) and v.dtype == np.float32:
                    assert np.array_equal(v, getattr(b, k).astype(np.float32)), k
                    n += 1
        assert n > 40

    cfg = _cfg()
    cp, sw, lw = make_synthetic(cfg, 70000, 4)
    both(cfg, cp, sw, lw, 1, 1)
    cfgm = mixed_config(2).consolidate()
    cpm, swm, lwm = make_mixed(cfgm, ncol=400)
    both(cfgm, cpm, swm, lwm, cfgm.nsw, cfgm.nlw)
