"""(De)serialisation of solver cases used as committed golden fixtures.

A case file holds the effective inputs of one `radsurf` call (everything
read_input + the namelist overrides produce, after the LW emission
preparation) and the outputs of the CPU oracle for them.  Files are written
by tests/golden/make_golden.py in the build container (where /root/reference
exists) and only READ everywhere else.
"""
import json
import os

import numpy as np

from spartacus_surface_b200 import (config_type, canopy_properties_type, sw_spectral_properties_type,
                                    lw_spectral_properties_type, canopy_flux_type, boundary_conds_out_type)
from spartacus_surface_b200.radsurf_canopy_flux import ALL_FIELDS
from spartacus_surface_b200.driver.spartacus_surface_driver import DriverResult, allocate_outputs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLUX_NAMES = ("sw_norm_dir", "sw_norm_diff", "lw_internal", "lw_norm")
BC_FIELDS = ("sw_albedo", "sw_albedo_dir", "lw_emissivity", "lw_emission")
CONFIG_KEYS = ("do_sw", "do_lw", "use_sw_direct_albedo", "do_vegetation", "do_urban",
               "n_vegetation_region_forest", "n_vegetation_region_urban", "nsw", "nlw",
               "n_stream_sw_forest", "n_stream_sw_urban", "n_stream_lw_forest", "n_stream_lw_urban",
               "use_symmetric_vegetation_scale_forest", "use_symmetric_vegetation_scale_urban",
               "vegetation_isolation_factor_forest", "vegetation_isolation_factor_urban",
               "min_vegetation_fraction", "min_building_fraction", "do_save_flux_profile", "iverbose")


def _dump_obj(out, prefix, obj):
    for k, v in vars(obj).items():
        if isinstance(v, np.ndarray):
            out[f"{prefix}.{k}"] = v


def _load_obj(npz, prefix, obj):
    for key in npz.files:
        if key.startswith(prefix + "."):
            setattr(obj, key[len(prefix) + 1:], np.ascontiguousarray(npz[key]))


def save_case(path, r, meta=None):
    out = {}
    cfg = {k: getattr(r.config, k) for k in CONFIG_KEYS}
    out["config_json"] = np.frombuffer(json.dumps({"config": cfg, "meta": meta or {}}).encode(), dtype=np.uint8)
    _dump_obj(out, "canopy", r.canopy_props)
    _dump_obj(out, "sw", r.sw_spectral_props)
    _dump_obj(out, "lw", r.lw_spectral_props)
    for name in ("top_flux_dn_sw", "top_flux_dn_direct_sw", "top_flux_dn_lw"):
        v = getattr(r, name, None)
        if v is not None:
            out["top." + name] = v
    for name in FLUX_NAMES:
        f = getattr(r, name)
        if f is not None:
            _dump_obj(out, "out." + name, f)
    _dump_obj(out, "out.bc", r.bc_out)
    np.savez_compressed(path, **out)


def load_case(path, legendre_gauss_init=None):
    """Return (r, expected): r has inputs and freshly allocated (zeroed) outputs,
    expected = {"sw_norm_dir": {field: array}, ..., "bc": {...}}."""
    if not os.path.isabs(path):
        path = os.path.join(GOLDEN_DIR, path)
    npz = np.load(path)
    info = json.loads(bytes(npz["config_json"]).decode())
    config = config_type()
    for k, v in info["config"].items():
        setattr(config, k, v)
    config.consolidate(legendre_gauss_init)
    r = DriverResult()
    r.meta = info["meta"]
    r.config = config
    r.canopy_props = canopy_properties_type()
    _load_obj(npz, "canopy", r.canopy_props)
    cp = r.canopy_props
    cp.ncol, cp.ntotlay = int(cp.nlay.size), int(cp.nlay.sum())
    r.sw_spectral_props = sw_spectral_properties_type(config.nsw)
    _load_obj(npz, "sw", r.sw_spectral_props)
    r.lw_spectral_props = lw_spectral_properties_type(config.nlw)
    _load_obj(npz, "lw", r.lw_spectral_props)
    for name in ("top_flux_dn_sw", "top_flux_dn_direct_sw", "top_flux_dn_lw"):
        setattr(r, name, npz["top." + name] if "top." + name in npz.files else None)
    allocate_outputs(r)
    expected = {}
    for name in FLUX_NAMES + ("bc",):
        d = {k[len("out." + name) + 1:]: npz[k] for k in npz.files if k.startswith("out." + name + ".")}
        if d:
            expected[name] = d
    return r, expected


TRUTH_DIR = os.path.join(GOLDEN_DIR, "truth")


def load_truth(case):
    """Ground truth of a golden case (tests/golden/make_truth.py: the oracle's
    _Float128 build on the stored inputs), same nested-dict form as `expected`."""
    npz = np.load(os.path.join(TRUTH_DIR, case if case.endswith(".npz") else case + ".npz"))
    truth = {}
    for key in npz.files:
        name, k = key.split(".", 1)
        truth.setdefault(name, {})[k] = npz[key]
    return truth


def outputs_of(r):
    """Current outputs of r in the same nested-dict form as `expected`."""
    got = {}
    for name in FLUX_NAMES:
        f = getattr(r, name)
        if f is not None:
            got[name] = {k: getattr(f, k) for k in ALL_FIELDS if getattr(f, k) is not None}
    got["bc"] = {k: getattr(r.bc_out, k) for k in BC_FIELDS if getattr(r.bc_out, k) is not None}
    return got


def max_rel_err(got, expected, atol=1e-12):
    """Largest |got-exp| / max(|exp|, scale floor) over all fields; returns (err, where)."""
    worst, where = 0.0, None
    for name, fields in expected.items():
        for k, e in fields.items():
            g = np.asarray(got[name][k])
            e = np.asarray(e)
            assert g.shape == e.shape, (name, k, g.shape, e.shape)
            if not np.all(np.isfinite(g)):
                return float("inf"), (name, k, "non-finite")
            err = np.abs(g - e) / np.maximum(np.abs(e), atol)
            mask = np.abs(g - e) <= atol
            err = np.where(mask, 0.0, err)
            if err.size and err.max() > worst:
                worst, where = float(err.max()), (name, k, int(np.argmax(err)))
    return worst, where


def list_cases():
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))
