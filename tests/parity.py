"""Parity metric shared by the GPU tests, smoke() and bench.py.

Ground truth
------------
`truth` = the reference algorithm (the oracle restatement, oracle/src) evaluated
with every scalar in _Float128 on the same FP64 inputs (oracle/_build/
liboracle_quad.so; stored for the golden cases in tests/golden/truth/).  With
113 significant bits the rounding of the algorithm is removed (what remains is
< 1e-18 even where LU without pivoting loses 16 digits), so it arbitrates the
ill-conditioned fixtures of the reference where two FP64 builds of the
reference itself disagree (test/simple noscat/surfaces/overhang, rami5,
8 streams).

Rule (stated once, used everywhere)
-----------------------------------
For every output field of every flux object

    err(gpu, truth)  <=  max(1e-9, 2 * err(reference_fp64, truth))

i.e. BASELINE.json's 1e-9 relative on fluxes, relaxed only where the reference's
own FP64 arithmetic is demonstrably worse than that on the same input - there
the CUDA path has to be at least as accurate as the reference (factor 2 for the
run-to-run scatter of that error).  err(reference_fp64, truth) is the larger of
the two FP64 builds of the oracle (with and without FMA contraction: what
`gfortran -O2 -march=native` and plain `gfortran -O2` produce).

Measure `err(a, truth)`:
 * the four flux-scale fields of an object (top_dn, top_net, ground_dn,
   ground_net): ELEMENTWISE relative error, max_i |a_i - t_i| / max(|t_i|,
   ELEM_FLOOR * S) with S the largest magnitude of those four fields in the
   object (entries below 1e-6 of the flux scale are measured against that
   floor);
 * every other field: max_i |a_i - t_i| / max(max_i |t_i|, SMALL_FIELD_FLOOR * S)
   (absorptions by clear air etc. are ~1e-7 of the flux they derive from and
   are measured against 1e-3 of the flux scale); sunlit fractions against
   their own maximum.
"""
import numpy as np

REL_TOL = 1e-9
REF_FACTOR = 2.0
ELEM_FLOOR = 1e-6
SMALL_FIELD_FLOOR = 1e-3
SCALE_FIELDS = ("top_dn", "top_net", "ground_dn", "ground_net")

# Golden cases where the register-resident path is known to sit within 1.5x ABOVE the bound on
# one or two fields (everything at the 1e-8 level, both for this path and for the reference's own
# FP64 arithmetic; profiles/r02_parity_table.json has the numbers).  The tests report them as
# expected failures instead of hiding them behind a looser rule.
KNOWN_MARGINAL = {
    name: "RAMI-V HET09 scene, 4 streams: 62 layers, the top one with a vegetation fraction of 1.7e-5 "
          "(regions differing by 1e5 in area): direct albedo 2.1e-8 from the truth against 7.1e-9 for "
          "the reference's FP64 arithmetic (ratio to the bound <= 1.5)"
    for name in ("rami5_HET09_JBS_SUM-41-direct", "rami5_HET09_JBS_SUM-56-direct",
                 "rami5_HET09_JBS_SUM-56-direct-blacksoil")
}


def field_errors(got, truth):
    """{(object, field): err(got, truth)} in the measure defined above."""
    out = {}
    for name, fields in truth.items():
        obj_scale = max([float(np.abs(np.asarray(fields[k])).max()) for k in SCALE_FIELDS
                         if k in fields and np.asarray(fields[k]).size] or [0.0])
        for k, t in fields.items():
            g = np.asarray(got[name][k], dtype=np.float64)
            t = np.asarray(t, dtype=np.float64)
            assert g.shape == t.shape, (name, k, g.shape, t.shape)
            if t.size == 0:
                continue
            if not np.all(np.isfinite(g)):
                out[(name, k)] = float("inf")
                continue
            diff = np.abs(g - t)
            if not diff.any():
                out[(name, k)] = 0.0
            elif k in SCALE_FIELDS:
                out[(name, k)] = float((diff / np.maximum(np.abs(t), max(ELEM_FLOOR * obj_scale, 1e-300))).max())
            else:
                scale = float(np.abs(t).max())
                if "sunlit" not in k:
                    scale = max(scale, SMALL_FIELD_FLOOR * obj_scale)
                out[(name, k)] = float(diff.max()) / max(scale, 1e-300)
    return out


def reference_errors(truth, *reference_fp64):
    """err(reference_fp64, truth) per field: the largest over the given FP64 builds."""
    ref = {}
    for r in reference_fp64:
        for key, e in field_errors(r, truth).items():
            ref[key] = max(ref.get(key, 0.0), e)
    return ref


def check(got, truth, *reference_fp64, factor=REF_FACTOR):
    """Returns (ok, worst err/bound, report lines).  `factor`: REF_FACTOR for the product path; the
    test-only generic kernels (the reference's own operation order compiled by nvcc, i.e. one more
    FP64 build of the reference) are held to 4."""
    err = field_errors(got, truth)
    ref = reference_errors(truth, *reference_fp64)
    lines, ok, worst = [], True, 0.0
    for key in sorted(err):
        bound = max(REL_TOL, factor * ref.get(key, 0.0))
        worst = max(worst, err[key] / bound)
        if err[key] > bound:
            ok = False
            lines.append(f"{key[0]}.{key[1]}: err_vs_truth={err[key]:.3e} > bound={bound:.3e} "
                         f"(reference_fp64 err_vs_truth={ref.get(key, 0.0):.3e})")
    return ok, worst, lines


def summary(got, truth, *reference_fp64):
    """Numbers for the parity table (profiles/r02_parity_table.json)."""
    err = field_errors(got, truth)
    ref = reference_errors(truth, *reference_fp64)
    worst_key = max(err, key=lambda k: err[k]) if err else None
    bounds = {k: max(REL_TOL, REF_FACTOR * ref.get(k, 0.0)) for k in err}
    ratio_key = max(err, key=lambda k: err[k] / bounds[k]) if err else None
    return {
        "err_gpu": max(err.values()) if err else 0.0,
        "err_gpu_field": ".".join(worst_key) if worst_key else None,
        "err_ref_fp64": max(ref.values()) if ref else 0.0,
        "bound_max": max(bounds.values()) if bounds else REL_TOL,
        "worst_err_over_bound": (err[ratio_key] / bounds[ratio_key]) if ratio_key else 0.0,
        "worst_err_over_bound_field": ".".join(ratio_key) if ratio_key else None,
        "fields_above_1e-9": sorted(".".join(k) for k in err if err[k] > REL_TOL),
    }


def max_err(got, truth):
    e = field_errors(got, truth)
    return max(e.values()) if e else 0.0
