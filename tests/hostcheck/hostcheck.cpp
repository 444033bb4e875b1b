// hostcheck.cpp - TEST INFRASTRUCTURE: compiles the product's solver bodies
// (spartacus_surface_b200/csrc/*.cuh, the same code the CUDA kernels wrap) for
// the host and runs them with a serial loop over "threads".  It lets the CPU
// test-suite check the kernels' arithmetic, scratch indexing and launch plan
// against the oracle without a GPU.  It is NOT a product path: nothing in the
// package loads it, and the library itself has no CPU fallback.
#include <vector>

#define SSB_HOSTCHECK_QR 1
#include "asymtx_qr.hpp"
#include "../../spartacus_surface_b200/csrc/ssb_driver.hpp"
#include "../../spartacus_surface_b200/csrc/ssb_fast_layer.cuh"
#include "../../spartacus_surface_b200/csrc/ssb_fast_sweeps.cuh"
#include "../../spartacus_surface_b200/csrc/ssb_fused.cuh"

namespace {

struct HostBackend {
  const ssb::Plan *plan = nullptr;
  std::vector<double> buf;
  int status = 0;
  bool fast = false;  // use the register-resident bodies where they exist (ns <= 4)
  bool fused = false;  // ... and the column-resident ones (ssb_fused.cuh; ns <= 2)
  template <int NSA, int NREG, bool URBAN>
  bool fused_shape_t(bool lw, int *pe, int *oe, int *geo) {
    if (lw) {
      *pe = ssb::LwFused<NREG, NSA, URBAN>::private_elems;
      *oe = ssb::LwFused<NREG, NSA, URBAN>::op_elems;
      *geo = ssb::LwSweepLayout<NREG, NSA, URBAN>::oGeo;
    } else {
      *pe = ssb::SwFused<NREG, NSA, URBAN>::private_elems;
      *oe = ssb::SwFused<NREG, NSA, URBAN>::op_elems;
      *geo = ssb::SwSweepLayout<NREG, NSA, URBAN>::oGeo;
    }
    return true;
  }
  template <int NSA>
  bool fused_shape_ns(const ssb::SolveCfg &c, bool lw, int *pe, int *oe, int *geo) {
    switch (c.nreg * 2 + (c.urban ? 1 : 0)) {
      case 2: return fused_shape_t<NSA, 1, false>(lw, pe, oe, geo);
      case 3: return fused_shape_t<NSA, 1, true>(lw, pe, oe, geo);
      case 4: return fused_shape_t<NSA, 2, false>(lw, pe, oe, geo);
      case 5: return fused_shape_t<NSA, 2, true>(lw, pe, oe, geo);
      case 6: return fused_shape_t<NSA, 3, false>(lw, pe, oe, geo);
      case 7: return fused_shape_t<NSA, 3, true>(lw, pe, oe, geo);
      default: return false;
    }
  }
  bool fused_shape(const ssb::SolveCfg &c, bool lw, int *pe, int *oe, int *geo) {
    if (!fused) return false;
    if (c.ns == 1) return fused_shape_ns<1>(c, lw, pe, oe, geo);
    if (c.ns == 2) return fused_shape_ns<2>(c, lw, pe, oe, geo);
    return false;
  }
  int fused_slots() { return 1; }
  int fused_flags() { return 1; }
  template <int NSA, int NREG, bool URBAN>
  void fused_cols(const ssb::ClassArgs &a, bool lw, long width) {
    double stack[512];
    const ssb::StateMem st{stack, 1};
    for (long t = 0; t < width; ++t) {
      if (lw)
        ssb::fused_column_lw<NREG, NSA, URBAN>(a, (int)t, true, st);
      else
        ssb::fused_column_sw<NREG, NSA, URBAN>(a, (int)t, true, st);
    }
  }
  template <int NSA>
  void fused_run_ns(const ssb::ClassArgs &a, bool lw, long width) {
    switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
      case 2: fused_cols<NSA, 1, false>(a, lw, width); break;
      case 3: fused_cols<NSA, 1, true>(a, lw, width); break;
      case 4: fused_cols<NSA, 2, false>(a, lw, width); break;
      case 5: fused_cols<NSA, 2, true>(a, lw, width); break;
      case 6: fused_cols<NSA, 3, false>(a, lw, width); break;
      case 7: fused_cols<NSA, 3, true>(a, lw, width); break;
      default: break;
    }
  }
  void fused_run(const ssb::ClassArgs &a, bool lw, long width) {
    if (a.cfg.ns == 1) fused_run_ns<1>(a, lw, width);
    else fused_run_ns<2>(a, lw, width);
  }
  void fork_passes() {}
  void begin_pass(bool) {}
  void end_passes() {}
  // level-major staging of the per-layer arrays (ssb_stage.cuh), with the register-resident bodies
  bool stage_layers = true;
  bool stage_supported(const ssb::SolveCfg &c) { return fast && stage_layers && c.ns <= 4; }
  void stage(const ssb::StageArgs &s, bool scatter, bool) { ssb::stage_host(s, scatter); }
  // record sweeps after the register-resident layer bodies (mode 3)
  bool records = false;
  bool records_shape(const ssb::SolveCfg &c, bool lw, int *oe) {
    if (!records || !fast || c.ns > 2) return false;
    int pe = 0, geo = 0;
    return c.ns == 1 ? fused_shape_ns<1>(c, lw, &pe, oe, &geo) : fused_shape_ns<2>(c, lw, &pe, oe, &geo);
  }
  template <int NSA, int NREG, bool URBAN>
  void record_cols(const ssb::ClassArgs &a, bool lw, long width) {
    double stack[512];
    const ssb::StateMem st{stack, 1};
    for (long t = 0; t < width; ++t) {
      if (lw)
        ssb::fused_column_lw<NREG, NSA, URBAN, 1>(a, (int)t, true, st);
      else
        ssb::fused_column_sw<NREG, NSA, URBAN, 1>(a, (int)t, true, st);
    }
    for (long t = 0; t < width; ++t) {
      if (lw)
        ssb::fused_column_lw<NREG, NSA, URBAN, 2>(a, (int)t, true, ssb::StateMem{nullptr, 0});
      else
        ssb::fused_column_sw<NREG, NSA, URBAN, 2>(a, (int)t, true, ssb::StateMem{nullptr, 0});
    }
  }
  template <int NSA>
  void records_run_ns(const ssb::ClassArgs &a, bool lw, long width) {
    switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
      case 2: record_cols<NSA, 1, false>(a, lw, width); break;
      case 3: record_cols<NSA, 1, true>(a, lw, width); break;
      case 4: record_cols<NSA, 2, false>(a, lw, width); break;
      case 5: record_cols<NSA, 2, true>(a, lw, width); break;
      case 6: record_cols<NSA, 3, false>(a, lw, width); break;
      case 7: record_cols<NSA, 3, true>(a, lw, width); break;
      default: break;
    }
  }
  void records_run(const ssb::ClassArgs &a, bool lw, long width) {
    if (a.cfg.ns == 1) records_run_ns<1>(a, lw, width);
    else records_run_ns<2>(a, lw, width);
  }
  size_t budget = (size_t)1 << 22;  // small on purpose: exercises the chunk loop
  const int *dev_cols(const ssb::Plan &p, size_t off) { return p.all_cols.data() + off; }
  // reversed column order inside every chunk: results must not depend on it
  std::vector<int> ordered;
  const int *order_chunk(const ssb::ClassArgs &a, const int *host_cols) {
    ordered.assign(host_cols, host_cols + a.ncols);
    std::reverse(ordered.begin(), ordered.end());
    return ordered.data();
  }
  const int *dev_nlay() { return plan->nlay.data(); }
  const int *dev_istartlay() { return plan->istartlay.data(); }
  const int *dev_irep() { return plan->irep.data(); }
  double *scratch(size_t n) {
    if (buf.size() < n) buf.resize(n);
    // poison so that reads of never-written scratch show up as NaN
    for (size_t i = 0; i < n; ++i) buf[i] = __builtin_nan("");
    return buf.data();
  }
  size_t scratch_budget_doubles() { return budget; }
  int *dev_status() { return &status; }
  // register-resident bodies for the actual stream count NSA (1..4), as the device runs them:
  // prepare pass (k_partition_layers), then every layer problem
  template <int NSA, bool LW>
  void fast_layers(const ssb::ClassArgs &a, long nt) {
    const long width = (long)a.ncols * a.cfg.nspec;
    double stack[256];
    const ssb::StateMem st{stack, 1};
    for (long t = 0; t < nt; ++t) ssb::fast_prepare_level(a, (int)(t % width), (int)(t / width));
    for (long t = 0; t < nt; ++t) {
      const int q = (int)(t % width), lev = (int)(t / width);
      if (LW) {
        if (a.cfg.nreg == 1) ssb::fast_layer_problem_lw<1, NSA>(a, q, lev, st);
        else if (a.cfg.nreg == 2) ssb::fast_layer_problem_lw<2, NSA>(a, q, lev, st);
        else ssb::fast_layer_problem_lw<3, NSA>(a, q, lev, st);
      } else {
        if (a.cfg.nreg == 1) ssb::fast_layer_problem_sw<1, NSA>(a, q, lev, st);
        else if (a.cfg.nreg == 2) ssb::fast_layer_problem_sw<2, NSA>(a, q, lev, st);
        else ssb::fast_layer_problem_sw<3, NSA>(a, q, lev, st);
      }
    }
  }
  template <bool LW>
  bool try_fast_layers(const ssb::ClassArgs &a, long nt) {
    if (!fast) return false;
    switch (a.cfg.ns) {
      case 1: fast_layers<1, LW>(a, nt); return true;
      case 2: fast_layers<2, LW>(a, nt); return true;
      case 3: fast_layers<3, LW>(a, nt); return true;
      case 4: fast_layers<4, LW>(a, nt); return true;
      default: return false;
    }
  }
  template <int NS>
  void layer_sw(const ssb::ClassArgs &a, long nt) {
    if (try_fast_layers<false>(a, nt)) return;
    const long width = (long)a.ncols * a.cfg.nspec;
    for (long t = 0; t < nt; ++t) ssb::layer_problem_sw<NS>(a, (int)(t % width), (int)(t / width));
  }
  template <int NS>
  void layer_lw(const ssb::ClassArgs &a, long nt) {
    if (try_fast_layers<true>(a, nt)) return;
    const long width = (long)a.ncols * a.cfg.nspec;
    for (long t = 0; t < nt; ++t) ssb::layer_problem_lw<NS>(a, (int)(t % width), (int)(t / width));
  }
  template <int NSA, bool LW, int NREG, bool URBAN>
  void fast_sweeps(const ssb::ClassArgs &a, long nt) {
    double state[256];
    const ssb::StateMem st{state, 1};
    for (long t = 0; t < nt; ++t) {
      if (LW)
        ssb::fast_column_sweeps_lw<NREG, NSA, URBAN>(a, (int)t, st);
      else
        ssb::fast_column_sweeps_sw<NREG, NSA, URBAN>(a, (int)t, st);
    }
  }
  template <int NSA, bool LW>
  bool fast_sweeps_ns(const ssb::ClassArgs &a, long nt) {
    switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
      case 2: fast_sweeps<NSA, LW, 1, false>(a, nt); return true;
      case 3: fast_sweeps<NSA, LW, 1, true>(a, nt); return true;
      case 4: fast_sweeps<NSA, LW, 2, false>(a, nt); return true;
      case 5: fast_sweeps<NSA, LW, 2, true>(a, nt); return true;
      case 6: fast_sweeps<NSA, LW, 3, false>(a, nt); return true;
      case 7: fast_sweeps<NSA, LW, 3, true>(a, nt); return true;
      default: return false;
    }
  }
  template <bool LW>
  bool try_fast_sweeps(const ssb::ClassArgs &a, long nt) {
    if (!fast) return false;
    switch (a.cfg.ns) {
      case 1: return fast_sweeps_ns<1, LW>(a, nt);
      case 2: return fast_sweeps_ns<2, LW>(a, nt);
      case 3: return fast_sweeps_ns<3, LW>(a, nt);
      case 4: return fast_sweeps_ns<4, LW>(a, nt);
      default: return false;
    }
  }
  template <int NS>
  void sweeps_sw(const ssb::ClassArgs &a, long nt) {
    if (try_fast_sweeps<false>(a, nt)) return;
    for (long t = 0; t < nt; ++t) ssb::column_sweeps_sw<NS>(a, (int)t);
  }
  template <int NS>
  void sweeps_lw(const ssb::ClassArgs &a, long nt) {
    if (try_fast_sweeps<true>(a, nt)) return;
    for (long t = 0; t < nt; ++t) ssb::column_sweeps_lw<NS>(a, (int)t);
  }
  void surface(const ssb::SurfaceArgs &s, int nsw_threads, int nlw_threads) {
    for (int t = 0; t < nsw_threads; ++t) ssb::surface_column_sw(s, t / s.nsw, t % s.nsw);
    for (int t = 0; t < nlw_threads; ++t) ssb::surface_column_lw(s, t / s.nlw, t % s.nlw);
  }
};

}  // namespace

extern "C" int hostcheck_radsurf(const ssb200_config *config, const ssb200_canopy_properties *cp,
                                 const ssb200_sw_spectral_properties *sw, const ssb200_lw_spectral_properties *lw,
                                 ssb200_boundary_conds_out *bc, int32_t istartcol, int32_t iendcol,
                                 ssb200_canopy_flux *sw_dir, ssb200_canopy_flux *sw_diff,
                                 ssb200_canopy_flux *lw_int, ssb200_canopy_flux *lw_norm, int64_t budget_doubles,
                                 int32_t fast) {
  ssb::CallArgs ca{config, cp, sw, lw, bc, sw_dir, sw_diff, lw_int, lw_norm};
  std::string err;
  int rc = ssb::validate_call(ca, err);
  if (rc) {
    std::fprintf(stderr, "hostcheck: %s\n", err.c_str());
    return rc;
  }
  int c1 = istartcol > 0 ? istartcol : 1, c2 = iendcol > 0 ? iendcol : cp->ncol;
  if (c2 > cp->ncol) c2 = cp->ncol;
  ssb::Plan plan;
  rc = ssb::build_plan(*config, *cp, c1 - 1, c2 - 1, plan, err);
  if (rc) {
    std::fprintf(stderr, "hostcheck: %s\n", err.c_str());
    return rc;
  }
  HostBackend be;
  be.plan = &plan;
  if (budget_doubles > 0) be.budget = (size_t)budget_doubles;
  be.stage_layers = (fast & 8) == 0;  // bit 3: the caller's arrays in place (no level-major staging)
  fast &= 7;
  be.fast = fast != 0;
  // fast == 4: generic bodies with the symmetrised Jacobi eigen-systems, as on the device
  ssb::host_generic_jacobi() = (fast == 4);
  if (fast == 4) be.fast = false;
  be.fused = fast == 2;
  be.records = fast == 3;
  ssb::Dispatcher<HostBackend> disp(be);
  rc = disp.run(ca, plan, err);
  if (rc) {
    std::fprintf(stderr, "hostcheck: %s\n", err.c_str());
    return rc;
  }
  return be.status;
}

// Unit entry: the register-resident layer formulation (csrc/ssb_layer_math.cuh) for ONE layer
// of `nreg` regions x NS streams described by scalars (the Gamma matrices are never formed in
// the product: LayerCoef evaluates their entries).  Outputs are column-major, full size.
//   lw != 0: R, T, int_flux (n x n), source, int_flux_source (n)         <- emission rate b (n)
//   lw == 0: R, T, int_diff (n x n), S_up, S_dn, int_dir_diff (n x d), E, int_dir (d x d)
namespace {
template <int NREG, int NS>
int fast_layer_unit(int lw, const ssb200_legendre_gauss *lgc, const double *ext, const double *ssa,
                    const double *frac, const double *fex, const double *fwall, double wall_ext,
                    double wall_factor, double cos_sza, double dz, const double *brate, double *out) {
  constexpr int n = NREG * NS, d = NREG;
  ssb::LgTable lg;
  lg.ns = NS;
  for (int i = 0; i < NS; ++i) {
    lg.mu[i] = lgc->mu[i];
    lg.tan_ang[i] = lgc->tan_ang[i];
    lg.weight[i] = lgc->weight[i];
    lg.hweight[i] = lgc->hweight[i];
    lg.vweight[i] = lgc->vweight[i];
  }
  lg.vadjustment = lgc->vadjustment;
  lg.vadjustment2 = lgc->vadjustment2;
  ssb::LayerCoef<NREG, NS> k;
  k.set_streams(&lg);
  for (int rf = 0; rf < NREG; ++rf) {
    double loss = 0.0;
    for (int rt = 0; rt < NREG; ++rt) {
      if (rt == rf) continue;
      loss += fex[rt + 3 * rf];
      k.dx[rt + NREG * rf] = fex[rt + 3 * rf];
    }
    k.loss[rf] = loss;
    k.ext[rf] = ext[rf];
    k.es[rf] = ext[rf] * ssa[rf];
    k.fw[rf] = fwall[rf];
    k.frac[rf] = (NREG == 1) ? 1.0 : frac[rf];
    k.rfrac[rf] = (NREG == 1) ? 1.0 : 1.0 / frac[rf];
  }
  k.wall_ext = wall_ext;
  k.wall_factor = wall_factor;
  const double zc = cos_sza > 1.0e-6 ? cos_sza : 1.0e-6;
  k.sin0 = lw ? 0.0 : std::sqrt(1.0 - zc * zc);
  k.tan0 = lw ? 0.0 : k.sin0 / zc;
  k.rcos = lw ? 0.0 : 1.0 / zc;
  const int ne = lw ? 3 * n * n + 2 * n : 3 * n * n + 3 * n * d + 2 * d * d;
  std::vector<double> P((size_t)ne * ssb::kScratchTile, __builtin_nan(""));
  double stack[512];
  const ssb::StateMem st{stack, 1};
  bool ok;
  if (lw)
    ok = ssb::layer_lw_solve<NREG, NS, NREG, 0>(k, brate, dz, P.data(), st);
  else
    ok = ssb::layer_sw_solve<NREG, NS, NREG, 0>(k, dz, P.data(), st);
  for (int e = 0; e < ne; ++e) out[e] = P[(size_t)e * ssb::kScratchTile];
  return ok ? 0 : 1;
}
}  // namespace

extern "C" int hostcheck_fast_layer(int32_t lw, int32_t nreg, const ssb200_legendre_gauss *lg, const double *ext,
                                    const double *ssa, const double *frac, const double *fex, const double *fwall,
                                    double wall_ext, double wall_factor, double cos_sza, double dz,
                                    const double *brate, double *out) {
#define SSB_UNIT(NR, NSV) \
  if (nreg == NR && lg->nstream == NSV) \
    return fast_layer_unit<NR, NSV>(lw, lg, ext, ssa, frac, fex, fwall, wall_ext, wall_factor, cos_sza, dz, brate, out);
  SSB_UNIT(1, 1) SSB_UNIT(1, 2) SSB_UNIT(1, 4) SSB_UNIT(2, 1) SSB_UNIT(2, 2) SSB_UNIT(2, 4)
  SSB_UNIT(3, 1) SSB_UNIT(3, 2) SSB_UNIT(3, 3) SSB_UNIT(3, 4)
#undef SSB_UNIT
  return -1;
}
