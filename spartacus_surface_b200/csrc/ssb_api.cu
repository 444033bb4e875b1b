// ssb_api.cu - kernels and C-ABI entry points of libspartacus_b200.so
// (include/spartacus_b200.h).  sm_100a only; there is no CPU fallback: every
// solve entry returns SSB200_ERR_NOGPU when no CUDA device is present.
#include <cuda_runtime.h>

#include <cub/cub.cuh>
#include <mutex>

#include "ssb_driver.hpp"
#include "ssb_fast.cuh"
#include "ssb_fast_layer.cuh"
#include "ssb_launch.hpp"

namespace ssb {
cudaError_t g_fast_error = cudaSuccess;
}

namespace {

thread_local std::string g_last_error;
std::mutex g_mutex;
long long g_launches = 0;

int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

#define SSB_CUDA(call)                                                                           \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(SSB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// Flat tiles and single-layer urban models (the generic kernels live in ssb_k_ns*_*.cu)
__global__ void k_surface(ssb::SurfaceArgs s, int nsw_threads, int nlw_threads) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nsw_threads) ssb::surface_column_sw(s, t / s.nsw, t % s.nsw);
  if (t < nlw_threads) ssb::surface_column_lw(s, t / s.nlw, t % s.nlw);
}

// sort keys of the columns of one launch chunk (column_segment_key)
// (columns are ordered inside groups of `group` neighbours only: a warp then still touches
// neighbouring lines of the per-column input and output arrays)
// The key is packed into as few bits as it needs - `lbits` for the segment pattern (one bit per
// layer; night-time columns get the largest value) below the group index - so that the radix sort
// runs 3 passes instead of 8: at 131,072 columns per GPU (8-GPU sharding) the sort is latency, not
// bandwidth.
__global__ void k_column_keys(ssb::ClassArgs a, unsigned long long *keys, int group, int lbits) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= a.ncols) return;
  const unsigned long long all = (1ull << lbits) - 1ull;
  const unsigned pattern = ssb::column_segment_key(a, a.cols[ic]);
  const unsigned long long k = (pattern == 0xffffffffu) ? all : ((unsigned long long)pattern & (all >> 1));
  keys[ic] = ((unsigned long long)(ic / group) << lbits) | k;
}

// ---------------------------------------------------------------------------
// canopy_flux_type%scale / %sum / %check on device
// ---------------------------------------------------------------------------
struct FieldList {
  double *p[32];
  const double *a[32], *b[32];
  int per_layer[32];  // 1: (nspec, ntotlay), 0: (nspec, ncol)
  int spectral[32];
  int n;
};

__global__ void k_scale(FieldList fl, const double *factor, const int *lay2col, int nspec, long ncol, long ntotlay) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  for (int f = 0; f < fl.n; ++f) {
    const long cnt = (fl.per_layer[f] ? ntotlay : ncol) * nspec;
    if (t < cnt) {
      const long idx = t / nspec;
      const int g = (int)(t % nspec);
      const long col = fl.per_layer[f] ? lay2col[idx] : idx;
      fl.p[f][t] = factor[g + (long)nspec * col] * fl.p[f][t];
    }
  }
}

__global__ void k_sum(FieldList fl, int nspec, long ncol, long ntotlay) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  for (int f = 0; f < fl.n; ++f) {
    const long cnt = (fl.per_layer[f] ? ntotlay : ncol) * (fl.spectral[f] ? nspec : 1);
    if (t < cnt) fl.p[f][t] = fl.a[f][t] + fl.b[f][t];
  }
}

__global__ void k_check(ssb200_canopy_flux f, const int *nlay, const int *istartlay, const int *irep, int ncol,
                        double *residual) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const int ns = f.nspec;
  const int l1 = istartlay[col] - 1, nl = nlay[col], rep = irep[col];
  auto sum_col = [&](const double *p) {
    double s = 0.0;
    for (int g = 0; g < ns; ++g) s += p[g + (size_t)ns * col];
    return s;
  };
  auto sum_lay = [&](const double *p) {
    double s = 0.0;
    if (p)
      for (int l = 0; l < nl; ++l)
        for (int g = 0; g < ns; ++g) s += p[g + (size_t)ns * (l1 + l)];
    return s;
  };
  const double ground_net = sum_col(f.ground_net), top_net = sum_col(f.top_net);
  const double clear = (rep != SSB200_TILE_FLAT) ? sum_lay(f.clear_air_abs) : 0.0;
  double roof = 0.0, wall = 0.0, veg = 0.0, vegair = 0.0;
  if (rep == SSB200_TILE_URBAN || rep == SSB200_TILE_VEGETATED_URBAN || rep == SSB200_TILE_SIMPLE_URBAN ||
      rep == SSB200_TILE_INFINITE_STREET) {
    roof = sum_lay(f.roof_net);
    wall = sum_lay(f.wall_net);
  }
  if (rep == SSB200_TILE_FOREST || rep == SSB200_TILE_VEGETATED_URBAN) {
    veg = sum_lay(f.veg_abs);
    vegair = sum_lay(f.veg_air_abs);
  }
  residual[col] = ground_net + clear + wall + roof + veg + vegair - top_net;
}

// sigma * emissivity * T^4 (emissivity NULL: 1) for elements [i0, i1) of (nspec, .) arrays, interval 1
__global__ void k_sigma_t4(double *out, const double *emissivity, const double *temperature, int nspec, long i0,
                           long i1) {
  const long i = i0 + blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= i1) return;
  const double t = temperature[i], t2 = t * t;
  const double sb = 5.67037321e-8;  // radtool/radiation_constants.F90:26
  out[(size_t)nspec * i] = emissivity ? (sb * emissivity[(size_t)nspec * i]) * (t2 * t2) : sb * (t2 * t2);
}

// ---- single-precision storage variant: float <-> double on the device (ssb200_radsurf_sp) ----
__global__ void k_f2d(double *dst, const float *src, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}
__global__ void k_d2f(float *dst, const double *src, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];  // round to nearest even
}

// ---- stages around radsurf for ssb200_radsurf_fluxes ----------------------------------------
__global__ void k_fill(double *p, double v, long i0, long i1) {
  const long i = i0 + blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < i1) p[i] = v;
}
// read_input's default veg_contact_fraction (driver/spartacus_surface_read_input.F90:155-166)
__global__ void k_vcf_default(double *vcf, const double *vf, const double *bf, double min_veg, long i0, long i1) {
  const long i = i0 + blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < i1) vcf[i] = fmin(1.0, vf[i] / fmax(min_veg, 1.0 - bf[i]));
}
__global__ void k_lay2col(const int *nlay, const int *istartlay, const int *irep, int ncol, int *lay2col) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncol || irep[j] == SSB200_TILE_FLAT) return;
  const int l0 = istartlay[j] - 1;
  for (int l = 0; l < nlay[j]; ++l) lay2col[l0 + l] = j;
}
// out = a * fa + b * fb per (interval, column) factor, every product and the sum rounded separately
// so that the result equals canopy_flux_type%scale followed by %sum bit for bit
// (radsurf/radsurf_canopy_flux.F90:212-282,399-460).  fa NULL: 1.  fb_minus non-NULL: fb - fb_minus.
// Non-spectral members (sunlit fractions) are summed unscaled.  Window: columns [c0, c1), layers [l0, l1).
__global__ void k_scale_sum(FieldList fl, const double *fa, const double *fb, const double *fb_minus,
                            const int *lay2col, int nspec, long c0, long c1, long l0, long l1) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  for (int f = 0; f < fl.n; ++f) {
    const int w = fl.spectral[f] ? nspec : 1;
    const long lo = (fl.per_layer[f] ? l0 : c0) * w, hi = (fl.per_layer[f] ? l1 : c1) * w;
    const long i = lo + t;
    if (i >= hi) continue;
    const double a = fl.a[f][i], b = fl.b[f][i];
    if (!fl.spectral[f]) {
      fl.p[f][i] = __dadd_rn(a, b);
      continue;
    }
    const long idx = i / nspec;
    const int g = (int)(i % nspec);
    const long col = fl.per_layer[f] ? lay2col[idx] : idx;
    const size_t k = (size_t)g + (size_t)nspec * (size_t)col;
    const double xa = fa ? __dmul_rn(fa[k], a) : a;
    const double fbk = fb_minus ? __dadd_rn(fb[k], -fb_minus[k]) : fb[k];
    fl.p[f][i] = __dadd_rn(xa, __dmul_rn(fbk, b));
  }
}

// register-resident independent DFMA chains: FP64 roofline denominator
__global__ void k_fp64_peak(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c);
    a1 = fma(a1, m, c);
    a2 = fma(a2, m, c);
    a3 = fma(a3, m, c);
    a4 = fma(a4, m, c);
    a5 = fma(a5, m, c);
    a6 = fma(a6, m, c);
    a7 = fma(a7, m, c);
  }
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------
// Context: cached device buffers, plan, scratch
// ---------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

struct Context {
  ssb::Plan plan;
  bool plan_uploaded = false;
  long uploaded_generation = -1;
  DevBuf d_nlay, d_istart, d_irep, d_cols, d_status, d_lay2col;
  static constexpr int kLanes = 3;  // host entry: upload / kernel / download streams; scratch: lane 0 device entry, lane 1 host entry
  DevBuf d_scratch[kLanes], d_perm[kLanes], d_sort[kLanes];
  cudaStream_t lane_stream[kLanes] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> blk_events;  // per pipeline block: upload done, kernels done
  int pipeline = 1;
  int pipeline_max_blocks = 16;
  std::vector<DevBuf> stage;  // staging mirrors of host arrays for ssb200_radsurf
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  size_t budget_doubles = 0;
  bool profiling = false;
  double times_ms[5] = {0, 0, 0, 0, 0};
  long long counts[5] = {0, 0, 0, 0, 0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_done = nullptr;  // end of the last device-entry call (cross-stream ordering)
  cudaStream_t last_stream = nullptr;
  bool last_stream_valid = false;
  cudaError_t first_error = cudaSuccess;
  int fast_mode = 1;
  size_t l2_set_aside = 0;
  long lay2col_generation = -1;
  // order the columns of a launch by their segment pattern (fast kernels): warps then take one code
  // path per layer, the segment kernels write and the sweeps skip whole 32-byte sectors.  Measured on
  // B200 (1 M columns, S2) TOGETHER with the level-major staging below, which makes the column order
  // irrelevant for the per-layer inputs and outputs: 59.8 -> 50.9 ms per step (without the staging the
  // ordering made the sweeps slower: round 1).
  int sort_columns = 1;
  int sort_group = 16384;  // ... inside groups of this many neighbouring columns (0: the whole chunk); measured 1024 .. 65536: 48.0 / 47.3 (4096) / 47.2 / 46.9 (16384) / 47.0 / 47.2 ms per step
  int partition = 1;  // group layer problems by solved sub-block before the fast layer kernels
  // column-resident kernels (ssb_fused.cuh) where they exist (1 and 2 streams); 0: split path
  // sweeps of the register-resident path at 1 and 2 streams: 0 = interface-state sweeps
  // (ssb_fast_sweeps.cuh), 1 = record sweeps (ssb_fused.cuh MODE 1 / 2)
  // shortwave and longwave pass of a call on two streams (separate scratch; no effect while profiling)
  int concurrent_passes = 1;
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int sweep_mode = 0;  // measured on B200: the record sweeps move fewer bytes and are 6 % slower
  // level-major staging of the per-layer arrays of a chunk (ssb_stage.cuh)
  int stage_layers = 1;
  int fused_mode = 0;  // measured on B200: 40 % fewer DRAM bytes, 28 % slower - off by default
  int fused_sort = 1;
  int fused_sort_group = 0;   // 0: the whole chunk
  int fused_sync = 0;         // block-aligned phases (ssb_fused.cuh: phase_sync): measured, no gain
  int fused_blocks_per_sm = 2;
  int fused_l2_persist = 0;   // pin the private tiles in L2 (access policy window): measured, 25 % slower
  int sm_count = 0;
};
Context g_ctx;

struct CudaBackend {
  Context &cx;
  int lane;
  int pass_lane;            // scratch / perm / sort buffers of the pass being dispatched
  cudaStream_t main_stream; // the stream of the call; the longwave pass may run on cx.aux_stream
  bool forked = false;
  size_t budget;
  bool layer_was_fast = false;  // the chunk's layer kernels were the register-resident ones
  explicit CudaBackend(Context &c, int lane_ = 0, size_t budget_ = 0)
      : cx(c), lane(lane_), pass_lane(lane_), main_stream(c.stream), budget(budget_ ? budget_ : c.budget_doubles) {}
  // Shortwave pass on the stream of the call, longwave pass on an auxiliary stream with its own
  // scratch (lane 2): the tails of one pass's launches are filled by the other's blocks, and the
  // memory-bound sweeps of one overlap the FP64-bound layer kernels of the other.
  void fork_passes() {
    if (!cx.concurrent_passes || cx.profiling || cx.first_error != cudaSuccess) return;
    if (!cx.aux_stream && cudaStreamCreateWithFlags(&cx.aux_stream, cudaStreamNonBlocking) != cudaSuccess) return;
    if (!cx.ev_fork) {
      if (cudaEventCreateWithFlags(&cx.ev_fork, cudaEventDisableTiming) != cudaSuccess) return;
      if (cudaEventCreateWithFlags(&cx.ev_join, cudaEventDisableTiming) != cudaSuccess) return;
    }
    cudaEventRecord(cx.ev_fork, main_stream);
    cudaStreamWaitEvent(cx.aux_stream, cx.ev_fork, 0);
    forked = true;
  }
  void begin_pass(bool lw) {
    if (!forked) return;
    cx.stream = lw ? cx.aux_stream : main_stream;
    pass_lane = lw ? 2 : lane;
  }
  void end_passes() {
    if (forked) {
      cudaEventRecord(cx.ev_join, cx.aux_stream);
      cudaStreamWaitEvent(main_stream, cx.ev_join, 0);
      forked = false;
    }
    cx.stream = main_stream;
    pass_lane = lane;
  }
  const int *dev_cols(const ssb::Plan &, size_t off) { return (const int *)cx.d_cols.p + off; }
  // columns of the chunk ordered by their segment pattern (device radix sort, stable)
  // (always for the column-resident kernels: a thread walks the layers of its column, so the
  // warps only stay on one code path when neighbouring columns have the same segment pattern)
  const int *order_chunk(const ssb::ClassArgs &a, const int *) {
    const bool want = a.fused ? cx.fused_sort != 0 : (cx.sort_columns && cx.partition);
    if (!cx.fast_mode || !want || a.ncols < 64 || a.cfg.ns > 4 || cx.first_error != cudaSuccess)
      return a.cols;
    const size_t n = (size_t)a.ncols, pad = (n + 63) & ~(size_t)63;
    typedef unsigned long long Key;
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const Key *)nullptr, (Key *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, (int)n, 0, 64, cx.stream);
    temp_bytes = (temp_bytes + 255) & ~(size_t)255;
    if (cx.d_sort[pass_lane].reserve(temp_bytes + pad * (2 * sizeof(Key) + sizeof(int))) != cudaSuccess) {
      cudaGetLastError();
      return a.cols;
    }
    char *base = (char *)cx.d_sort[pass_lane].p;
    Key *keys = (Key *)(base + temp_bytes), *keys_out = keys + pad;
    int *cols_out = (int *)(keys_out + pad);
    const int group = a.fused ? (cx.fused_sort_group > 0 ? cx.fused_sort_group : 0x7fffffff)
                              : (cx.sort_group > 0 ? cx.sort_group : 0x7fffffff);
    const int lbits = (a.lmax < 31 ? a.lmax : 31) + 1;
    int gbits = 0;
    for (size_t ngroups = (n + (size_t)group - 1) / (size_t)group; ((size_t)1 << gbits) < ngroups; ++gbits) {
    }
    k_column_keys<<<(unsigned)((n + 127) / 128), 128, 0, cx.stream>>>(a, keys, group, lbits);
    cub::DeviceRadixSort::SortPairs(base, temp_bytes, keys, keys_out, a.cols, cols_out, (int)n, 0, lbits + gbits,
                                    cx.stream);
    g_launches += 2;
    check_launch();
    return cols_out;
  }
  const int *dev_nlay() { return (const int *)cx.d_nlay.p; }
  const int *dev_istartlay() { return (const int *)cx.d_istart.p; }
  const int *dev_irep() { return (const int *)cx.d_irep.p; }
  int *dev_status() { return (int *)cx.d_status.p; }
  size_t scratch_budget_doubles() { return budget; }
  double *scratch(size_t n) {
    cudaError_t e = cx.d_scratch[pass_lane].reserve(n * sizeof(double));
    if (e != cudaSuccess && cx.first_error == cudaSuccess) cx.first_error = e;
    return (double *)cx.d_scratch[pass_lane].p;
  }
  void tick(int family, bool start) {
    if (!cx.profiling) return;
    if (start) {
      cudaEventRecord(cx.ev0, cx.stream);
    } else {
      cudaEventRecord(cx.ev1, cx.stream);
      cudaEventSynchronize(cx.ev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, cx.ev0, cx.ev1);
      cx.times_ms[family] += ms;
      cx.counts[family] += 1;
    }
  }
  void check_launch() {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = ssb::g_fast_error;
    ssb::g_fast_error = cudaSuccess;
    if (e != cudaSuccess && cx.first_error == cudaSuccess) cx.first_error = e;
  }
  template <class F>
  void launch(F launcher, const ssb::ClassArgs &a, long nt, int family) {
    if (nt <= 0 || cx.first_error != cudaSuccess) return;
    tick(family, true);
    launcher(a, nt, cx.stream);
    check_launch();
    tick(family, false);
  }
  template <int NS>
  void layer_sw(const ssb::ClassArgs &a, long nt) {
    if (cx.fast_mode && nt > 0 && cx.first_error == cudaSuccess) {
      tick(0, true);
      ssb::ClassArgs b = a;
      if (cx.partition && cx.d_perm[pass_lane].reserve(sizeof(int) * (3 * (size_t)nt + 4)) == cudaSuccess) {
        b.perm_count = (int *)cx.d_perm[pass_lane].p;
        b.perm = b.perm_count + 4;
      }
      const bool done = ssb::fast_layer_sw<NS>(b, nt, cx.stream);
      layer_was_fast = done;
      if (done && b.perm) g_launches += 3;
      if (done) check_launch();
      tick(0, false);
      if (done) return;
    }
    if (a.lstride) {  // staged arrays are only understood by the register-resident kernels
      if (cx.first_error == cudaSuccess) cx.first_error = cudaErrorNotSupported;
      return;
    }
    launch(ssb::launch_layer_sw<NS>, a, nt, 0);
  }
  template <int NS>
  void layer_lw(const ssb::ClassArgs &a, long nt) {
    if (cx.fast_mode && nt > 0 && cx.first_error == cudaSuccess) {
      tick(2, true);
      ssb::ClassArgs b = a;
      if (cx.partition && cx.d_perm[pass_lane].reserve(sizeof(int) * (3 * (size_t)nt + 4)) == cudaSuccess) {
        b.perm_count = (int *)cx.d_perm[pass_lane].p;
        b.perm = b.perm_count + 4;
      }
      const bool done = ssb::fast_layer_lw<NS>(b, nt, cx.stream);
      layer_was_fast = done;
      if (done && b.perm) g_launches += 3;
      if (done) check_launch();
      tick(2, false);
      if (done) return;
    }
    if (a.lstride) {
      if (cx.first_error == cudaSuccess) cx.first_error = cudaErrorNotSupported;
      return;
    }
    launch(ssb::launch_layer_lw<NS>, a, nt, 2);
  }
  template <int NS>
  void sweeps_sw(const ssb::ClassArgs &a, long nt) {
    // the register-resident sweeps read what the partition pass of the layer kernels prepared
    if (cx.fast_mode && (layer_was_fast || a.lmax == 0) && nt > 0 && cx.first_error == cudaSuccess) {
      tick(1, true);
      const bool done = ssb::fast_sweeps_sw<NS>(a, nt, cx.stream);
      if (done) check_launch();
      tick(1, false);
      if (done) return;
    }
    if (a.lstride) {
      if (cx.first_error == cudaSuccess) cx.first_error = cudaErrorNotSupported;
      return;
    }
    launch(ssb::launch_sweeps_sw<NS>, a, nt, 1);
  }
  template <int NS>
  void sweeps_lw(const ssb::ClassArgs &a, long nt) {
    if (cx.fast_mode && (layer_was_fast || a.lmax == 0) && nt > 0 && cx.first_error == cudaSuccess) {
      tick(3, true);
      const bool done = ssb::fast_sweeps_lw<NS>(a, nt, cx.stream);
      if (done) check_launch();
      tick(3, false);
      if (done) return;
    }
    if (a.lstride) {
      if (cx.first_error == cudaSuccess) cx.first_error = cudaErrorNotSupported;
      return;
    }
    launch(ssb::launch_sweeps_lw<NS>, a, nt, 3);
  }
  bool fused_shape(const ssb::SolveCfg &c, bool lw, int *pe, int *oe, int *geo) {
    if (!cx.fast_mode || !cx.fused_mode || c.ns > 2) return false;
    return ssb::fused_shape(c, lw, pe, oe, geo);
  }
  bool stage_supported(const ssb::SolveCfg &c) {
    return cx.fast_mode && cx.partition && cx.stage_layers && c.ns <= 4 && c.nspec <= 1024;
  }
  void stage(const ssb::StageArgs &s, bool scatter, bool lw) {
    if (cx.first_error != cudaSuccess) return;
    const int fam = 4;  // booked with the surface kernels ("surface" family of ssb200_last_kernel_times_ms)
    (void)lw;
    tick(fam, true);
    long n = 0;
    ssb::stage_launch(s, scatter, cx.stream, &n);
    g_launches += n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && cx.first_error == cudaSuccess) cx.first_error = e;
    tick(fam, false);
  }
  bool records_shape(const ssb::SolveCfg &c, bool lw, int *oe) {
    if (!cx.fast_mode || cx.sweep_mode != 1 || !cx.partition || c.ns > 2) return false;
    int pe = 0, geo = 0;
    return ssb::fused_shape(c, lw, &pe, oe, &geo);
  }
  void records_run(const ssb::ClassArgs &a, bool lw, long width) {
    if (width <= 0 || cx.first_error != cudaSuccess) return;
    if (!layer_was_fast && a.lmax > 0) {  // the records are built from the register-resident layer scratch
      cx.first_error = cudaErrorNotSupported;
      return;
    }
    const int fam = lw ? 3 : 1;
    tick(fam, true);
    ssb::records_launch(a, lw, width, cx.stream);
    g_launches += 1;  // (check_launch counts one)
    check_launch();
    tick(fam, false);
  }
  int fused_slots() {
    if (cx.sm_count <= 0) {
      int dev = 0, n = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      cx.sm_count = n > 0 ? n : 148;
    }
    return cx.fused_blocks_per_sm * cx.sm_count;  // __launch_bounds__(128, 2)
  }
  int fused_flags() { return 1 | (cx.fused_sync ? 2 : 0); }
  // The private tiles are re-read and re-written for every layer of every column while the operator
  // records, inputs and outputs stream through L2 once: keep the former resident with a persisting
  // access-policy window on the launch stream (the set-aside is capped by the device limit).
  void l2_window(const void *base, size_t bytes) {
    if (!cx.fused_l2_persist) return;
    int dev = 0, max_persist = 0, max_window = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_persist <= 0 || max_window <= 0) return;
    const size_t set_aside = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
    if (bytes > 0 && cx.l2_set_aside != set_aside) {
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside) == cudaSuccess) cx.l2_set_aside = set_aside;
      cudaGetLastError();
    }
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    v.accessPolicyWindow.num_bytes = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    v.accessPolicyWindow.hitRatio = bytes > 0 ? (float)((double)set_aside / (double)bytes) : 0.0f;
    if (v.accessPolicyWindow.hitRatio > 1.0f) v.accessPolicyWindow.hitRatio = 1.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    cudaStreamSetAttribute(cx.stream, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();
  }
  void fused_run(const ssb::ClassArgs &a, bool lw, long width) {
    if (width <= 0 || cx.first_error != cudaSuccess) return;
    const int fam = lw ? 2 : 0;  // booked under the layer families (kernel_times: sw_layer / lw_layer)
    tick(fam, true);
    const long tiles = (width + ssb::kScratchTile - 1) / ssb::kScratchTile;
    const int grid = (int)(tiles < (long)fused_slots() ? tiles : (long)fused_slots());
    l2_window(a.layer, (size_t)grid * (size_t)a.ne_layer * ssb::kScratchTile * sizeof(double));
    ssb::fused_launch(a, lw, width, grid, cx.stream);
    check_launch();
    l2_window(nullptr, 0);
    tick(fam, false);
  }
  void surface(const ssb::SurfaceArgs &s, int nsw_threads, int nlw_threads) {
    const int nt = nsw_threads > nlw_threads ? nsw_threads : nlw_threads;
    if (nt <= 0 || cx.first_error != cudaSuccess) return;
    tick(4, true);
    k_surface<<<(nt + 127) / 128, 128, 0, cx.stream>>>(s, nsw_threads, nlw_threads);
    check_launch();
    tick(4, false);
  }
};

int ensure_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(SSB200_ERR_NOGPU, "no CUDA device visible: libspartacus_b200 has no CPU fallback");
  }
  return 0;
}

int upload_plan(Context &cx, const ssb200_canopy_properties &cp) {
  const size_t ncol = (size_t)cp.ncol;
  SSB_CUDA(cx.d_nlay.reserve(sizeof(int) * (ncol + 1)));
  SSB_CUDA(cx.d_istart.reserve(sizeof(int) * (ncol + 1)));
  SSB_CUDA(cx.d_irep.reserve(sizeof(int) * (ncol + 1)));
  SSB_CUDA(cx.d_cols.reserve(sizeof(int) * (cx.plan.all_cols.size() + 1)));
  SSB_CUDA(cudaMemcpyAsync(cx.d_nlay.p, cx.plan.nlay.data(), sizeof(int) * ncol, cudaMemcpyHostToDevice, cx.stream));
  SSB_CUDA(cudaMemcpyAsync(cx.d_istart.p, cx.plan.istartlay.data(), sizeof(int) * ncol, cudaMemcpyHostToDevice,
                           cx.stream));
  SSB_CUDA(cudaMemcpyAsync(cx.d_irep.p, cx.plan.irep.data(), sizeof(int) * ncol, cudaMemcpyHostToDevice, cx.stream));
  SSB_CUDA(cudaMemcpyAsync(cx.d_cols.p, cx.plan.all_cols.data(), sizeof(int) * cx.plan.all_cols.size(),
                           cudaMemcpyHostToDevice, cx.stream));
  // the host vectors must stay unchanged until the copies are done
  SSB_CUDA(cudaStreamSynchronize(cx.stream));
  cx.plan_uploaded = true;
  cx.uploaded_generation = cx.plan.generation;
  return 0;
}

// Core of both entry points; all double arrays are device pointers here.
// `blocks`: when non-null the call only prepares (plan, status, budget) and returns the
// clamped 1-based column range in blocks[0..1]; the caller then dispatches column
// windows itself with run_window().
int radsurf_device_locked(Context &cx, const ssb::CallArgs &ca, int istartcol, int iendcol, cudaStream_t stream,
                          int32_t *status_out, int *blocks = nullptr) {
  std::string err;
  int rc = ssb::validate_call(ca, err);
  if (rc) return fail(rc, err);
  const ssb200_canopy_properties &cp = *ca.cp;
  int c1 = istartcol > 0 ? istartcol : 1;
  int c2 = iendcol > 0 ? iendcol : cp.ncol;
  if (c2 > cp.ncol) c2 = cp.ncol;
  if (c1 > c2) return 0;
  cx.stream = stream;
  cx.first_error = cudaSuccess;
  rc = ssb::build_plan(*ca.config, cp, c1 - 1, c2 - 1, cx.plan, err);
  if (rc) {
    cx.plan.valid = false;
    return fail(rc, err);
  }
  rc = ssb::validate_members(ca, cx.plan, err);
  if (rc) return fail(rc, err);
  // One Context (scratch, status word, plan buffers) serves every call: work of an earlier call
  // that is still in flight on ANOTHER stream must finish before this call reuses them.
  if (cx.ev_done && cx.last_stream_valid && cx.last_stream != stream) SSB_CUDA(cudaStreamWaitEvent(stream, cx.ev_done, 0));
  if (cx.plan.generation != cx.uploaded_generation) cx.plan_uploaded = false;
  if (!cx.plan_uploaded) {
    rc = upload_plan(cx, cp);
    if (rc) return rc;
  }
  SSB_CUDA(cx.d_status.reserve(sizeof(int)));
  SSB_CUDA(cudaMemsetAsync(cx.d_status.p, 0, sizeof(int), stream));
  if (cx.budget_doubles == 0) {
    size_t free_b = 0, total_b = 0;
    SSB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // per scratch lane (two lanes are live when the SW and LW passes run concurrently)
    size_t b = (free_b + cx.d_scratch[0].bytes + cx.d_scratch[1].bytes + cx.d_scratch[2].bytes) / 4;
    const size_t cap = (size_t)12 << 30;  // (more, smaller chunks overlap the two passes better: 47.0 -> 46.7 ms against 24 GiB)
    if (b > cap) b = cap;
    cx.budget_doubles = b / sizeof(double);
  }
  if (cx.profiling) {
    if (!cx.ev0) {
      cudaEventCreate(&cx.ev0);
      cudaEventCreate(&cx.ev1);
    }
    for (double &t : cx.times_ms) t = 0.0;
    for (long long &n : cx.counts) n = 0;
  }
  if (blocks) {
    blocks[0] = c1;
    blocks[1] = c2;
    return 0;
  }
  CudaBackend be(cx);
  ssb::Dispatcher<CudaBackend> disp(be);
  rc = disp.run(ca, cx.plan, err);
  if (rc) return fail(rc, err);
  if (cx.first_error != cudaSuccess)
    return fail(SSB200_ERR_CUDA, std::string("kernel launch / scratch allocation: ") + cudaGetErrorString(cx.first_error));
  if (status_out)
    SSB_CUDA(cudaMemcpyAsync(status_out, cx.d_status.p, sizeof(int), cudaMemcpyDeviceToDevice, stream));
  if (!cx.ev_done) SSB_CUDA(cudaEventCreateWithFlags(&cx.ev_done, cudaEventDisableTiming));
  SSB_CUDA(cudaEventRecord(cx.ev_done, stream));
  cx.last_stream = stream;
  cx.last_stream_valid = true;
  return 0;
}

// dispatch the columns [col_lo, col_hi) (0-based) of the prepared plan on one lane
int run_window(Context &cx, const ssb::CallArgs &ca, int lane, size_t budget, int col_lo, int col_hi) {
  std::string err;
  cx.stream = cx.lane_stream[lane];
  CudaBackend be(cx, lane, budget);
  ssb::Dispatcher<CudaBackend> disp(be);
  disp.col_lo = col_lo;
  disp.col_hi = col_hi;
  const int rc = disp.run(ca, cx.plan, err);
  if (rc) return fail(rc, err);
  if (cx.first_error != cudaSuccess)
    return fail(SSB200_ERR_CUDA, std::string("kernel launch / scratch allocation: ") + cudaGetErrorString(cx.first_error));
  return 0;
}

static int collect_fields(ssb200_canopy_flux *o, const ssb200_canopy_flux *a, const ssb200_canopy_flux *b,
                          FieldList &fl, bool with_sunlit) {
  fl.n = 0;
  auto add = [&](double *p, const double *pa, const double *pb, int per_layer, int spectral) {
    if (!p) return;
    if (a && (!pa || !pb)) return;
    fl.p[fl.n] = p;
    fl.a[fl.n] = pa;
    fl.b[fl.n] = pb;
    fl.per_layer[fl.n] = per_layer;
    fl.spectral[fl.n] = spectral;
    ++fl.n;
  };
#define F(m, pl, sp) add(o->m, a ? a->m : nullptr, b ? b->m : nullptr, pl, sp)
  F(ground_dn, 0, 1);
  F(ground_net, 0, 1);
  F(ground_vertical_diff, 0, 1);
  F(top_dn, 0, 1);
  F(top_net, 0, 1);
  F(ground_dn_dir, 0, 1);
  F(top_dn_dir, 0, 1);
  F(roof_in, 1, 1);
  F(roof_net, 1, 1);
  F(wall_in, 1, 1);
  F(wall_net, 1, 1);
  F(roof_in_dir, 1, 1);
  F(wall_in_dir, 1, 1);
  F(clear_air_abs, 1, 1);
  F(veg_abs, 1, 1);
  F(veg_air_abs, 1, 1);
  F(veg_abs_dir, 1, 1);
  F(flux_dn_layer_top, 1, 1);
  F(flux_up_layer_top, 1, 1);
  F(flux_dn_layer_base, 1, 1);
  F(flux_up_layer_base, 1, 1);
  F(flux_dn_dir_layer_top, 1, 1);
  F(flux_dn_dir_layer_base, 1, 1);
  if (with_sunlit) {
    F(ground_sunlit_frac, 0, 0);
    F(roof_sunlit_frac, 1, 0);
    F(wall_sunlit_frac, 1, 0);
    F(veg_sunlit_frac, 1, 0);
  }
#undef F
  return fl.n;
}

// --- host-pointer staging --------------------------------------------------
struct Stager {
  Context &cx;
  cudaStream_t stream;
  int next = 0;
  int rc = 0;
  struct Arr {
    double *host;    // (sp: really float *)
    double *dev;
    float *dev32;    // sp: device copy in the caller's precision
    size_t width;    // doubles per column (per_layer = false) or per packed layer (true)
    bool per_layer, upload, download;
  };
  std::vector<Arr> arrs;
  // single-precision storage (ssb200_radsurf_sp): the caller's arrays hold float; they cross PCIe as
  // float and are widened / rounded on the device, the kernels see double mirrors as always
  bool sp = false;
  Stager(Context &c, cudaStream_t s, bool sp_ = false) : cx(c), stream(s), sp(sp_) {}
  // device mirror of `host` (total = rows * width doubles); copies are issued per window
  double *mirror(const double *host, size_t rows, size_t width, bool per_layer, bool upload, bool download) {
    if (!host || rc) return nullptr;
    if ((size_t)next >= cx.stage.size()) cx.stage.resize((size_t)next + 16);
    DevBuf &b = cx.stage[next++];
    const size_t total = rows * width;
    cudaError_t e = b.reserve((total > 0 ? total : 1) * sizeof(double));
    if (e != cudaSuccess) {
      rc = fail(SSB200_ERR_CUDA, std::string("staging allocation: ") + cudaGetErrorString(e));
      return nullptr;
    }
    float *d32 = nullptr;
    if (sp) {
      if ((size_t)next >= cx.stage.size()) cx.stage.resize((size_t)next + 16);
      DevBuf &b32 = cx.stage[next++];
      e = b32.reserve((total > 0 ? total : 1) * sizeof(float));
      if (e != cudaSuccess) {
        rc = fail(SSB200_ERR_CUDA, std::string("staging allocation: ") + cudaGetErrorString(e));
        return nullptr;
      }
      d32 = (float *)b32.p;
    }
    arrs.push_back(Arr{const_cast<double *>(host), (double *)cx.stage[next - (sp ? 2 : 1)].p, d32, width, per_layer, upload, download});
    return arrs.back().dev;
  }
  // device-only array (no host counterpart: never copied)
  double *device_only(size_t total) {
    if (rc) return nullptr;
    if ((size_t)next >= cx.stage.size()) cx.stage.resize((size_t)next + 16);
    DevBuf &b = cx.stage[next++];
    cudaError_t e = b.reserve((total > 0 ? total : 1) * sizeof(double));
    if (e != cudaSuccess) {
      rc = fail(SSB200_ERR_CUDA, std::string("staging allocation: ") + cudaGetErrorString(e));
      return nullptr;
    }
    return (double *)b.p;
  }
  // copy the slices of columns [c0, c1) / packed layers [l0, l1) on stream `st`
  int copy_window(bool to_device, size_t c0, size_t c1, size_t l0, size_t l1, cudaStream_t st) {
    for (const Arr &a : arrs) {
      if (to_device ? !a.upload : !a.download) continue;
      const size_t off = (a.per_layer ? l0 : c0) * a.width, cnt = ((a.per_layer ? l1 : c1) * a.width) - off;
      if (cnt == 0) continue;
      cudaError_t e;
      if (sp) {
        float *h32 = reinterpret_cast<float *>(a.host);
        const unsigned blocks = (unsigned)((cnt + 255) / 256);
        if (to_device) {
          e = cudaMemcpyAsync(a.dev32 + off, h32 + off, cnt * sizeof(float), cudaMemcpyHostToDevice, st);
          k_f2d<<<blocks, 256, 0, st>>>(a.dev + off, a.dev32 + off, cnt);
        } else {
          k_d2f<<<blocks, 256, 0, st>>>(a.dev32 + off, a.dev + off, cnt);
          e = cudaMemcpyAsync(h32 + off, a.dev32 + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, st);
        }
        ++g_launches;
        if (e == cudaSuccess) e = cudaGetLastError();
      } else {
        e = to_device ? cudaMemcpyAsync(a.dev + off, a.host + off, cnt * sizeof(double), cudaMemcpyHostToDevice, st)
                      : cudaMemcpyAsync(a.host + off, a.dev + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, st);
      }
      if (e != cudaSuccess)
        return fail(SSB200_ERR_CUDA, std::string(to_device ? "H2D copy: " : "D2H copy: ") + cudaGetErrorString(e));
    }
    return 0;
  }
};

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char *ssb200_version(void) { return "spartacus_surface_b200 0.1 (sm_100a, generic + register-resident kernels)"; }

const char *ssb200_last_error(void) { return g_last_error.c_str(); }

int ssb200_abi_sizes(int64_t out[7]) {
  if (!out) return fail(SSB200_ERR_ARG, "NULL argument");
  out[0] = sizeof(ssb200_legendre_gauss);
  out[1] = sizeof(ssb200_config);
  out[2] = sizeof(ssb200_canopy_properties);
  out[3] = sizeof(ssb200_sw_spectral_properties);
  out[4] = sizeof(ssb200_lw_spectral_properties);
  out[5] = sizeof(ssb200_canopy_flux);
  out[6] = sizeof(ssb200_boundary_conds_out);
  return 0;
}

int ssb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ssb200_set_device(int device) {
  int rc = ensure_device();
  if (rc) return rc;
  SSB_CUDA(cudaSetDevice(device));
  return 0;
}

// legendre_gauss_type%initialize (radtool/radtool_legendre_gauss.F90:52-100,119-170):
// Newton iteration for the Legendre nodes mapped to [0,1]; host side, once per config.
int ssb200_legendre_gauss_init(int32_t nstream, ssb200_legendre_gauss *lg) {
  if (!lg || nstream < 1 || nstream > SSB200_MAX_NSTREAM) return fail(SSB200_ERR_ARG, "nstream outside 1..16");
  const int n = nstream;
  const double pi = SSB_PI, eps = SSB_EPS;
  double y[SSB200_MAX_NSTREAM], y0[SSB200_MAX_NSTREAM], dp[SSB200_MAX_NSTREAM];
  double P[SSB200_MAX_NSTREAM + 1][SSB200_MAX_NSTREAM];
  const float c027 = 0.27f / (float)n;  // single-precision constant expression in the reference (:142)
  for (int k = 1; k <= n; ++k) {
    y[k - 1] = cos((2 * (k - 1) + 1) * pi / (2 * n)) + (double)c027 * sin(pi * (-1.0 + 2.0 * k) / (n + 1));
    y0[k - 1] = 2.0;
  }
  for (;;) {
    double md = 0.0;
    for (int i = 0; i < n; ++i) md = fmax(md, fabs(y[i] - y0[i]));
    if (!(md > eps)) break;
    for (int i = 0; i < n; ++i) {
      P[0][i] = 1.0;
      P[1][i] = y[i];
    }
    for (int k = 2; k <= n; ++k)
      for (int i = 0; i < n; ++i) P[k][i] = ((2 * k - 1) * y[i] * P[k - 1][i] - (k - 1) * P[k - 2][i]) / k;
    for (int i = 0; i < n; ++i) dp[i] = (n + 1) * (P[n - 1][i] - y[i] * P[n][i]) / (1.0 - y[i] * y[i]);
    for (int i = 0; i < n; ++i) {
      y0[i] = y[i];
      y[i] = y0[i] - P[n][i] / dp[i];
    }
  }
  memset(lg, 0, sizeof(*lg));
  lg->nstream = n;
  double sh = 0.0, sv = 0.0, s2 = 0.0;
  for (int i = 0; i < n; ++i) {
    lg->mu[i] = 0.5 * (0.0 * (1.0 - y[i]) + 1.0 * (1.0 - y[i]));  // map as written in the reference (:165)
    lg->weight[i] = (((n + 1) * (n + 1)) / (double)(n * n)) * (1.0 - 0.0) / ((1.0 - y[i] * y[i]) * dp[i] * dp[i]);
    lg->sin_ang[i] = sqrt(1.0 - lg->mu[i] * lg->mu[i]);
    lg->tan_ang[i] = lg->sin_ang[i] / lg->mu[i];
    lg->hweight[i] = lg->weight[i] * lg->mu[i];
    lg->vweight[i] = lg->weight[i] * lg->sin_ang[i];
  }
  for (int i = 0; i < n; ++i) {
    sh += lg->hweight[i];
    sv += lg->vweight[i];
  }
  for (int i = 0; i < n; ++i) {
    lg->hweight[i] = lg->hweight[i] / sh;
    lg->vweight[i] = lg->vweight[i] / sv;
  }
  lg->vadjustment = 1.0;
  for (int i = 0; i < n; ++i) s2 += lg->weight[i] * lg->sin_ang[i];
  lg->vadjustment2 = (pi / 4.0) / s2;
  return 0;
}

int ssb200_radsurf_device(const ssb200_config *config, const ssb200_canopy_properties *canopy_props,
                          const ssb200_sw_spectral_properties *sw, const ssb200_lw_spectral_properties *lw,
                          ssb200_boundary_conds_out *bc_out, int32_t istartcol, int32_t iendcol,
                          ssb200_canopy_flux *sw_norm_dir, ssb200_canopy_flux *sw_norm_diff,
                          ssb200_canopy_flux *lw_internal, ssb200_canopy_flux *lw_norm, void *stream,
                          int32_t *status_out) {
  int rc = ensure_device();
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_mutex);
  ssb::CallArgs ca{config, canopy_props, sw, lw, bc_out, sw_norm_dir, sw_norm_diff, lw_internal, lw_norm};
  return radsurf_device_locked(g_ctx, ca, istartcol, iendcol, (cudaStream_t)stream, status_out);
}

// Body of both host entries.  drv == NULL: ssb200_radsurf (the four normalised flux objects are
// host arrays).  drv != NULL: ssb200_radsurf_fluxes - the normalised objects exist on the device
// only (shaped like sw_flux / lw_flux), NULL inputs take read_input's defaults or are derived from
// the temperatures, and every window ends with the fused scale + sum into sw_flux / lw_flux.
static int radsurf_host(const ssb200_config *config, const ssb200_canopy_properties *cp,
                        const ssb200_sw_spectral_properties *sw, const ssb200_lw_spectral_properties *lw,
                        const ssb200_driver_inputs *drv, ssb200_boundary_conds_out *bc, int32_t istartcol,
                        int32_t iendcol, ssb200_canopy_flux *sw_dir, ssb200_canopy_flux *sw_diff,
                        ssb200_canopy_flux *lw_int, ssb200_canopy_flux *lw_norm, ssb200_canopy_flux *sw_flux,
                        ssb200_canopy_flux *lw_flux, bool sp = false) {
  int rc = ensure_device();
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_mutex);
  Context &cx = g_ctx;
  if (!config || !cp || !bc) return fail(SSB200_ERR_ARG, "config, canopy_props and bc_out must not be NULL");
  if (drv) {
    if ((config->do_sw && (!sw_flux || !drv->top_flux_dn_sw || !drv->top_flux_dn_direct_sw)) ||
        (config->do_lw && (!lw_flux || !drv->top_flux_dn_lw)))
      return fail(SSB200_ERR_ARG, "radsurf_fluxes: flux object or top-of-canopy flux missing");
    sw_dir = sw_diff = sw_flux;  // shape donors for the argument checks
    lw_int = lw_norm = lw_flux;
  }
  {
    ssb::CallArgs probe{config, cp, sw, lw, bc, sw_dir, sw_diff, lw_int, lw_norm};
    std::string err;
    rc = ssb::validate_call(probe, err);
    if (rc) return fail(rc, err);
  }
  if (!cx.own_stream) SSB_CUDA(cudaStreamCreateWithFlags(&cx.own_stream, cudaStreamNonBlocking));
  cudaStream_t st = cx.own_stream;
  int c1 = istartcol > 0 ? istartcol : 1, c2 = iendcol > 0 ? iendcol : cp->ncol;
  if (c2 > cp->ncol) c2 = cp->ncol;
  if (c1 > c2) return 0;
  const size_t ncol = (size_t)cp->ncol, ntot = (size_t)cp->ntotlay;
  // packed layer range touched by the columns of this call; check that it is
  // contiguous so that per-layer outputs need no upload
  size_t l1 = ntot, l2 = 0;
  bool contiguous = true;
  size_t expect = 0;
  bool first = true;
  for (int j = c1 - 1; j < c2; ++j) {
    if (cp->i_representation[j] == SSB200_TILE_FLAT || cp->nlay[j] <= 0) continue;
    const size_t s = (size_t)cp->istartlay[j] - 1, e = s + (size_t)cp->nlay[j];
    if (cp->istartlay[j] < 1 || e > ntot) return fail(SSB200_ERR_SHAPE, "layer range outside 1..ntotlay");
    if (!first && s != expect) contiguous = false;
    first = false;
    expect = e;
    if (s < l1) l1 = s;
    if (e > l2) l2 = e;
  }
  if (l2 < l1) l1 = l2 = 0;
  const size_t cO = (size_t)(c1 - 1), cN = (size_t)(c2 - c1 + 1), lN = l2 - l1;
  if (drv && !contiguous)
    return fail(SSB200_ERR_UNSUPPORTED, "radsurf_fluxes needs the packed layers of the selected columns to be contiguous");
  Stager sg(cx, st, sp);
  ssb200_canopy_properties dcp = *cp;
  auto lay1 = [&](const double *h) { return (const double *)sg.mirror(h, ntot, 1, true, true, false); };
  dcp.cos_sza = sg.mirror(cp->cos_sza, ncol, 1, false, true, false);
  dcp.dz = lay1(cp->dz);
  dcp.building_fraction = lay1(cp->building_fraction);
  dcp.building_scale = lay1(cp->building_scale);
  dcp.veg_fraction = lay1(cp->veg_fraction);
  dcp.veg_scale = lay1(cp->veg_scale);
  dcp.veg_ext = lay1(cp->veg_ext);
  dcp.veg_fsd = lay1(cp->veg_fsd);
  dcp.veg_contact_fraction = lay1(cp->veg_contact_fraction);
  // read_input's defaults for members the caller leaves NULL (ssb200_radsurf_fluxes only)
  struct Fill {
    double *p;
    double v;
    size_t n;
  };
  struct Emis {  // out = sigma [emissivity] T^4 on the device, per window
    double *out;
    const double *emissivity, *temperature;
    bool per_layer;
  };
  std::vector<Emis> emis;
  std::vector<Fill> fills;
  auto filled = [&](size_t width, double v) -> const double * {
    double *p = sg.device_only(ntot * width);
    fills.push_back(Fill{p, v, ntot * width});
    return p;
  };
  double *d_vcf_default = nullptr;
  if (drv && !cp->veg_contact_fraction && cp->veg_fraction && cp->building_fraction) {
    d_vcf_default = sg.device_only(ntot);
    dcp.veg_contact_fraction = d_vcf_default;
  }
  ssb200_sw_spectral_properties dsw;
  ssb200_lw_spectral_properties dlw;
  memset(&dsw, 0, sizeof(dsw));
  memset(&dlw, 0, sizeof(dlw));
  if (config->do_sw) {
    const size_t g = (size_t)config->nsw;
    auto L = [&](const double *h) { return (const double *)sg.mirror(h, ntot, g, true, true, false); };
    auto Cc = [&](const double *h) { return (const double *)sg.mirror(h, ncol, g, false, true, false); };
    dsw.nspec = sw->nspec;
    dsw.air_ext = L(sw->air_ext);
    dsw.air_ssa = L(sw->air_ssa);
    dsw.veg_ssa = L(sw->veg_ssa);
    dsw.ground_albedo = Cc(sw->ground_albedo);
    dsw.roof_albedo = L(sw->roof_albedo);
    dsw.wall_albedo = L(sw->wall_albedo);
    dsw.wall_specular_frac = L(sw->wall_specular_frac);
    dsw.ground_albedo_dir = Cc(sw->ground_albedo_dir);
    dsw.roof_albedo_dir = L(sw->roof_albedo_dir);
    if (drv) {  // driver/spartacus_surface_read_input.F90:258-269,335-344 (roof_albedo_dir NULL: the kernels use roof_albedo)
      if (!sw->air_ext) dsw.air_ext = filled(g, 1.0e-5);
      if (!sw->air_ssa) dsw.air_ssa = filled(g, 0.999);
      if (!sw->wall_specular_frac && sw->wall_albedo) dsw.wall_specular_frac = filled(g, 0.0);
    }
  }
  if (config->do_lw) {
    const size_t g = (size_t)config->nlw;
    auto L = [&](const double *h) { return (const double *)sg.mirror(h, ntot, g, true, true, false); };
    auto Cc = [&](const double *h) { return (const double *)sg.mirror(h, ncol, g, false, true, false); };
    dlw.nspec = lw->nspec;
    dlw.air_ext = L(lw->air_ext);
    dlw.air_ssa = L(lw->air_ssa);
    dlw.clear_air_planck = L(lw->clear_air_planck);
    dlw.veg_ssa = L(lw->veg_ssa);
    dlw.veg_planck = L(lw->veg_planck);
    dlw.veg_air_planck = L(lw->veg_air_planck);
    dlw.ground_emissivity = Cc(lw->ground_emissivity);
    dlw.ground_emission = Cc(lw->ground_emission);
    dlw.roof_emissivity = L(lw->roof_emissivity);
    dlw.wall_emissivity = L(lw->wall_emissivity);
    dlw.roof_emission = L(lw->roof_emission);
    dlw.wall_emission = L(lw->wall_emission);
    if (drv) {  // driver/spartacus_surface_read_input.F90:362-365; radsurf_simple_spectrum.F90:41-66
      if (!lw->air_ext) dlw.air_ext = filled(g, 1.0e-5);
      if (!lw->air_ssa) dlw.air_ssa = filled(g, 0.0);
      // (the same host array given for several temperatures - the driver passes one air temperature
      // for clear air, vegetation and the air in vegetation - is uploaded once)
      std::vector<std::pair<const double *, const double *>> t_seen;
      auto T1 = [&](const double *h, size_t rows, bool per_layer) {
        for (const auto &s : t_seen)
          if (s.first == h) return s.second;
        const double *d = (const double *)sg.mirror(h, rows, 1, per_layer, true, false);
        t_seen.push_back({h, d});
        return d;
      };
      const bool need_t = !lw->ground_emission || !lw->roof_emission || !lw->wall_emission ||
                          !lw->clear_air_planck || !lw->veg_planck || !lw->veg_air_planck;
      if (need_t && g != 1)
        return fail(SSB200_ERR_ARG, "Simple longwave spectrum only possible with one input spectral interval");
      if (!lw->ground_emission && drv->ground_temperature) {
        emis.push_back(Emis{sg.device_only(ncol), dlw.ground_emissivity, T1(drv->ground_temperature, ncol, false), false});
        dlw.ground_emission = emis.back().out;
      }
      auto layer_emis = [&](const double *host_out, const double *&slot, const double *emissivity, const double *t) {
        if (host_out || !t) return;
        emis.push_back(Emis{sg.device_only(ntot), emissivity, T1(t, ntot, true), true});
        slot = emis.back().out;
      };
      layer_emis(lw->roof_emission, dlw.roof_emission, dlw.roof_emissivity, drv->roof_temperature);
      layer_emis(lw->wall_emission, dlw.wall_emission, dlw.wall_emissivity, drv->wall_temperature);
      layer_emis(lw->clear_air_planck, dlw.clear_air_planck, nullptr, drv->clear_air_temperature);
      layer_emis(lw->veg_planck, dlw.veg_planck, nullptr, drv->veg_temperature);
      layer_emis(lw->veg_air_planck, dlw.veg_air_planck, nullptr, drv->veg_air_temperature);
    }
  }
  // outputs: per-column members are uploaded first (Flat tiles and night-time
  // columns leave some of them untouched); per-layer members are fully
  // rewritten by the kernels when the layer range is contiguous
  ssb200_boundary_conds_out dbc;
  memset(&dbc, 0, sizeof(dbc));
  {
    const size_t gs = (size_t)config->nsw, gl = (size_t)config->nlw;
    if (config->do_sw) {
      dbc.sw_albedo = sg.mirror(bc->sw_albedo, ncol, gs, false, true, true);
      dbc.sw_albedo_dir = sg.mirror(bc->sw_albedo_dir, ncol, gs, false, true, true);
    }
    if (config->do_lw) {
      dbc.lw_emissivity = sg.mirror(bc->lw_emissivity, ncol, gl, false, true, true);
      dbc.lw_emission = sg.mirror(bc->lw_emission, ncol, gl, false, true, true);
    }
  }
  auto stage_flux = [&](const ssb200_canopy_flux *h, ssb200_canopy_flux &d) {
    d = *h;
    const size_t g = (size_t)h->nspec;
    auto Cc = [&](double *p) { return sg.mirror(p, ncol, g, false, true, true); };
    auto L = [&](double *p) { return sg.mirror(p, ntot, g, true, !contiguous, true); };
    auto C1 = [&](double *p) { return sg.mirror(p, ncol, 1, false, true, true); };
    auto L1 = [&](double *p) { return sg.mirror(p, ntot, 1, true, !contiguous, true); };
    d.ground_dn = Cc(h->ground_dn);
    d.ground_net = Cc(h->ground_net);
    d.ground_vertical_diff = Cc(h->ground_vertical_diff);
    d.top_dn = Cc(h->top_dn);
    d.top_net = Cc(h->top_net);
    d.ground_dn_dir = Cc(h->ground_dn_dir);
    d.top_dn_dir = Cc(h->top_dn_dir);
    d.ground_sunlit_frac = C1(h->ground_sunlit_frac);
    d.roof_in = L(h->roof_in);
    d.roof_net = L(h->roof_net);
    d.wall_in = L(h->wall_in);
    d.wall_net = L(h->wall_net);
    d.roof_in_dir = L(h->roof_in_dir);
    d.wall_in_dir = L(h->wall_in_dir);
    d.roof_sunlit_frac = L1(h->roof_sunlit_frac);
    d.wall_sunlit_frac = L1(h->wall_sunlit_frac);
    d.clear_air_abs = L(h->clear_air_abs);
    d.veg_abs = L(h->veg_abs);
    d.veg_air_abs = L(h->veg_air_abs);
    d.veg_abs_dir = L(h->veg_abs_dir);
    d.veg_sunlit_frac = L1(h->veg_sunlit_frac);
    d.flux_dn_layer_top = L(h->flux_dn_layer_top);
    d.flux_up_layer_top = L(h->flux_up_layer_top);
    d.flux_dn_layer_base = L(h->flux_dn_layer_base);
    d.flux_up_layer_base = L(h->flux_up_layer_base);
    d.flux_dn_dir_layer_top = L(h->flux_dn_dir_layer_top);
    d.flux_dn_dir_layer_base = L(h->flux_dn_dir_layer_base);
  };
  // a normalised flux object that lives on the device only, with the members of `tmpl`
  std::vector<std::pair<double *, size_t>> zeroed;
  auto device_flux = [&](const ssb200_canopy_flux *tmpl, ssb200_canopy_flux &d) {
    d = *tmpl;
    const size_t g = (size_t)tmpl->nspec;
    auto make = [&](double *present, size_t total, bool always_zero) -> double * {
      if (!present) return nullptr;
      double *p = sg.device_only(total);
      // members the kernels may leave untouched start from zero (canopy_flux_type%zero_all, driver:184)
      if (always_zero || !contiguous) zeroed.push_back({p, total});
      return p;
    };
#define SSB_DC(m) d.m = make(tmpl->m, ncol * g, true)
#define SSB_DL(m) d.m = make(tmpl->m, ntot * g, false)
    SSB_DC(ground_dn); SSB_DC(ground_net); SSB_DC(ground_vertical_diff); SSB_DC(top_dn); SSB_DC(top_net);
    SSB_DC(ground_dn_dir); SSB_DC(top_dn_dir);
    d.ground_sunlit_frac = make(tmpl->ground_sunlit_frac, ncol, true);
    SSB_DL(roof_in); SSB_DL(roof_net); SSB_DL(wall_in); SSB_DL(wall_net); SSB_DL(roof_in_dir); SSB_DL(wall_in_dir);
    d.roof_sunlit_frac = make(tmpl->roof_sunlit_frac, ntot, false);
    d.wall_sunlit_frac = make(tmpl->wall_sunlit_frac, ntot, false);
    SSB_DL(clear_air_abs); SSB_DL(veg_abs); SSB_DL(veg_air_abs); SSB_DL(veg_abs_dir);
    d.veg_sunlit_frac = make(tmpl->veg_sunlit_frac, ntot, false);
    SSB_DL(flux_dn_layer_top); SSB_DL(flux_up_layer_top); SSB_DL(flux_dn_layer_base); SSB_DL(flux_up_layer_base);
    SSB_DL(flux_dn_dir_layer_top); SSB_DL(flux_dn_dir_layer_base);
#undef SSB_DC
#undef SSB_DL
  };
  ssb200_canopy_flux d1, d2, d3, d4, dsum_sw, dsum_lw;
  const double *d_top_sw = nullptr, *d_top_dir = nullptr, *d_top_lw = nullptr;
  if (drv) {
    if (config->do_sw) {
      device_flux(sw_flux, d1);
      device_flux(sw_flux, d2);
      stage_flux(sw_flux, dsum_sw);
      d_top_sw = sg.mirror(drv->top_flux_dn_sw, ncol, (size_t)config->nsw, false, true, false);
      d_top_dir = sg.mirror(drv->top_flux_dn_direct_sw, ncol, (size_t)config->nsw, false, true, false);
    }
    if (config->do_lw) {
      device_flux(lw_flux, d3);
      device_flux(lw_flux, d4);
      stage_flux(lw_flux, dsum_lw);
      d_top_lw = sg.mirror(drv->top_flux_dn_lw, ncol, (size_t)config->nlw, false, true, false);
    }
  } else {
    if (config->do_sw) {
      stage_flux(sw_dir, d1);
      stage_flux(sw_diff, d2);
    }
    if (config->do_lw) {
      stage_flux(lw_int, d3);
      stage_flux(lw_norm, d4);
    }
  }
  if (sg.rc) return sg.rc;
  for (const Fill &f : fills)
    if (f.n > 0) {
      k_fill<<<(unsigned)((f.n + 255) / 256), 256, 0, st>>>(f.p, f.v, 0, (long)f.n);
      ++g_launches;
    }
  for (const auto &z : zeroed) SSB_CUDA(cudaMemsetAsync(z.first, 0, z.second * sizeof(double), st));
  ssb::CallArgs ca{config, &dcp, config->do_sw ? &dsw : nullptr, config->do_lw ? &dlw : nullptr, &dbc,
                   config->do_sw ? &d1 : nullptr, config->do_sw ? &d2 : nullptr,
                   config->do_lw ? &d3 : nullptr, config->do_lw ? &d4 : nullptr};
  // Pipeline: the column range is cut into blocks whose input slices are uploaded, solved
  // and downloaded as three overlapping stages (pinned host memory needed for true
  // overlap).  Requires a contiguous packed-layer range; profiling serialises.
  int range[2];
  rc = radsurf_device_locked(cx, ca, c1, c2, st, nullptr, range);
  if (rc) return rc;
  if (drv && cx.lay2col_generation != cx.plan.generation) {
    SSB_CUDA(cx.d_lay2col.reserve(sizeof(int) * (ntot + 1)));
    k_lay2col<<<(unsigned)((ncol + 127) / 128), 128, 0, st>>>((const int *)cx.d_nlay.p, (const int *)cx.d_istart.p,
                                                               (const int *)cx.d_irep.p, (int)ncol, (int *)cx.d_lay2col.p);
    ++g_launches;
    cx.lay2col_generation = cx.plan.generation;
  }
  SSB_CUDA(cudaStreamSynchronize(st));  // plan upload and status reset are visible to every lane
  const size_t total_work = lN + cN;
  int nblk = 1;
  if (cx.pipeline && contiguous && !cx.profiling && total_work >= ((size_t)1 << 17)) {
    // blocks of at least ~30 k columns x 16 layers: smaller launches are bound by launch latency (a rank
    // of an 8-GPU run holds 131,072 columns: 4 blocks, not 16)
    nblk = (int)(total_work >> 19);
    if (nblk < 2) nblk = 2;
    if (nblk > cx.pipeline_max_blocks) nblk = cx.pipeline_max_blocks;
  }
  // three stages on three streams, chained per block by events: uploads in block order on
  // lane 0, kernels on lane 1 (one scratch area with the full budget, reused block after
  // block), downloads on lane 2 - both copy engines stay busy once the pipeline has filled
  for (int l = 0; l < Context::kLanes; ++l)
    if (!cx.lane_stream[l]) SSB_CUDA(cudaStreamCreateWithFlags(&cx.lane_stream[l], cudaStreamNonBlocking));
  cudaStream_t s_up = cx.lane_stream[0], s_run = cx.lane_stream[1], s_down = cx.lane_stream[2];
  if (nblk == 1) s_up = s_down = s_run;
  if ((int)cx.blk_events.size() < 2 * nblk) {
    const size_t old_n = cx.blk_events.size();
    cx.blk_events.resize((size_t)2 * nblk, nullptr);
    for (size_t i = old_n; i < cx.blk_events.size(); ++i)
      SSB_CUDA(cudaEventCreateWithFlags(&cx.blk_events[i], cudaEventDisableTiming));
  }
  // first packed layer of every column >= j (Flat tiles own no layers): block boundaries
  auto layer_begin = [&](int j) -> size_t {
    for (; j < c2; ++j)
      if (cp->i_representation[j] != SSB200_TILE_FLAT && cp->nlay[j] > 0) return (size_t)cp->istartlay[j] - 1;
    return l2;
  };
  // an error return inside the loop must not leave queued copies reading / writing the caller's
  // host arrays after the function has returned (the caller may free them): drain the lanes first
  struct LaneDrain {
    Context &cx;
    bool armed = true;
    ~LaneDrain() {
      if (armed)
        for (int l = 0; l < Context::kLanes; ++l) cudaStreamSynchronize(cx.lane_stream[l]);
    }
  } drain{cx};
  for (int b = 0; b < nblk; ++b) {
    const int cb0 = (c1 - 1) + (int)(((long long)cN * b) / nblk), cb1 = (c1 - 1) + (int)(((long long)cN * (b + 1)) / nblk);
    if (cb1 <= cb0) continue;
    const size_t lb0 = (b == 0) ? l1 : layer_begin(cb0), lb1 = (b == nblk - 1) ? l2 : layer_begin(cb1);
    rc = sg.copy_window(true, (size_t)cb0, (size_t)cb1, lb0, lb1, s_up);
    if (rc) return rc;
    if (nblk > 1) {
      SSB_CUDA(cudaEventRecord(cx.blk_events[2 * b], s_up));
      SSB_CUDA(cudaStreamWaitEvent(s_run, cx.blk_events[2 * b], 0));
    }
    if (drv) {  // input stage of the window: default contact fraction, emission from temperatures
      if (d_vcf_default && lb1 > lb0) {
        k_vcf_default<<<(unsigned)((lb1 - lb0 + 255) / 256), 256, 0, s_run>>>(
            d_vcf_default, dcp.veg_fraction, dcp.building_fraction, config->min_vegetation_fraction, (long)lb0, (long)lb1);
        ++g_launches;
      }
      for (const Emis &e : emis) {
        const long i0 = e.per_layer ? (long)lb0 : (long)cb0, i1 = e.per_layer ? (long)lb1 : (long)cb1;
        if (i1 <= i0) continue;
        k_sigma_t4<<<(unsigned)((i1 - i0 + 255) / 256), 256, 0, s_run>>>(e.out, e.emissivity, e.temperature, 1, i0, i1);
        ++g_launches;
      }
    }
    rc = run_window(cx, ca, 1, cx.budget_doubles, cb0, cb1);
    if (rc) return rc;
    if (drv) {  // scale + sum of the window (driver:250-261), still on the kernel lane
      const long cnt = (long)std::max((size_t)(cb1 - cb0), lb1 - lb0);
      auto scale_sum = [&](ssb200_canopy_flux &o, ssb200_canopy_flux &x, ssb200_canopy_flux &y, const double *fa,
                           const double *fb, const double *fbm, int nspec) {
        FieldList fl;
        collect_fields(&o, &x, &y, fl, true);
        const long nthr = cnt * nspec;
        if (fl.n == 0 || nthr <= 0) return;
        k_scale_sum<<<(unsigned)((nthr + 255) / 256), 256, 0, s_run>>>(fl, fa, fb, fbm, (const int *)cx.d_lay2col.p, nspec,
                                                                       (long)cb0, (long)cb1, (long)lb0, (long)lb1);
        ++g_launches;
      };
      if (config->do_sw) scale_sum(dsum_sw, d1, d2, d_top_dir, d_top_sw, d_top_dir, config->nsw);
      if (config->do_lw) scale_sum(dsum_lw, d3, d4, nullptr, d_top_lw, nullptr, config->nlw);
      SSB_CUDA(cudaGetLastError());
    }
    if (nblk > 1) {
      SSB_CUDA(cudaEventRecord(cx.blk_events[2 * b + 1], s_run));
      SSB_CUDA(cudaStreamWaitEvent(s_down, cx.blk_events[2 * b + 1], 0));
    }
    rc = sg.copy_window(false, (size_t)cb0, (size_t)cb1, lb0, lb1, s_down);
    if (rc) return rc;
  }
  for (int l = 0; l < Context::kLanes; ++l) SSB_CUDA(cudaStreamSynchronize(cx.lane_stream[l]));
  drain.armed = false;
  int status = 0;
  SSB_CUDA(cudaMemcpy(&status, cx.d_status.p, sizeof(int), cudaMemcpyDeviceToHost));
  return status;
}

int ssb200_radsurf(const ssb200_config *config, const ssb200_canopy_properties *cp,
                   const ssb200_sw_spectral_properties *sw, const ssb200_lw_spectral_properties *lw,
                   ssb200_boundary_conds_out *bc, int32_t istartcol, int32_t iendcol, ssb200_canopy_flux *sw_dir,
                   ssb200_canopy_flux *sw_diff, ssb200_canopy_flux *lw_int, ssb200_canopy_flux *lw_norm) {
  return radsurf_host(config, cp, sw, lw, nullptr, bc, istartcol, iendcol, sw_dir, sw_diff, lw_int, lw_norm, nullptr,
                      nullptr);
}

int ssb200_radsurf_fluxes(const ssb200_config *config, const ssb200_canopy_properties *cp,
                          const ssb200_sw_spectral_properties *sw, const ssb200_lw_spectral_properties *lw,
                          const ssb200_driver_inputs *drv, ssb200_boundary_conds_out *bc, int32_t istartcol,
                          int32_t iendcol, ssb200_canopy_flux *sw_flux, ssb200_canopy_flux *lw_flux) {
  if (!drv) return fail(SSB200_ERR_ARG, "driver_inputs must not be NULL");
  return radsurf_host(config, cp, sw, lw, drv, bc, istartcol, iendcol, nullptr, nullptr, nullptr, nullptr, sw_flux,
                      lw_flux);
}

// Single-precision storage variant (-DSINGLE_PRECISION builds of the reference: jprb = real32,
// utilities/parkind1.F90:45-49).  The _sp structs are layout-identical to the double ones (pointers
// and 32-bit integers only), so the body is shared; the Stager widens / rounds on the device.
int ssb200_radsurf_sp(const ssb200_config *config, const ssb200_canopy_properties_sp *cp,
                      const ssb200_sw_spectral_properties_sp *sw, const ssb200_lw_spectral_properties_sp *lw,
                      ssb200_boundary_conds_out_sp *bc, int32_t istartcol, int32_t iendcol, ssb200_canopy_flux_sp *sw_dir,
                      ssb200_canopy_flux_sp *sw_diff, ssb200_canopy_flux_sp *lw_int, ssb200_canopy_flux_sp *lw_norm) {
  static_assert(sizeof(ssb200_canopy_properties_sp) == sizeof(ssb200_canopy_properties) &&
                    sizeof(ssb200_sw_spectral_properties_sp) == sizeof(ssb200_sw_spectral_properties) &&
                    sizeof(ssb200_lw_spectral_properties_sp) == sizeof(ssb200_lw_spectral_properties) &&
                    sizeof(ssb200_canopy_flux_sp) == sizeof(ssb200_canopy_flux) &&
                    sizeof(ssb200_boundary_conds_out_sp) == sizeof(ssb200_boundary_conds_out),
                "single-precision structs must mirror the double-precision ones");
  return radsurf_host(config, reinterpret_cast<const ssb200_canopy_properties *>(cp),
                      reinterpret_cast<const ssb200_sw_spectral_properties *>(sw),
                      reinterpret_cast<const ssb200_lw_spectral_properties *>(lw), nullptr,
                      reinterpret_cast<ssb200_boundary_conds_out *>(bc), istartcol, iendcol,
                      reinterpret_cast<ssb200_canopy_flux *>(sw_dir), reinterpret_cast<ssb200_canopy_flux *>(sw_diff),
                      reinterpret_cast<ssb200_canopy_flux *>(lw_int), reinterpret_cast<ssb200_canopy_flux *>(lw_norm), nullptr,
                      nullptr, true);
}

int64_t ssb200_kernel_launch_count(void) { return (int64_t)g_launches; }

int ssb200_set_profiling(int enable) {
  std::lock_guard<std::mutex> lock(g_mutex);
  g_ctx.profiling = enable != 0;
  return 0;
}

int ssb200_last_kernel_times_ms(double out[5]) {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (int i = 0; i < 5; ++i) out[i] = g_ctx.times_ms[i];
  return 0;
}

int ssb200_last_kernel_counts(int64_t out[5]) {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (int i = 0; i < 5; ++i) out[i] = g_ctx.counts[i];
  return 0;
}

int ssb200_set_option(const char *name, int64_t value) {
  std::lock_guard<std::mutex> lock(g_mutex);
  const std::string n(name ? name : "");
  if (n == "scratch_budget_bytes") {
    g_ctx.budget_doubles = (size_t)value / sizeof(double);
    return 0;
  }
  if (n == "fast_kernels") {
    g_ctx.fast_mode = value != 0;
    return 0;
  }
  if (n == "pipeline") {
    g_ctx.pipeline = value != 0;
    return 0;
  }
  if (n == "sort_columns") {
    g_ctx.sort_columns = value != 0;
    return 0;
  }
  if (n == "sort_group") {
    g_ctx.sort_group = (int)value;
    return 0;
  }
  if (n == "pipeline_max_blocks") {
    g_ctx.pipeline_max_blocks = value < 1 ? 1 : (int)value;
    return 0;
  }
  if (n == "concurrent_passes") {
    g_ctx.concurrent_passes = value != 0;
    return 0;
  }
  if (n == "stage_layers") {
    g_ctx.stage_layers = value != 0;
    return 0;
  }
  if (n == "record_sweeps") {
    g_ctx.sweep_mode = value != 0 ? 1 : 0;
    return 0;
  }
  if (n == "fused_kernels") {  // column-resident kernels where they exist (1 and 2 streams)
    g_ctx.fused_mode = value != 0;
    return 0;
  }
  if (n == "fused_sort") {
    g_ctx.fused_sort = value != 0;
    return 0;
  }
  if (n == "fused_sort_group") {
    g_ctx.fused_sort_group = (int)value;
    return 0;
  }
  if (n == "fused_blocks_per_sm") {
    g_ctx.fused_blocks_per_sm = value < 1 ? 1 : (value > 2 ? 2 : (int)value);
    return 0;
  }
  if (n == "fused_sync") {
    g_ctx.fused_sync = value != 0;
    return 0;
  }
  if (n == "fused_l2_persist") {
    g_ctx.fused_l2_persist = value != 0;
    return 0;
  }
  if (n == "partition_layers") {
    g_ctx.partition = value != 0;
    return 0;
  }
  return fail(SSB200_ERR_ARG, "unknown option " + n);
}

int ssb200_release(void) {
  std::lock_guard<std::mutex> lock(g_mutex);
  Context &cx = g_ctx;
  if (ssb200_device_count() > 0) {
    cudaDeviceSynchronize();
    for (DevBuf *b : {&cx.d_nlay, &cx.d_istart, &cx.d_irep, &cx.d_cols, &cx.d_status, &cx.d_lay2col}) b->release();
    for (int l = 0; l < Context::kLanes; ++l) {
      cx.d_scratch[l].release();
      cx.d_perm[l].release();
      cx.d_sort[l].release();
    }
    for (DevBuf &b : cx.stage) b.release();
  }
  cx.plan = ssb::Plan();
  cx.plan_uploaded = false;
  cx.budget_doubles = 0;
  return 0;
}

int ssb200_canopy_flux_scale_device(ssb200_canopy_flux *flux, const int32_t *nlay, const int32_t *istartlay,
                                    const double *factor, void *stream) {
  int rc = ensure_device();
  if (rc) return rc;
  if (!flux || !nlay || !istartlay || !factor) return fail(SSB200_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(g_mutex);
  Context &cx = g_ctx;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int> lay2col((size_t)flux->ntotlay, 0);
  for (int j = 0; j < flux->ncol; ++j)
    for (int l = 0; l < nlay[j]; ++l) {
      const long il = (long)istartlay[j] - 1 + l;
      if (il < 0 || il >= flux->ntotlay) return fail(SSB200_ERR_SHAPE, "layer range outside 1..ntotlay");
      lay2col[(size_t)il] = j;
    }
  SSB_CUDA(cx.d_lay2col.reserve(sizeof(int) * (lay2col.size() + 1)));
  SSB_CUDA(cudaMemcpyAsync(cx.d_lay2col.p, lay2col.data(), sizeof(int) * lay2col.size(), cudaMemcpyHostToDevice, st));
  SSB_CUDA(cudaStreamSynchronize(st));
  FieldList fl;
  collect_fields(flux, nullptr, nullptr, fl, false);
  const long nmax = (long)(flux->ntotlay > flux->ncol ? flux->ntotlay : flux->ncol) * flux->nspec;
  if (nmax > 0) {
    k_scale<<<(unsigned)((nmax + 255) / 256), 256, 0, st>>>(fl, factor, (const int *)cx.d_lay2col.p, flux->nspec,
                                                            flux->ncol, flux->ntotlay);
    ++g_launches;
    SSB_CUDA(cudaGetLastError());
  }
  return 0;
}

int ssb200_canopy_flux_sum_device(ssb200_canopy_flux *out, const ssb200_canopy_flux *a, const ssb200_canopy_flux *b,
                                  void *stream) {
  int rc = ensure_device();
  if (rc) return rc;
  if (!out || !a || !b) return fail(SSB200_ERR_ARG, "NULL argument");
  FieldList fl;
  collect_fields(out, a, b, fl, true);
  const long nmax = (long)(out->ntotlay > out->ncol ? out->ntotlay : out->ncol) * out->nspec;
  if (nmax > 0) {
    k_sum<<<(unsigned)((nmax + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fl, out->nspec, out->ncol, out->ntotlay);
    ++g_launches;
    SSB_CUDA(cudaGetLastError());
  }
  return 0;
}

int ssb200_canopy_flux_check_device(const ssb200_canopy_flux *flux, const ssb200_canopy_properties *cp,
                                    double *residual, void *stream) {
  int rc = ensure_device();
  if (rc) return rc;
  if (!flux || !cp || !residual) return fail(SSB200_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(g_mutex);
  Context &cx = g_ctx;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ncol = (size_t)cp->ncol;
  SSB_CUDA(cx.d_nlay.reserve(sizeof(int) * (ncol + 1)));
  SSB_CUDA(cx.d_istart.reserve(sizeof(int) * (ncol + 1)));
  SSB_CUDA(cx.d_irep.reserve(sizeof(int) * (ncol + 1)));
  SSB_CUDA(cudaMemcpyAsync(cx.d_nlay.p, cp->nlay, sizeof(int) * ncol, cudaMemcpyHostToDevice, st));
  SSB_CUDA(cudaMemcpyAsync(cx.d_istart.p, cp->istartlay, sizeof(int) * ncol, cudaMemcpyHostToDevice, st));
  SSB_CUDA(cudaMemcpyAsync(cx.d_irep.p, cp->i_representation, sizeof(int) * ncol, cudaMemcpyHostToDevice, st));
  SSB_CUDA(cudaStreamSynchronize(st));
  cx.plan_uploaded = false;  // the index buffers were overwritten
  cx.plan.valid = false;
  if (ncol > 0) {
    k_check<<<(unsigned)((ncol + 127) / 128), 128, 0, st>>>(*flux, (const int *)cx.d_nlay.p, (const int *)cx.d_istart.p,
                                                            (const int *)cx.d_irep.p, (int)ncol, residual);
    ++g_launches;
    SSB_CUDA(cudaGetLastError());
  }
  return 0;
}

int ssb200_calc_simple_spectrum_lw_device(ssb200_lw_spectral_properties *lw, int32_t ncol, int32_t ntotlay,
                                          int32_t istartcol, int32_t iendcol, int32_t ilay1, int32_t ilay2,
                                          const double *ground_temperature, const double *roof_temperature,
                                          const double *wall_temperature, const double *clear_air_temperature,
                                          const double *veg_temperature, const double *veg_air_temperature,
                                          void *stream) {
  int rc = ensure_device();
  if (rc) return rc;
  if (!lw) return fail(SSB200_ERR_ARG, "NULL argument");
  if (lw->nspec > 1 && (clear_air_temperature || veg_temperature || veg_air_temperature))
    return fail(SSB200_ERR_ARG, "Simple longwave spectrum only possible with one input spectral interval");
  const long c0 = (istartcol > 0 ? istartcol : 1) - 1, c1 = iendcol > 0 ? (iendcol < ncol ? iendcol : ncol) : ncol;
  const long l0 = (ilay1 > 0 ? ilay1 : 1) - 1, l1 = ilay2 < ntotlay ? ilay2 : ntotlay;
  cudaStream_t st = (cudaStream_t)stream;
  // (the members are inputs of radsurf, hence const in the struct; this stage fills them)
  auto run = [&](const double *out_c, const double *emis, const double *temp, long i0, long i1, bool need_emis) -> int {
    double *out = const_cast<double *>(out_c);
    if (!temp || i1 <= i0) return 0;
    if (!out || (need_emis && !emis)) return fail(SSB200_ERR_ARG, "temperature given but emission / emissivity array missing");
    k_sigma_t4<<<(unsigned)((i1 - i0 + 255) / 256), 256, 0, st>>>(out, emis, temp, lw->nspec, i0, i1);
    ++g_launches;
    SSB_CUDA(cudaGetLastError());
    return 0;
  };
  if ((rc = run(lw->ground_emission, lw->ground_emissivity, ground_temperature, c0, c1, true))) return rc;
  if ((rc = run(lw->roof_emission, lw->roof_emissivity, roof_temperature, l0, l1, true))) return rc;
  if ((rc = run(lw->wall_emission, lw->wall_emissivity, wall_temperature, l0, l1, true))) return rc;
  if ((rc = run(lw->clear_air_planck, nullptr, clear_air_temperature, l0, l1, false))) return rc;
  if ((rc = run(lw->veg_planck, nullptr, veg_temperature, l0, l1, false))) return rc;
  if ((rc = run(lw->veg_air_planck, nullptr, veg_air_temperature, l0, l1, false))) return rc;
  return 0;
}

double ssb200_measure_fp64_peak_tflops(int iters) {
  if (ensure_device()) return -1.0;
  if (iters <= 0) iters = 1 << 16;
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1.0;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double *out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_fp64_peak<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  ++g_launches;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  return best;
}

}  // extern "C"
