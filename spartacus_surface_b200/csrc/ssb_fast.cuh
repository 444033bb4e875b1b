// ssb_fast.cuh - hooks of the register-resident ("fast") kernels.  A hook returns false when
// it has no kernel for the requested shape, in which case the generic one-thread-per-problem
// kernel runs.  The kernels of one stream count are built in their own translation units
// (ssb_f_ns*_{sw,lw}.cu: fast_*_ns<k>); the hooks are indexed by the stream CAPACITY the
// dispatcher instantiates (1, 2, 4, 8 streams).
#pragma once
#include <cuda_runtime.h>

#include "ssb_solver.cuh"

namespace ssb {
// first error returned by a runtime call of the launchers (attribute settings and memsets may
// report - and thereby clear - the error of an earlier launch); read by the backend
extern cudaError_t g_fast_error;
inline void fast_note(cudaError_t e) {
  if (e != cudaSuccess && g_fast_error == cudaSuccess) g_fast_error = e;
}
#define SSB_FAST_DECL(k)                                                   \
  bool fast_layer_sw_ns##k(const ClassArgs &, long, cudaStream_t);        \
  bool fast_layer_lw_ns##k(const ClassArgs &, long, cudaStream_t);        \
  bool fast_sweeps_sw_ns##k(const ClassArgs &, long, cudaStream_t);       \
  bool fast_sweeps_lw_ns##k(const ClassArgs &, long, cudaStream_t);
SSB_FAST_DECL(1)
SSB_FAST_DECL(2)
SSB_FAST_DECL(3)
SSB_FAST_DECL(4)
#undef SSB_FAST_DECL

#define SSB_FAST_HOOK(name)                                                                  \
  template <int CAP>                                                                         \
  inline bool name(const ClassArgs &a, long nt, cudaStream_t st) {                           \
    if (CAP == 1) return name##_ns1(a, nt, st);                                              \
    if (CAP == 2) return name##_ns2(a, nt, st);                                              \
    if (CAP == 4) return a.cfg.ns == 3 ? name##_ns3(a, nt, st) : name##_ns4(a, nt, st);      \
    return false; /* 8 streams: generic kernels */                                           \
  }
SSB_FAST_HOOK(fast_layer_sw)
SSB_FAST_HOOK(fast_layer_lw)
SSB_FAST_HOOK(fast_sweeps_sw)
SSB_FAST_HOOK(fast_sweeps_lw)
#undef SSB_FAST_HOOK

// column-resident kernels (ssb_g_ns*_{sw,lw}.cu): 1 and 2 streams
#define SSB_FUSED_DECL(k)                                                                \
  bool fused_shape_sw_ns##k(const SolveCfg &, int *, int *, int *);                     \
  bool fused_shape_lw_ns##k(const SolveCfg &, int *, int *, int *);                     \
  bool fused_sw_ns##k(const ClassArgs &, long, int, cudaStream_t);                      \
  bool fused_lw_ns##k(const ClassArgs &, long, int, cudaStream_t);                      \
  bool records_sw_ns##k(const ClassArgs &, long, cudaStream_t);                         \
  bool records_lw_ns##k(const ClassArgs &, long, cudaStream_t);
SSB_FUSED_DECL(1)
SSB_FUSED_DECL(2)
#undef SSB_FUSED_DECL
inline bool fused_shape(const SolveCfg &c, bool lw, int *pe, int *oe, int *geo) {
  if (lw) return c.ns == 1 ? fused_shape_lw_ns1(c, pe, oe, geo) : (c.ns == 2 ? fused_shape_lw_ns2(c, pe, oe, geo) : false);
  return c.ns == 1 ? fused_shape_sw_ns1(c, pe, oe, geo) : (c.ns == 2 ? fused_shape_sw_ns2(c, pe, oe, geo) : false);
}
inline bool fused_launch(const ClassArgs &a, bool lw, long nt, int grid, cudaStream_t st) {
  if (lw) return a.cfg.ns == 1 ? fused_lw_ns1(a, nt, grid, st) : (a.cfg.ns == 2 ? fused_lw_ns2(a, nt, grid, st) : false);
  return a.cfg.ns == 1 ? fused_sw_ns1(a, nt, grid, st) : (a.cfg.ns == 2 ? fused_sw_ns2(a, nt, grid, st) : false);
}
// record sweeps (upward pass over the layer scratch -> operator records -> downward pass)
inline bool records_launch(const ClassArgs &a, bool lw, long nt, cudaStream_t st) {
  if (lw) return a.cfg.ns == 1 ? records_lw_ns1(a, nt, st) : (a.cfg.ns == 2 ? records_lw_ns2(a, nt, st) : false);
  return a.cfg.ns == 1 ? records_sw_ns1(a, nt, st) : (a.cfg.ns == 2 ? records_sw_ns2(a, nt, st) : false);
}
}  // namespace ssb
