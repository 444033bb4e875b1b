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
}  // namespace ssb
