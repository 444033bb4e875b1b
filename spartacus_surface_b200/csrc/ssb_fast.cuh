// ssb_fast.cuh - hooks of the register-resident ("fast") kernels.  A hook
// returns false when it has no kernel for the requested shape, in which case
// the generic one-thread-per-problem kernel runs.  Specialisations for the
// stream counts that have fast kernels are defined in ssb_f_ns*_*.cu.
#pragma once
#include <cuda_runtime.h>

#include "ssb_solver.cuh"

namespace ssb {
template <int NS>
inline bool fast_layer_sw(const ClassArgs &, long, cudaStream_t) {
  return false;
}
template <int NS>
inline bool fast_layer_lw(const ClassArgs &, long, cudaStream_t) {
  return false;
}
template <int NS>
inline bool fast_sweeps_sw(const ClassArgs &, long, cudaStream_t) {
  return false;
}
template <int NS>
inline bool fast_sweeps_lw(const ClassArgs &, long, cudaStream_t) {
  return false;
}
template <> bool fast_sweeps_sw<1>(const ClassArgs &, long, cudaStream_t);
template <> bool fast_sweeps_sw<2>(const ClassArgs &, long, cudaStream_t);
template <> bool fast_sweeps_lw<1>(const ClassArgs &, long, cudaStream_t);
template <> bool fast_sweeps_lw<2>(const ClassArgs &, long, cudaStream_t);
template <> bool fast_layer_sw<1>(const ClassArgs &, long, cudaStream_t);
template <> bool fast_layer_sw<2>(const ClassArgs &, long, cudaStream_t);
template <> bool fast_layer_lw<1>(const ClassArgs &, long, cudaStream_t);
template <> bool fast_layer_lw<2>(const ClassArgs &, long, cudaStream_t);
}  // namespace ssb
