// ssb_fast.cuh - sub-warp kernels for the hot configurations (filled in by
// ssb_fast_impl.cuh); every hook returns false when it has no kernel for the
// requested shape, in which case the generic one-thread-per-problem kernel runs.
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
}  // namespace ssb
