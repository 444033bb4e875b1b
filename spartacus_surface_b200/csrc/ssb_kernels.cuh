// ssb_kernels.cuh - generic kernels: one CUDA thread per problem, thin index
// wrappers around the bodies of ssb_solver.cuh.  Included by ssb_k_ns*_*.cu
// with SSB_NS (stream capacity) and SSB_KIND_SW / SSB_KIND_LW defined.
#pragma once
#include "ssb_launch.hpp"

namespace ssb {

constexpr int kGenericBlock = 64;

#ifdef SSB_KIND_SW
template <int NS>
__global__ void __launch_bounds__(kGenericBlock) k_layer_sw(ClassArgs a, long nt) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const long width = (long)a.ncols * a.cfg.nspec;
  layer_problem_sw<NS>(a, (int)(t % width), (int)(t / width));
}
template <int NS>
__global__ void __launch_bounds__(kGenericBlock) k_sweeps_sw(ClassArgs a, long nt) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  column_sweeps_sw<NS>(a, (int)t);
}
template <>
void launch_layer_sw<SSB_NS>(const ClassArgs &a, long nt, cudaStream_t st) {
  k_layer_sw<SSB_NS><<<(unsigned)((nt + kGenericBlock - 1) / kGenericBlock), kGenericBlock, 0, st>>>(a, nt);
}
template <>
void launch_sweeps_sw<SSB_NS>(const ClassArgs &a, long nt, cudaStream_t st) {
  k_sweeps_sw<SSB_NS><<<(unsigned)((nt + kGenericBlock - 1) / kGenericBlock), kGenericBlock, 0, st>>>(a, nt);
}
#endif

#ifdef SSB_KIND_LW
template <int NS>
__global__ void __launch_bounds__(kGenericBlock) k_layer_lw(ClassArgs a, long nt) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const long width = (long)a.ncols * a.cfg.nspec;
  layer_problem_lw<NS>(a, (int)(t % width), (int)(t / width));
}
template <int NS>
__global__ void __launch_bounds__(kGenericBlock) k_sweeps_lw(ClassArgs a, long nt) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  column_sweeps_lw<NS>(a, (int)t);
}
template <>
void launch_layer_lw<SSB_NS>(const ClassArgs &a, long nt, cudaStream_t st) {
  k_layer_lw<SSB_NS><<<(unsigned)((nt + kGenericBlock - 1) / kGenericBlock), kGenericBlock, 0, st>>>(a, nt);
}
template <>
void launch_sweeps_lw<SSB_NS>(const ClassArgs &a, long nt, cudaStream_t st) {
  k_sweeps_lw<SSB_NS><<<(unsigned)((nt + kGenericBlock - 1) / kGenericBlock), kGenericBlock, 0, st>>>(a, nt);
}
#endif

}  // namespace ssb
