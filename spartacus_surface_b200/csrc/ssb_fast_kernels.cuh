// ssb_fast_kernels.cuh - kernels around the register-resident layer bodies
// (ssb_fast_layer.cuh).  Included by ssb_f_ns*_{sw,lw}.cu with SSB_NS and
// SSB_KIND_SW / SSB_KIND_LW defined.
#pragma once
#include "ssb_fast.cuh"
#include "ssb_fast_layer.cuh"
#include "ssb_fast_sweeps.cuh"

namespace ssb {

constexpr int kFastBlock = 128;

#ifdef SSB_KIND_SW
template <int NREG, int NS>
__global__ void __launch_bounds__(kFastBlock) k_fast_layer_sw(ClassArgs a, long nt) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const long width = (long)a.ncols * a.cfg.nspec;
  fast_layer_problem_sw<NREG, NS>(a, (int)(t % width), (int)(t / width));
}
template <>
bool fast_layer_sw<SSB_NS>(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS) return false;
  const unsigned grid = (unsigned)((nt + kFastBlock - 1) / kFastBlock);
  switch (a.cfg.nreg) {
    case 1: k_fast_layer_sw<1, SSB_NS><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 2: k_fast_layer_sw<2, SSB_NS><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 3: k_fast_layer_sw<3, SSB_NS><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    default: return false;
  }
}
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFastBlock) k_fast_sweeps_sw(ClassArgs a, long nt) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  fast_column_sweeps_sw<NREG, NS, URBAN>(a, (int)t);
}
template <>
bool fast_sweeps_sw<SSB_NS>(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS) return false;
  const unsigned grid = (unsigned)((nt + kFastBlock - 1) / kFastBlock);
  const int key = a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0);
  switch (key) {
    case 2: k_fast_sweeps_sw<1, SSB_NS, false><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 3: k_fast_sweeps_sw<1, SSB_NS, true><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 4: k_fast_sweeps_sw<2, SSB_NS, false><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 5: k_fast_sweeps_sw<2, SSB_NS, true><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 6: k_fast_sweeps_sw<3, SSB_NS, false><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 7: k_fast_sweeps_sw<3, SSB_NS, true><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    default: return false;
  }
}
#endif

#ifdef SSB_KIND_LW
template <int NREG, int NS>
__global__ void __launch_bounds__(kFastBlock) k_fast_layer_lw(ClassArgs a, long nt) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const long width = (long)a.ncols * a.cfg.nspec;
  fast_layer_problem_lw<NREG, NS>(a, (int)(t % width), (int)(t / width));
}
template <>
bool fast_layer_lw<SSB_NS>(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS) return false;
  const unsigned grid = (unsigned)((nt + kFastBlock - 1) / kFastBlock);
  switch (a.cfg.nreg) {
    case 1: k_fast_layer_lw<1, SSB_NS><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 2: k_fast_layer_lw<2, SSB_NS><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 3: k_fast_layer_lw<3, SSB_NS><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    default: return false;
  }
}
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFastBlock) k_fast_sweeps_lw(ClassArgs a, long nt) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  fast_column_sweeps_lw<NREG, NS, URBAN>(a, (int)t);
}
template <>
bool fast_sweeps_lw<SSB_NS>(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS) return false;
  const unsigned grid = (unsigned)((nt + kFastBlock - 1) / kFastBlock);
  const int key = a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0);
  switch (key) {
    case 2: k_fast_sweeps_lw<1, SSB_NS, false><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 3: k_fast_sweeps_lw<1, SSB_NS, true><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 4: k_fast_sweeps_lw<2, SSB_NS, false><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 5: k_fast_sweeps_lw<2, SSB_NS, true><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 6: k_fast_sweeps_lw<3, SSB_NS, false><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    case 7: k_fast_sweeps_lw<3, SSB_NS, true><<<grid, kFastBlock, 0, st>>>(a, nt); return true;
    default: return false;
  }
}
#endif

}  // namespace ssb
