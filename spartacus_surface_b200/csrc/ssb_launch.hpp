// ssb_launch.hpp - launchers of the generic kernels.  The kernels are
// instantiated per stream capacity NS in separate translation units
// (ssb_k_ns*_{sw,lw}.cu) so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include "ssb_solver.cuh"

namespace ssb {
template <int NS> void launch_layer_sw(const ClassArgs &a, long nt, cudaStream_t st);
template <int NS> void launch_sweeps_sw(const ClassArgs &a, long nt, cudaStream_t st);
template <int NS> void launch_layer_lw(const ClassArgs &a, long nt, cudaStream_t st);
template <int NS> void launch_sweeps_lw(const ClassArgs &a, long nt, cudaStream_t st);
}  // namespace ssb
