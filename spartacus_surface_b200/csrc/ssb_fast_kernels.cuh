// ssb_fast_kernels.cuh - kernels around the register-resident layer bodies
// (ssb_fast_layer.cuh).  Included by ssb_f_ns*_{sw,lw}.cu with SSB_NS and
// SSB_KIND_SW / SSB_KIND_LW defined.
#pragma once
#include "ssb_fast.cuh"
#define SSB_CAT2(a, b) a##b
#define SSB_CAT(a, b) SSB_CAT2(a, b)
#include "ssb_fast_layer.cuh"
#include "ssb_fast_sweeps.cuh"

namespace ssb {

// dynamic shared-memory limit of a kernel: set once per (kernel instantiation, device), not per launch
#define SSB_SMEM_ONCE(kernel, smem)                                                                    \
  do {                                                                                                 \
    static int ssb_dev_done = -1;                                                                      \
    int ssb_dev = 0;                                                                                   \
    cudaGetDevice(&ssb_dev);                                                                           \
    if (ssb_dev_done != ssb_dev) {                                                                     \
      fast_note(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem))); \
      ssb_dev_done = ssb_dev;                                                                          \
    }                                                                                                  \
  } while (0)

// Where the per-thread slice lives.  Shared memory ([element][thread], conflict-free) as long as two
// blocks of the kernel fit one SM; a slice too large for that (3 regions x 4 streams: 1.4-1.7 KB per
// thread, one 128-thread block per SM) goes to the thread's local memory instead, which keeps two blocks
// per SM and with them the warps that hide the latency of everything else: 102 -> 87 ms per step at 4
// streams (SSB_SLICE_LOCAL_BYTES: the shared-memory bytes per block above which the local variant is
// used; 0 switches it off).
#ifndef SSB_SLICE_LOCAL_BYTES
#define SSB_SLICE_LOCAL_BYTES (112 * 1024)
#endif
SSB_HD constexpr bool slice_is_local(int doubles, int threads) {
  return SSB_SLICE_LOCAL_BYTES > 0 && (long)sizeof(double) * doubles * threads > (long)SSB_SLICE_LOCAL_BYTES;
}

constexpr int kFastBlock = 128;  // sweeps: one thread per (column, interval)
#ifndef SSB_SWEEP_MINB
#define SSB_SWEEP_MINB 2  // resident blocks per SM the sweeps are compiled for (3: 168 registers, 1-1.8 KB of spills: 51.7 -> 68.4 ms)
#endif
// layer kernels: block size and the resident threads per SM that the register budget is
// set for (__launch_bounds__).  Measured on B200: aligning the warps of a block with
// barriers at phase boundaries (one 384-thread block per SM) does not pay.
#ifndef SSB_LAYER_BLOCK
#define SSB_LAYER_BLOCK 128
#endif
#ifndef SSB_LAYER_THREADS
#define SSB_LAYER_THREADS 256
#endif
constexpr int kLayerBlock = SSB_LAYER_BLOCK;
constexpr int kLayerMinB = SSB_LAYER_THREADS / SSB_LAYER_BLOCK;

// First pass over the layer problems of a launch (one thread each): evaluates the layer
// geometry once and writes the geometry block (fast_prepare_level); solves the problems of
// layers without vegetation on the spot (clear region only: order NS, a few hundred flops,
// not worth a launch and a second pass over their inputs); groups the other problems by the
// sub-block of regions they solve, so that every warp of the layer kernels runs one code
// path.  Warp-aggregated append: the order inside a segment is not deterministic, the results
// are (every problem is independent).
// resident blocks per SM the partition pass is compiled for: at 1 and 2 streams 64 registers (4 blocks) hide more
// of its load latency than the 80 it would take (47.4 -> 47.2 ms per step); the 4-stream version needs its 128-162
#ifndef SSB_PARTITION_MINB
#define SSB_PARTITION_MINB (SSB_NS <= 2 ? 4 : 1)
#endif
constexpr int kPartitionBlock = 256;
template <int NREG, int NS, bool LW>
static __global__ void __launch_bounds__(kPartitionBlock, SSB_PARTITION_MINB) k_partition_layers(ClassArgs a, long nt) {
  extern __shared__ double ssb_stack[];
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  int seg = -1;
  if (t < nt) {
    const long width = (long)a.ncols * a.cfg.nspec;
    const int q = (int)(t % width), lev = (int)(t / width);
    seg = fast_prepare_level(a, q, lev);
    if (NREG > 1 && seg == 1) {
      const StateMem st{ssb_stack + threadIdx.x, kPartitionBlock};
      if (LW)
        fast_layer_problem_lw_impl<NREG, NS, 1>(a, q, lev, st);
      else
        fast_layer_problem_sw_seg<NREG, NS, 1>(a, q, lev, st);
      seg = -1;
    }
  }
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int k = 0; k < 3; k += 2) {
    const unsigned mask = __ballot_sync(0xffffffffu, seg == k);
    if (mask == 0u) continue;
    int base = 0;
    const int leader = __ffs(mask) - 1;
    if ((int)lane == leader) base = atomicAdd(&a.perm_count[k], __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (seg == k) a.perm[(size_t)k * (size_t)nt + base + __popc(mask & ((1u << lane) - 1u))] = (int)t;
  }
}
template <int NREG, int NS, bool LW>
static void launch_partition_layers(const ClassArgs &a, long nt, cudaStream_t st) {
  const size_t smem = sizeof(double) * kPartitionBlock *
                      (LW ? LayerStack<1, NS>::lw_doubles : LayerStack<1, NS>::sw_doubles);
  if (smem > 48 * 1024) SSB_SMEM_ONCE((k_partition_layers<NREG, NS, LW>), smem);
  fast_note(cudaMemsetAsync(a.perm_count, 0, 3 * sizeof(int), st));
  k_partition_layers<NREG, NS, LW><<<(unsigned)((nt + kPartitionBlock - 1) / kPartitionBlock), kPartitionBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
}

#ifdef SSB_KIND_SW
template <int NREG, int NS, int SEG>
__global__ void __launch_bounds__(kLayerBlock, kLayerMinB) k_fast_layer_sw_seg(ClassArgs a, long nt) {
  extern __shared__ double ssb_stack[];  // [stack element][thread]: conflict-free per-thread slices
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= a.perm_count[SEG]) return;
  const long id = a.perm[(size_t)SEG * (size_t)nt + t];
  const long width = (long)a.ncols * a.cfg.nspec;
  constexpr int NR = (SEG == 0) ? NREG : (SEG == 1 ? 1 : (NREG > 1 ? NREG - 1 : 1));
  if constexpr (slice_is_local(LayerStack<NR, NS>::sw_doubles, kLayerBlock)) {
    double slice[LayerStack<NR, NS>::sw_doubles];
    fast_layer_problem_sw_seg<NREG, NS, SEG>(a, (int)(id % width), (int)(id / width), StateMem{slice, 1});
  } else {
    const StateMem st{ssb_stack + threadIdx.x, kLayerBlock};
    fast_layer_problem_sw_seg<NREG, NS, SEG>(a, (int)(id % width), (int)(id / width), st);
  }
}
template <int NREG, int NS, int SEG>
static void launch_fast_layer_sw_seg(const ClassArgs &a, long nt, unsigned grid, cudaStream_t st) {
  constexpr int NR = (SEG == 0) ? NREG : (SEG == 1 ? 1 : (NREG > 1 ? NREG - 1 : 1));
  const size_t smem = slice_is_local(LayerStack<NR, NS>::sw_doubles, kLayerBlock)
                          ? 0
                          : sizeof(double) * LayerStack<NR, NS>::sw_doubles * kLayerBlock;
  if (smem > 0) SSB_SMEM_ONCE((k_fast_layer_sw_seg<NREG, NS, SEG>), smem);
  k_fast_layer_sw_seg<NREG, NS, SEG><<<grid, kLayerBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
}
template <int NREG, int NS>
static void launch_fast_layer_sw(const ClassArgs &a, long nt, cudaStream_t st) {
  const unsigned grid = (unsigned)((nt + kLayerBlock - 1) / kLayerBlock);
  launch_partition_layers<NREG, NS, false>(a, nt, st);
  launch_fast_layer_sw_seg<NREG, NS, 0>(a, nt, grid, st);
  if (NREG > 1) launch_fast_layer_sw_seg<NREG, NS, 2>(a, nt, grid, st);
}
bool SSB_CAT(fast_layer_sw_ns, SSB_NS)(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS || a.perm == nullptr) return false;
  switch (a.cfg.nreg) {
    case 1: launch_fast_layer_sw<1, SSB_NS>(a, nt, st); return true;
    case 2: launch_fast_layer_sw<2, SSB_NS>(a, nt, st); return true;
    case 3: launch_fast_layer_sw<3, SSB_NS>(a, nt, st); return true;
    default: return false;
  }
}
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFastBlock, SSB_SWEEP_MINB) k_fast_sweeps_sw(ClassArgs a, long nt) {
  extern __shared__ double ssb_state[];  // [state element][thread]: conflict-free per-thread slices
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  if constexpr (slice_is_local(SwSweepLayout<NREG, NS, URBAN>::state_doubles, kFastBlock)) {
    double slice[SwSweepLayout<NREG, NS, URBAN>::state_doubles];
    fast_column_sweeps_sw<NREG, NS, URBAN>(a, (int)t, StateMem{slice, 1});
  } else {
    const StateMem st{ssb_state + threadIdx.x, kFastBlock};
    fast_column_sweeps_sw<NREG, NS, URBAN>(a, (int)t, st);
  }
}
// (3 or 4 blocks per SM at 168 / 128 registers were measured: the spills cost more than the
// extra warps hide)
template <int NREG, int NS, bool URBAN>
static void launch_fast_sweeps_sw(const ClassArgs &a, long nt, cudaStream_t st) {
  const unsigned grid = (unsigned)((nt + kFastBlock - 1) / kFastBlock);
  const size_t smem = slice_is_local(SwSweepLayout<NREG, NS, URBAN>::state_doubles, kFastBlock)
                          ? 0
                          : sizeof(double) * SwSweepLayout<NREG, NS, URBAN>::state_doubles * kFastBlock;
  if (smem > 0) SSB_SMEM_ONCE((k_fast_sweeps_sw<NREG, NS, URBAN>), smem);
  k_fast_sweeps_sw<NREG, NS, URBAN><<<grid, kFastBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
}
bool SSB_CAT(fast_sweeps_sw_ns, SSB_NS)(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS) return false;
  switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
    case 2: launch_fast_sweeps_sw<1, SSB_NS, false>(a, nt, st); return true;
    case 3: launch_fast_sweeps_sw<1, SSB_NS, true>(a, nt, st); return true;
    case 4: launch_fast_sweeps_sw<2, SSB_NS, false>(a, nt, st); return true;
    case 5: launch_fast_sweeps_sw<2, SSB_NS, true>(a, nt, st); return true;
    case 6: launch_fast_sweeps_sw<3, SSB_NS, false>(a, nt, st); return true;
    case 7: launch_fast_sweeps_sw<3, SSB_NS, true>(a, nt, st); return true;
    default: return false;
  }
}
#endif

#ifdef SSB_KIND_LW
template <int NREG, int NS, int SEG>
__global__ void __launch_bounds__(kLayerBlock, kLayerMinB) k_fast_layer_lw_seg(ClassArgs a, long nt) {
  extern __shared__ double ssb_stack[];  // [stack element][thread]: conflict-free per-thread slices
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= a.perm_count[SEG]) return;
  const long id = a.perm[(size_t)SEG * (size_t)nt + t];
  const long width = (long)a.ncols * a.cfg.nspec;
  constexpr int NR = (SEG == 0) ? NREG : (SEG == 1 ? 1 : (NREG > 1 ? NREG - 1 : 1));
  if constexpr (slice_is_local(LayerStack<NR, NS>::lw_doubles, kLayerBlock)) {
    double slice[LayerStack<NR, NS>::lw_doubles];
    fast_layer_problem_lw_impl<NREG, NS, SEG>(a, (int)(id % width), (int)(id / width), StateMem{slice, 1});
  } else {
    const StateMem st{ssb_stack + threadIdx.x, kLayerBlock};
    fast_layer_problem_lw_impl<NREG, NS, SEG>(a, (int)(id % width), (int)(id / width), st);
  }
}
template <int NREG, int NS, int SEG>
static void launch_fast_layer_lw_seg(const ClassArgs &a, long nt, unsigned grid, cudaStream_t st) {
  constexpr int NR = (SEG == 0) ? NREG : (SEG == 1 ? 1 : (NREG > 1 ? NREG - 1 : 1));
  const size_t smem = slice_is_local(LayerStack<NR, NS>::lw_doubles, kLayerBlock)
                          ? 0
                          : sizeof(double) * LayerStack<NR, NS>::lw_doubles * kLayerBlock;
  if (smem > 0) SSB_SMEM_ONCE((k_fast_layer_lw_seg<NREG, NS, SEG>), smem);
  k_fast_layer_lw_seg<NREG, NS, SEG><<<grid, kLayerBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
}
template <int NREG, int NS>
static void launch_fast_layer_lw(const ClassArgs &a, long nt, cudaStream_t st) {
  const unsigned grid = (unsigned)((nt + kLayerBlock - 1) / kLayerBlock);
  launch_partition_layers<NREG, NS, true>(a, nt, st);
  launch_fast_layer_lw_seg<NREG, NS, 0>(a, nt, grid, st);
  if (NREG > 1) launch_fast_layer_lw_seg<NREG, NS, 2>(a, nt, grid, st);
}
bool SSB_CAT(fast_layer_lw_ns, SSB_NS)(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS || a.perm == nullptr) return false;
  switch (a.cfg.nreg) {
    case 1: launch_fast_layer_lw<1, SSB_NS>(a, nt, st); return true;
    case 2: launch_fast_layer_lw<2, SSB_NS>(a, nt, st); return true;
    case 3: launch_fast_layer_lw<3, SSB_NS>(a, nt, st); return true;
    default: return false;
  }
}
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFastBlock, SSB_SWEEP_MINB) k_fast_sweeps_lw(ClassArgs a, long nt) {
  extern __shared__ double ssb_state[];  // [state element][thread]: conflict-free per-thread slices
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= nt) return;
  if constexpr (slice_is_local(LwSweepLayout<NREG, NS, URBAN>::state_doubles, kFastBlock)) {
    double slice[LwSweepLayout<NREG, NS, URBAN>::state_doubles];
    fast_column_sweeps_lw<NREG, NS, URBAN>(a, (int)t, StateMem{slice, 1});
  } else {
    const StateMem st{ssb_state + threadIdx.x, kFastBlock};
    fast_column_sweeps_lw<NREG, NS, URBAN>(a, (int)t, st);
  }
}
// (3 or 4 blocks per SM at 168 / 128 registers were measured: the spills cost more than the
// extra warps hide)
template <int NREG, int NS, bool URBAN>
static void launch_fast_sweeps_lw(const ClassArgs &a, long nt, cudaStream_t st) {
  const unsigned grid = (unsigned)((nt + kFastBlock - 1) / kFastBlock);
  const size_t smem = slice_is_local(LwSweepLayout<NREG, NS, URBAN>::state_doubles, kFastBlock)
                          ? 0
                          : sizeof(double) * LwSweepLayout<NREG, NS, URBAN>::state_doubles * kFastBlock;
  if (smem > 0) SSB_SMEM_ONCE((k_fast_sweeps_lw<NREG, NS, URBAN>), smem);
  k_fast_sweeps_lw<NREG, NS, URBAN><<<grid, kFastBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
}
bool SSB_CAT(fast_sweeps_lw_ns, SSB_NS)(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS) return false;
  switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
    case 2: launch_fast_sweeps_lw<1, SSB_NS, false>(a, nt, st); return true;
    case 3: launch_fast_sweeps_lw<1, SSB_NS, true>(a, nt, st); return true;
    case 4: launch_fast_sweeps_lw<2, SSB_NS, false>(a, nt, st); return true;
    case 5: launch_fast_sweeps_lw<2, SSB_NS, true>(a, nt, st); return true;
    case 6: launch_fast_sweeps_lw<3, SSB_NS, false>(a, nt, st); return true;
    case 7: launch_fast_sweeps_lw<3, SSB_NS, true>(a, nt, st); return true;
    default: return false;
  }
}
#endif

}  // namespace ssb
