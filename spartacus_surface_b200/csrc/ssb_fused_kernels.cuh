// ssb_fused_kernels.cuh - kernels around the column-resident bodies (ssb_fused.cuh).  Included by
// ssb_g_ns*_{sw,lw}.cu with SSB_NS, SSB_FUSED_TU and SSB_KIND_SW / SSB_KIND_LW defined.
//
// Persistent grid: as many blocks as the device holds at once (2 per SM); a block walks over
// tiles of 128 (column, interval) problems and owns ONE private scratch tile for all of them, so
// the layer matrices of the ~38 k problems in flight (75 MB at 2 streams) stay in the 126 MB L2.
#pragma once
#include "ssb_fast.cuh"
#define SSB_CAT2(a, b) a##b
#define SSB_CAT(a, b) SSB_CAT2(a, b)
#include "ssb_fused.cuh"

namespace ssb {

constexpr int kFusedBlock = kScratchTile;  // one thread per problem of a tile
#ifndef SSB_REC_DOWN_BLOCKS
#define SSB_REC_DOWN_BLOCKS 4  // resident blocks per SM the downward record kernel is compiled for
#endif

#ifdef SSB_KIND_SW
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFusedBlock, 2) k_fused_sw(ClassArgs a, long nt) {
  extern __shared__ double ssb_stack[];  // [element][thread]: conflict-free per-thread slices
  const StateMem st{ssb_stack + threadIdx.x, kFusedBlock};
  const long ntiles = (nt + kFusedBlock - 1) / kFusedBlock;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long q = tile * kFusedBlock + threadIdx.x;
    fused_column_sw<NREG, NS, URBAN>(a, (int)q, q < nt, st);
  }
}
template <int NREG, int NS, bool URBAN>
static void launch_fused_sw(const ClassArgs &a, long nt, int grid, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = sizeof(double) * SwFused<NREG, NS, URBAN>::smem_doubles * kFusedBlock;
  if (!configured) {
    fast_note(cudaFuncSetAttribute(k_fused_sw<NREG, NS, URBAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  k_fused_sw<NREG, NS, URBAN><<<grid, kFusedBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
}
// sizes of the two scratch areas: private tile elements, operator record elements
bool SSB_CAT(fused_shape_sw_ns, SSB_NS)(const SolveCfg &c, int *private_elems, int *op_elems, int *geo_first) {
  if (c.ns != SSB_NS) return false;
#define SSB_SHAPE(NR, UR)                                             \
  {                                                                    \
    *private_elems = SwFused<NR, SSB_NS, UR>::private_elems;          \
    *op_elems = SwFused<NR, SSB_NS, UR>::op_elems;                    \
    *geo_first = SwSweepLayout<NR, SSB_NS, UR>::oGeo;                 \
    return true;                                                       \
  }
  switch (c.nreg * 2 + (c.urban ? 1 : 0)) {
    case 2: SSB_SHAPE(1, false)
    case 3: SSB_SHAPE(1, true)
    case 4: SSB_SHAPE(2, false)
    case 5: SSB_SHAPE(2, true)
    case 6: SSB_SHAPE(3, false)
    case 7: SSB_SHAPE(3, true)
    default: return false;
  }
#undef SSB_SHAPE
}
bool SSB_CAT(fused_sw_ns, SSB_NS)(const ClassArgs &a, long nt, int grid, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS || !a.fused) return false;
  switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
    case 2: launch_fused_sw<1, SSB_NS, false>(a, nt, grid, st); return true;
    case 3: launch_fused_sw<1, SSB_NS, true>(a, nt, grid, st); return true;
    case 4: launch_fused_sw<2, SSB_NS, false>(a, nt, grid, st); return true;
    case 5: launch_fused_sw<2, SSB_NS, true>(a, nt, grid, st); return true;
    case 6: launch_fused_sw<3, SSB_NS, false>(a, nt, grid, st); return true;
    case 7: launch_fused_sw<3, SSB_NS, true>(a, nt, grid, st); return true;
    default: return false;
  }
}
// ---- record sweeps after the split layer kernels (ssb_fused.cuh, MODE 1 / 2) ----------------
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFusedBlock, 2) k_rec_up_sw(ClassArgs a, long nt) {
  extern __shared__ double ssb_stack[];
  const long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (q >= nt) return;
  const StateMem st{ssb_stack + threadIdx.x, kFusedBlock};
  fused_column_sw<NREG, NS, URBAN, 1>(a, (int)q, true, st);
}
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFusedBlock, SSB_REC_DOWN_BLOCKS) k_rec_down_sw(ClassArgs a, long nt) {
  const long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (q >= nt) return;
  fused_column_sw<NREG, NS, URBAN, 2>(a, (int)q, true, StateMem{nullptr, 0});
}
template <int NREG, int NS, bool URBAN>
static void launch_records_sw(const ClassArgs &a, long nt, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = sizeof(double) * SwFused<NREG, NS, URBAN>::step_doubles * kFusedBlock;
  if (!configured) {
    fast_note(cudaFuncSetAttribute(k_rec_up_sw<NREG, NS, URBAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const unsigned grid = (unsigned)((nt + kFusedBlock - 1) / kFusedBlock);
  k_rec_up_sw<NREG, NS, URBAN><<<grid, kFusedBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
  k_rec_down_sw<NREG, NS, URBAN><<<grid, kFusedBlock, 0, st>>>(a, nt);
  fast_note(cudaGetLastError());
}
bool SSB_CAT(records_sw_ns, SSB_NS)(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS || a.fused) return false;
  switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
    case 2: launch_records_sw<1, SSB_NS, false>(a, nt, st); return true;
    case 3: launch_records_sw<1, SSB_NS, true>(a, nt, st); return true;
    case 4: launch_records_sw<2, SSB_NS, false>(a, nt, st); return true;
    case 5: launch_records_sw<2, SSB_NS, true>(a, nt, st); return true;
    case 6: launch_records_sw<3, SSB_NS, false>(a, nt, st); return true;
    case 7: launch_records_sw<3, SSB_NS, true>(a, nt, st); return true;
    default: return false;
  }
}
#endif

#ifdef SSB_KIND_LW
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFusedBlock, 2) k_fused_lw(ClassArgs a, long nt) {
  extern __shared__ double ssb_stack[];
  const StateMem st{ssb_stack + threadIdx.x, kFusedBlock};
  const long ntiles = (nt + kFusedBlock - 1) / kFusedBlock;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long q = tile * kFusedBlock + threadIdx.x;
    fused_column_lw<NREG, NS, URBAN>(a, (int)q, q < nt, st);
  }
}
template <int NREG, int NS, bool URBAN>
static void launch_fused_lw(const ClassArgs &a, long nt, int grid, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = sizeof(double) * LwFused<NREG, NS, URBAN>::smem_doubles * kFusedBlock;
  if (!configured) {
    fast_note(cudaFuncSetAttribute(k_fused_lw<NREG, NS, URBAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  k_fused_lw<NREG, NS, URBAN><<<grid, kFusedBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
}
bool SSB_CAT(fused_shape_lw_ns, SSB_NS)(const SolveCfg &c, int *private_elems, int *op_elems, int *geo_first) {
  if (c.ns != SSB_NS) return false;
#define SSB_SHAPE(NR, UR)                                             \
  {                                                                    \
    *private_elems = LwFused<NR, SSB_NS, UR>::private_elems;          \
    *op_elems = LwFused<NR, SSB_NS, UR>::op_elems;                    \
    *geo_first = LwSweepLayout<NR, SSB_NS, UR>::oGeo;                 \
    return true;                                                       \
  }
  switch (c.nreg * 2 + (c.urban ? 1 : 0)) {
    case 2: SSB_SHAPE(1, false)
    case 3: SSB_SHAPE(1, true)
    case 4: SSB_SHAPE(2, false)
    case 5: SSB_SHAPE(2, true)
    case 6: SSB_SHAPE(3, false)
    case 7: SSB_SHAPE(3, true)
    default: return false;
  }
#undef SSB_SHAPE
}
bool SSB_CAT(fused_lw_ns, SSB_NS)(const ClassArgs &a, long nt, int grid, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS || !a.fused) return false;
  switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
    case 2: launch_fused_lw<1, SSB_NS, false>(a, nt, grid, st); return true;
    case 3: launch_fused_lw<1, SSB_NS, true>(a, nt, grid, st); return true;
    case 4: launch_fused_lw<2, SSB_NS, false>(a, nt, grid, st); return true;
    case 5: launch_fused_lw<2, SSB_NS, true>(a, nt, grid, st); return true;
    case 6: launch_fused_lw<3, SSB_NS, false>(a, nt, grid, st); return true;
    case 7: launch_fused_lw<3, SSB_NS, true>(a, nt, grid, st); return true;
    default: return false;
  }
}
// ---- record sweeps after the split layer kernels (ssb_fused.cuh, MODE 1 / 2) ----------------
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFusedBlock, 2) k_rec_up_lw(ClassArgs a, long nt) {
  extern __shared__ double ssb_stack[];
  const long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (q >= nt) return;
  const StateMem st{ssb_stack + threadIdx.x, kFusedBlock};
  fused_column_lw<NREG, NS, URBAN, 1>(a, (int)q, true, st);
}
template <int NREG, int NS, bool URBAN>
__global__ void __launch_bounds__(kFusedBlock, SSB_REC_DOWN_BLOCKS) k_rec_down_lw(ClassArgs a, long nt) {
  const long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (q >= nt) return;
  fused_column_lw<NREG, NS, URBAN, 2>(a, (int)q, true, StateMem{nullptr, 0});
}
template <int NREG, int NS, bool URBAN>
static void launch_records_lw(const ClassArgs &a, long nt, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = sizeof(double) * LwFused<NREG, NS, URBAN>::step_doubles * kFusedBlock;
  if (!configured) {
    fast_note(cudaFuncSetAttribute(k_rec_up_lw<NREG, NS, URBAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const unsigned grid = (unsigned)((nt + kFusedBlock - 1) / kFusedBlock);
  k_rec_up_lw<NREG, NS, URBAN><<<grid, kFusedBlock, smem, st>>>(a, nt);
  fast_note(cudaGetLastError());
  k_rec_down_lw<NREG, NS, URBAN><<<grid, kFusedBlock, 0, st>>>(a, nt);
  fast_note(cudaGetLastError());
}
bool SSB_CAT(records_lw_ns, SSB_NS)(const ClassArgs &a, long nt, cudaStream_t st) {
  if (a.cfg.ns != SSB_NS || a.fused) return false;
  switch (a.cfg.nreg * 2 + (a.cfg.urban ? 1 : 0)) {
    case 2: launch_records_lw<1, SSB_NS, false>(a, nt, st); return true;
    case 3: launch_records_lw<1, SSB_NS, true>(a, nt, st); return true;
    case 4: launch_records_lw<2, SSB_NS, false>(a, nt, st); return true;
    case 5: launch_records_lw<2, SSB_NS, true>(a, nt, st); return true;
    case 6: launch_records_lw<3, SSB_NS, false>(a, nt, st); return true;
    case 7: launch_records_lw<3, SSB_NS, true>(a, nt, st); return true;
    default: return false;
  }
}
#endif

}  // namespace ssb
