// ssb_stage.cuh - level-major staging of the per-layer arrays of one chunk of columns.
//
// The reference layout packs the layers of a column contiguously ((nspec, ntotlay) arrays, layer
// il = istartlay(col) - 1 + lev).  The register-resident kernels run one thread per column (or per
// column and level), so in that layout the threads of a warp touch addresses nlay * nspec doubles
// apart: every 8-byte access moves its own 32-byte sector, and with ~40 k threads in flight the
// partially used sectors do not survive in L2 until their neighbours are touched (ncu, round 2:
// 3.1 KB of DRAM traffic per (column, layer) in a kernel whose useful bytes are 0.9 KB).  The C-ABI
// therefore does what the north-star describes for the shim: it repacks the per-layer inputs of a
// chunk into structure-of-arrays buffers ordered [level][column][interval] ("gather"), lets the
// kernels read and write those with consecutive addresses per warp, and copies the per-layer outputs
// back into the caller's arrays ("scatter").  Both directions go through a shared-memory tile so
// that reads and writes are full sectors on both sides.
#pragma once
#include "ssb_solver.cuh"

namespace ssb {

constexpr int kStageMaxArrays = 40;
struct StageList {
  int n;
  double *ref[kStageMaxArrays];     // the caller's array, reference layout (inputs are not written)
  double *staged[kStageMaxArrays];  // [lev][ic][g] buffer of the chunk
  int nspec[kStageMaxArrays];
};
struct StageArgs {
  StageList list;
  const int *cols, *nlay, *istartlay;
  int ncols, lmax;
};

// serial form (host check)
inline void stage_host(const StageArgs &s, bool scatter) {
  for (int k = 0; k < s.list.n; ++k) {
    const int ns = s.list.nspec[k];
    for (int ic = 0; ic < s.ncols; ++ic) {
      const int col = s.cols[ic], il1 = s.istartlay[col] - 1;
      for (int lev = 0; lev < s.nlay[col]; ++lev)
        for (int g = 0; g < ns; ++g) {
          double &r = s.list.ref[k][(size_t)g + (size_t)ns * (size_t)(il1 + lev)];
          double &t = s.list.staged[k][(size_t)g + (size_t)ns * ((size_t)ic + (size_t)lev * (size_t)s.ncols)];
          if (scatter)
            r = t;
          else
            t = r;
        }
    }
  }
}

#if defined(__CUDACC__)
constexpr int kStageBlock = 256;
constexpr int kStageTile = 4096;  // doubles of shared memory per block
// blockIdx.x: tile of `tc` chunk columns, blockIdx.y: array.  `lb` levels per pass of the tile.
template <bool SCATTER>
__global__ void __launch_bounds__(kStageBlock) k_stage(StageArgs s, int tc, int lb) {
  __shared__ double tile[kStageTile + 2 * 64];
  const int k = blockIdx.y, ns = s.list.nspec[k];
  double *ref = s.list.ref[k], *stg = s.list.staged[k];
  const int ic0 = blockIdx.x * tc;
  const int ncl = min(tc, s.ncols - ic0);
  if (ncl <= 0) return;
  const int row = ncl * ns, pitch = row | 1;  // odd pitch: the two access patterns stay conflict-poor
  for (int l0 = 0; l0 < s.lmax; l0 += lb) {
    const int nl = min(lb, s.lmax - l0);
    const int total = row * nl;
    if (!SCATTER) {
      for (int idx = threadIdx.x; idx < total; idx += kStageBlock) {  // (l, g) fastest: contiguous in `ref`
        const int c = idx / (nl * ns), r = idx % (nl * ns), l = r / ns, g = r % ns;
        const int col = s.cols[ic0 + c];
        if (l0 + l < s.nlay[col])
          tile[l * pitch + c * ns + g] = ref[(size_t)g + (size_t)ns * (size_t)(s.istartlay[col] - 1 + l0 + l)];
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < total; idx += kStageBlock) {  // (c, g) fastest: contiguous in `stg`
        const int l = idx / row, r = idx % row, c = r / ns;
        if (l0 + l < s.nlay[s.cols[ic0 + c]])
          stg[(size_t)(r % ns) + (size_t)ns * ((size_t)(ic0 + c) + (size_t)(l0 + l) * (size_t)s.ncols)] = tile[l * pitch + r];
      }
    } else {
      for (int idx = threadIdx.x; idx < total; idx += kStageBlock) {
        const int l = idx / row, r = idx % row, c = r / ns;
        if (l0 + l < s.nlay[s.cols[ic0 + c]])
          tile[l * pitch + r] = stg[(size_t)(r % ns) + (size_t)ns * ((size_t)(ic0 + c) + (size_t)(l0 + l) * (size_t)s.ncols)];
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < total; idx += kStageBlock) {
        const int c = idx / (nl * ns), r = idx % (nl * ns), l = r / ns, g = r % ns;
        const int col = s.cols[ic0 + c];
        if (l0 + l < s.nlay[col])
          ref[(size_t)g + (size_t)ns * (size_t)(s.istartlay[col] - 1 + l0 + l)] = tile[l * pitch + c * ns + g];
      }
    }
    __syncthreads();
  }
}

// Arrays with one value per layer (every array when nsw = nlw = 1): no tile needed.  A thread moves four
// vertically adjacent layers of one column of every array: on the reference side that is one full 32-byte
// sector per array, on the staged side four rows in which neighbouring threads are neighbouring columns.
template <bool SCATTER>
__global__ void __launch_bounds__(kStageBlock) k_stage1(StageArgs s) {
  // neighbouring threads take neighbouring groups of four layers of ONE column: on the reference side a
  // warp then covers whole 128-byte lines (8 columns x 16 layers when every column has 16 layers)
  const long t = blockIdx.x * (long)kStageBlock + threadIdx.x;
  const int ng = (s.lmax + 3) / 4;
  const int ic = (int)(t / ng), l0 = 4 * (int)(t % ng);
  if (ic >= s.ncols) return;
  const int col = s.cols[ic];
  const int cnt = s.nlay[col] - l0;
  if (cnt <= 0) return;
  const size_t rbase = (size_t)(s.istartlay[col] - 1 + l0);
  const size_t sbase = (size_t)ic + (size_t)l0 * (size_t)s.ncols;
  const size_t pitch = (size_t)s.ncols;
  const bool vec = cnt >= 4 && (rbase & 1) == 0;  // two aligned 16-byte accesses on the reference side
#pragma unroll 4
  for (int k = 0; k < s.list.n; ++k) {
    double *ref = s.list.ref[k] + rbase, *stg = s.list.staged[k] + sbase;
    double v[4];
    if (!SCATTER) {
      if (vec) {
        const double2 a = __ldcs((const double2 *)ref), b = __ldcs((const double2 *)ref + 1);
        v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < cnt) v[j] = __ldcs(ref + j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < cnt) stg[j * pitch] = v[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < cnt) v[j] = __ldcs(stg + j * pitch);
      if (vec) {
        __stcs((double2 *)ref, make_double2(v[0], v[1]));
        __stcs((double2 *)ref + 1, make_double2(v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < cnt) __stcs(ref + j, v[j]);
      }
    }
  }
}

// launches one grid per group of arrays with the same spectral width (the tile shape depends on it)
inline void stage_launch(const StageArgs &s, bool scatter, cudaStream_t st, long *launches) {
  if (s.list.n == 0 || s.ncols <= 0 || s.lmax <= 0) return;
  int done[kStageMaxArrays] = {0};
  for (int k0 = 0; k0 < s.list.n; ++k0) {
    if (done[k0]) continue;
    StageArgs g = s;
    g.list.n = 0;
    const int ns = s.list.nspec[k0];
    for (int k = k0; k < s.list.n; ++k)
      if (!done[k] && s.list.nspec[k] == ns) {
        g.list.ref[g.list.n] = s.list.ref[k];
        g.list.staged[g.list.n] = s.list.staged[k];
        g.list.nspec[g.list.n] = ns;
        ++g.list.n;
        done[k] = 1;
      }
    if (ns == 1) {
      const long nthreads = (long)s.ncols * ((s.lmax + 3) / 4);
      const unsigned blocks = (unsigned)((nthreads + kStageBlock - 1) / kStageBlock);
      if (scatter)
        k_stage1<true><<<blocks, kStageBlock, 0, st>>>(g);
      else
        k_stage1<false><<<blocks, kStageBlock, 0, st>>>(g);
      if (launches) ++*launches;
      continue;
    }
    int lb = s.lmax < 32 ? s.lmax : 32;
    while (lb > 1 && lb * ns > kStageTile) lb /= 2;  // (stage_supported caps the spectral width at kStageTile)
    int tc = kStageTile / (lb * ns);
    tc = tc > 64 ? 64 : (tc < 1 ? 1 : tc);
    const dim3 grid((unsigned)((s.ncols + tc - 1) / tc), (unsigned)g.list.n);
    if (scatter)
      k_stage<true><<<grid, kStageBlock, 0, st>>>(g, tc, lb);
    else
      k_stage<false><<<grid, kStageBlock, 0, st>>>(g, tc, lb);
    if (launches) ++*launches;
  }
}
#endif

}  // namespace ssb
