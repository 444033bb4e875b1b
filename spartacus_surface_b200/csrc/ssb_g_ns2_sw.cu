// column-resident sw kernels, 2 stream(s) per hemisphere
#define SSB_NS 2
#define SSB_KIND_SW
#define SSB_FUSED_TU
#include "ssb_fused_kernels.cuh"
