// column-resident lw kernels, 1 stream(s) per hemisphere
#define SSB_NS 1
#define SSB_KIND_LW
#define SSB_FUSED_TU
#include "ssb_fused_kernels.cuh"
