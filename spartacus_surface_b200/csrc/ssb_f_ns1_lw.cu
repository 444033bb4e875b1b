// register-resident lw layer kernels, 1 stream(s) per hemisphere
#define SSB_NS 1
#define SSB_KIND_LW
#include "ssb_fast_kernels.cuh"
