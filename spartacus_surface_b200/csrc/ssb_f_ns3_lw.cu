// register-resident lw layer kernels, 3 stream(s) per hemisphere
#define SSB_NS 3
#define SSB_KIND_LW
#include "ssb_fast_kernels.cuh"
