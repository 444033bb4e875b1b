// register-resident lw layer kernels, 2 stream(s) per hemisphere
#define SSB_NS 2
#define SSB_KIND_LW
#include "ssb_fast_kernels.cuh"
