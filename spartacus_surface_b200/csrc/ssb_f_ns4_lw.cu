// register-resident lw layer kernels, 4 stream(s) per hemisphere
#define SSB_NS 4
#define SSB_KIND_LW
#include "ssb_fast_kernels.cuh"
