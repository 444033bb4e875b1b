// register-resident sw layer kernels, 4 stream(s) per hemisphere
#define SSB_NS 4
#define SSB_KIND_SW
#include "ssb_fast_kernels.cuh"
