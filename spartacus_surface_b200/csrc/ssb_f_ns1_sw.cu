// register-resident sw layer kernels, 1 stream(s) per hemisphere
#define SSB_NS 1
#define SSB_KIND_SW
#include "ssb_fast_kernels.cuh"
