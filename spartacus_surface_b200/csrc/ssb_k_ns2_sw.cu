// generic sw kernels, stream capacity 2
#define SSB_NS 2
#define SSB_KIND_SW
#include "ssb_kernels.cuh"
