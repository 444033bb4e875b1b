// generic sw kernels, stream capacity 1
#define SSB_NS 1
#define SSB_KIND_SW
#include "ssb_kernels.cuh"
