// generic sw kernels, stream capacity 8
#define SSB_NS 8
#define SSB_KIND_SW
#include "ssb_kernels.cuh"
