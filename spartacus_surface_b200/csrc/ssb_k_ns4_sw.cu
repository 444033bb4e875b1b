// generic sw kernels, stream capacity 4
#define SSB_NS 4
#define SSB_KIND_SW
#include "ssb_kernels.cuh"
