// generic lw kernels, stream capacity 2
#define SSB_NS 2
#define SSB_KIND_LW
#include "ssb_kernels.cuh"
