// generic lw kernels, stream capacity 1
#define SSB_NS 1
#define SSB_KIND_LW
#include "ssb_kernels.cuh"
