// generic lw kernels, stream capacity 8
#define SSB_NS 8
#define SSB_KIND_LW
#include "ssb_kernels.cuh"
