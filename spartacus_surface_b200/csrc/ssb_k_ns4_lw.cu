// generic lw kernels, stream capacity 4
#define SSB_NS 4
#define SSB_KIND_LW
#include "ssb_kernels.cuh"
