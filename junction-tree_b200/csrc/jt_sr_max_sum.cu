// Kernels and launchers of libjt_b200 for the max_sum semiring (see jt_kernels.cuh, jt_launch.cuh).
#include "jt_launch.cuh"

JT_DEFINE_SEMIRING(SrMaxSum, 3, jt_sr_max_sum)
