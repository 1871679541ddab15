// Kernels and launchers of libjt_b200 for the sum_product semiring (see jt_kernels.cuh, jt_launch.cuh).
#include "jt_launch.cuh"

JT_DEFINE_SEMIRING(SrSumProduct, 0, jt_sr_sum_product)
