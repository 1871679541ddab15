// Kernels and launchers of libjt_b200 for the max_product semiring (see jt_kernels.cuh, jt_launch.cuh).
#include "jt_launch.cuh"

JT_DEFINE_SEMIRING(SrMaxProduct, 1, jt_sr_max_product)
