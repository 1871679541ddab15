// Kernels and launchers of libjt_b200 for the log_sum_exp semiring (see jt_kernels.cuh, jt_launch.cuh).
#include "jt_launch.cuh"

JT_DEFINE_SEMIRING(SrLogSumExp, 2, jt_sr_log_sum_exp)
