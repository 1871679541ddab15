// FP64 tensor-pipe throughput on this GPU: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) issued back to back
// by W warps per SM with C independent accumulator chains each.  Prints TFLOP/s per configuration.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_bench dmma_bench.cu && ./dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void k(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c[C][2];
#pragma unroll
    for (int i = 0; i < C; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int C>
void run(int warps_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * warps_per_sm * 32);
    const int iters = 20000;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    k<C><<<sms, warps_per_sm * 32>>>(out, 100);
    cudaEventRecord(t0);
    k<C><<<sms, warps_per_sm * 32>>>(out, iters);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    const double flops = 2.0 * 8 * 8 * 4 * (double)C * iters * warps_per_sm * sms;
    printf("warps/SM %2d chains %d: %.3f ms  %.2f TFLOP/s  (%.1f clk per DMMA per SM at 1.9 GHz)\n", warps_per_sm, C, ms,
           flops / ms / 1e9, ms * 1e-3 * 1.9e9 / ((double)C * iters * warps_per_sm));
    cudaFree(out);
}

int main() {
    run<1>(4);
    run<4>(4);
    run<8>(4);
    run<8>(8);
    run<8>(16);
    run<16>(16);
    return 0;
}
