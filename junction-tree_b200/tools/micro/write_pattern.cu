// HBM ceilings for the belief writers' access pattern (context for DESIGN.md's roofline notes):
// rows of B float64 (B = 65536: 512 KB per row); a CTA owns a 8 KB piece (512 x 16 B) of `rpc`
// consecutive rows -- what jt_project_tma_kernel / jt_beta_kernel write -- against a contiguous fill.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a write_pattern.cu -o write_pattern && ./write_pattern
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>   // 0: plain st, 1: st.cs (evict-first), 2: + one row read per 6 written
__global__ void __launch_bounds__(256) rows_kernel(double2* out, const double2* in, int rows, int rpc, long long Bv) {
    const long long col = (long long)blockIdx.y * 512 + threadIdx.x;
    const int r0 = blockIdx.x * rpc;
    const int r1 = min(rows, r0 + rpc);
    double2 m = make_double2(1.0, 2.0);
    for (int r = r0; r < r1; ++r) {
        if (MODE == 2 && (r % 6) == 0) {
            const double2 a = in[(long long)(r / 6) * Bv + col], b = in[(long long)(r / 6) * Bv + col + 256];
            m.x = a.x * b.x; m.y = a.y + b.y;
        }
        double2* p = out + (long long)r * Bv + col;
        if (MODE == 0) { p[0] = m; p[256] = m; }
        else {
            asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(m.x), "d"(m.y) : "memory");
            asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" :: "l"(p + 256), "d"(m.x), "d"(m.y) : "memory");
        }
    }
}

__global__ void fill_kernel(double2* out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_double2(1.0, 2.0);
}

template <typename F>
float timed(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) f();
    float best = 1e9f;
    for (int i = 0; i < 5; ++i) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    const long long B = 65536, Bv = B / 2;
    const int rows = 14000;                       // 7.3 GB written, like the largest belief launch of config 2
    double2 *out, *in;
    CK(cudaMalloc(&out, (size_t)rows * B * 8));
    CK(cudaMalloc(&in, (size_t)(rows / 6 + 1) * B * 8));
    CK(cudaMemset(in, 0, (size_t)(rows / 6 + 1) * B * 8));
    const double gb = (double)rows * B * 8 / 1e9;
    printf("contiguous fill            : %.0f GB/s\n", gb / timed([&] { fill_kernel<<<148 * 8, 1024>>>(out, (long long)rows * Bv); }) * 1e3);
    for (int rpc : {16, 64, 256, 1024}) {
        dim3 grid((rows + rpc - 1) / rpc, 64);
        printf("rows x 8 KB pieces, %4d rows per CTA: plain %.0f GB/s", rpc,
               gb / timed([&] { rows_kernel<0><<<grid, 256>>>(out, in, rows, rpc, Bv); }) * 1e3);
        printf(", st.cs %.0f GB/s", gb / timed([&] { rows_kernel<1><<<grid, 256>>>(out, in, rows, rpc, Bv); }) * 1e3);
        printf(", st.cs + 1/6 reads %.0f GB/s (r+w bytes)\n",
               gb * (1.0 + 1.0 / 6) / timed([&] { rows_kernel<2><<<grid, 256>>>(out, in, rows, rpc, Bv); }) * 1e3);
    }
    return 0;
}
