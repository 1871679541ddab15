// Read bandwidth of 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) by copy size: one producer
// warp per CTA streams rows of a 512 KB-pitch matrix into a 4-stage shared-memory ring of 16 KB
// stages (16 copies of 1 KB ... 2 copies of 8 KB per stage), one consumer warp releases the stages.
// Context for the tile width of jt_dense_kernel (1 KB copies) against jt_project_tma_kernel (8 KB).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a bulk_copy_size.cu -o bulk_copy_size
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kStages = 4, kStageBytes = 16384;

// grid (x, tiles): tile = a column range of `piece` bytes of every row; x splits the rows
__global__ void __launch_bounds__(64) ring_kernel(const char* in, long long pitch, int rows, int rpc, int piece, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kStages * kStageBytes);
    const uint32_t full = smem_u32(bars), empty = smem_u32(bars + kStages), ring = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int per = kStageBytes / piece;                       // copies per stage
    const int r0 = blockIdx.x * rpc, r1 = min(rows, r0 + rpc);
    const char* base = in + (long long)blockIdx.y * piece;
    int stage = 0; uint32_t phase = 0;
    if (warp == 0) {
        for (int r = r0; r < r1; r += per) {
            const int n = min(per, r1 - r);
            mbar_wait(empty + 8 * stage, phase ^ 1);
            if (lane == 0) mbar_expect_tx(full + 8 * stage, (uint32_t)n * piece);
            __syncwarp();
            if (lane < n) bulk_g2s(ring + stage * kStageBytes + lane * piece, base + (long long)(r + lane) * pitch, piece, full + 8 * stage);
            __syncwarp();
            if (lane == 0) mbar_arrive(full + 8 * stage);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else {
        double acc = 0.0;
        for (int r = r0; r < r1; r += per) {
            mbar_wait(full + 8 * stage, phase);
            acc += reinterpret_cast<const double*>(smem + stage * kStageBytes)[lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + 8 * stage);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (acc == 12345.678) sink[0] = acc;
    }
}

int main() {
    const long long pitch = 524288;               // 65536 float64 per row
    const int rows = 8192;                        // 4.3 GB
    char* in; double* sink;
    CK(cudaMalloc(&in, (size_t)rows * pitch));
    CK(cudaMemset(in, 0, (size_t)rows * pitch));
    CK(cudaMalloc(&sink, 8));
    const int smem = kStages * kStageBytes + 64;
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int piece : {512, 1024, 2048, 4096, 8192}) {
        for (int rpc : {64, 256, 1024}) {
            dim3 grid((rows + rpc - 1) / rpc, (unsigned)(pitch / piece));
            float best = 1e9f;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0);
                ring_kernel<<<grid, 64, smem>>>(in, pitch, rows, rpc, piece, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it) best = ms < best ? ms : best;
            }
            CK(cudaGetLastError());
            setvbuf(stdout, nullptr, _IONBF, 0);
            printf("copy %5d B, %4d rows per CTA (%6u x %4u CTAs): %.0f GB/s\n", piece, rpc, grid.x, grid.y,
                   (double)rows * pitch / 1e9 / best * 1e3);
        }
    }
    return 0;
}
