// libjt_b200: sm_100a kernels + C ABI for batched junction-tree sum-product propagation.
// ABI and task semantics: include/jt_b200.h.  Schedule producer: junctiontree/schedule.py.
//
// Everything on this path is HBM-bound (2 flops per 8..16 bytes), so the kernels are built
// around coalesced 16-byte accesses on the batch-innermost layout [entry][B]: a thread owns
// VEC consecutive instances of one output index s and walks the remaining clique axes r;
// all index arithmetic is table-driven, warp-uniform whenever the batch tile is >= 32 wide.

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/jt_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// error handling

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define JT_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(JT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));      \
    } while (0)

// ------------------------------------------------------------------------------------------
// device-side descriptors

struct DTask {
    long long src, out, beta, bel, own;  // entry offsets, -1 = absent
    int n_s, n_r, n_slo, n_rlo;
    int src_shi, src_slo, src_rhi, src_rlo;
    int rmsg_begin, rmsg_end, smsg_begin, smsg_end;
    int kind, out_space;
};

struct DMsg {
    long long off;    // entry offset (multiplied by B on the device)
    long long eoff;   // element offset added as is (jt_contract operands; 0 inside a plan)
    int a_hi, a_lo, b_hi, b_lo;
    int fid, pad;
};

struct KArgs {
    const DTask* tasks;   // first task of this launch
    const DMsg* msgs;     // all messages of the plan
    const int* tab;       // all index tables
    const int* prefix;    // [n_tasks + 1] first block of each task for this launch / tile shape
    void* work;
    void* fout;
    const void* fin;
    const int* fbase;     // [F][B] per-instance factor base offsets, or null
    long long B;          // instances (row pitch in elements)
    long long Bv;         // B / VEC
    int n_tasks;
    int bx_log2;          // batch-tile width in vectors (log2)
    int sy_log2;          // rows of s per block (log2)
    int flags;
    int fin_batched;
};

constexpr int kThreads = 256;
constexpr int kMaxSyLog2 = 12;  // largest chunk of s per block: 4096
constexpr int kRegMsgs = 4;   // r-dependent messages kept in registers
constexpr int kUnroll = 4;    // independent row loads in flight per thread

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack {
    T v[VEC];
};

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> pack_fill(T x) {
    Pack<T, VEC> p;
#pragma unroll
    for (int i = 0; i < VEC; ++i) p.v[i] = x;
    return p;
}

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld(const T* p) {
    return *reinterpret_cast<const Pack<T, VEC>*>(p);
}

template <typename T, int VEC>
__device__ __forceinline__ void st(T* p, const Pack<T, VEC>& x) {
    *reinterpret_cast<Pack<T, VEC>*>(p) = x;
}

template <typename T, int VEC>
__device__ __forceinline__ void mul(Pack<T, VEC>& a, const Pack<T, VEC>& b) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) a.v[i] *= b.v[i];
}

template <typename T, int VEC>
__device__ __forceinline__ void add(Pack<T, VEC>& a, const Pack<T, VEC>& b) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) a.v[i] += b.v[i];
}

// Locate the task of this block; s0 = first output index of the block's chunk of 2^sy_log2.
__device__ __forceinline__ const DTask* locate_chunk(const KArgs& a, int& s0) {
    const int bid = blockIdx.x;
    int lo = 0, hi = a.n_tasks;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.prefix + mid) <= bid) lo = mid; else hi = mid;
    }
    s0 = (bid - __ldg(a.prefix + lo)) << a.sy_log2;
    return a.tasks + lo;
}

// Locate the task of this block and the output index s of this thread (2^sy_log2 rows of s per
// block, 2^bx_log2 batch vectors per row).
__device__ __forceinline__ const DTask* locate(const KArgs& a, int& s, long long& bv) {
    const int tx = threadIdx.x & ((1 << a.bx_log2) - 1);
    const int ty = threadIdx.x >> a.bx_log2;
    bv = ((long long)blockIdx.y << a.bx_log2) + tx;
    int s0;
    const DTask* tk = locate_chunk(a, s0);
    s = s0 + ty;
    return tk;
}

// ------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) primitives, inline PTX for sm_100a

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "JT_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra JT_DONE_%=;\n"
        "bra JT_WAIT_%=;\n"
        "JT_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------
// evidence slicing (V1): per-instance base offset of every factor table
//   fbase[f][b] = sum over observed axes k of factor f:  state[b][var_k] * stride_k
// Pure integer arithmetic; out-of-range states are clamped and counted.

__global__ void __launch_bounds__(kThreads)
jt_evidence_kernel(const int* __restrict__ evidence, int n_evid, const int* __restrict__ ev_card,
                   const int* __restrict__ evf_ptr, const int* __restrict__ evf_var,
                   const int* __restrict__ evf_stride, int n_factors, long long B,
                   int* __restrict__ fbase, unsigned long long* __restrict__ errors) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int* row = evidence + b * n_evid;
    unsigned bad = 0;
    for (int f = 0; f < n_factors; ++f) {
        int acc = 0;
        for (int k = evf_ptr[f]; k < evf_ptr[f + 1]; ++k) {
            const int var = evf_var[k];
            int state = row[var];
            const int card = ev_card[var];
            if (state < 0 || state >= card) {
                ++bad;
                state = state < 0 ? 0 : card - 1;
            }
            acc += state * evf_stride[k];
        }
        fbase[(long long)f * B + b] = acc;
    }
    if (bad) atomicAdd(errors, (unsigned long long)bad);
}

// ------------------------------------------------------------------------------------------
// clique initialisation (E0 + V1): psi_C[s][b] = prod_f phi_f[ A_f(s) + fbase[f][b] ]

// A thread owns VEC batch columns and walks its rows of the block's chunk of s, so the
// per-instance factor offsets (which do not depend on s) are read once and kept in registers.
constexpr int kInitRegFactors = 6;

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) jt_init_kernel(const KArgs a) {
    typedef Pack<T, VEC> P;
    int s;
    long long bv;
    const DTask* tk = locate(a, s, bv);
    const int n_s = tk->n_s;
    if (s >= n_s || bv >= a.Bv) return;
    const int rows = kThreads >> a.bx_log2;                       // rows of s handled per step
    const int s_end = min(n_s, (s - (int)(threadIdx.x >> a.bx_log2)) + (1 << a.sy_log2));
    const int* __restrict__ tab = a.tab;
    const DMsg* __restrict__ msgs = a.msgs;
    const long long B = a.B, col = bv * VEC;
    const int n_slo = tk->n_slo;
    int s_hi = 0, s_lo = s;
    if (n_slo < n_s) {
        s_hi = s / n_slo;
        s_lo = s - s_hi * n_slo;
    }
    const T* __restrict__ fin = static_cast<const T*>(a.fin);
    const int f0 = tk->smsg_begin, nf = tk->smsg_end - f0;
    const bool gather = !a.fin_batched && a.fbase != nullptr;

    // per-factor, per-instance base offsets (evidence slicing), resident in registers
    int fb[kInitRegFactors][VEC];
#pragma unroll
    for (int j = 0; j < kInitRegFactors; ++j) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) fb[j][u] = 0;
        if (gather && j < nf) {
            const int* p = a.fbase + (long long)msgs[f0 + j].fid * B + col;
#pragma unroll
            for (int u = 0; u < VEC; ++u) fb[j][u] = p[u];
        }
    }

    T* out = static_cast<T*>(a.work) + tk->out * B + col;
    for (; s < s_end; s += rows) {
        P val = pack_fill<T, VEC>(T(1));
#pragma unroll
        for (int j = 0; j < kInitRegFactors; ++j) {
            if (j < nf) {
                const DMsg* m = msgs + f0 + j;
                const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
                if (a.fin_batched) {
                    mul(val, ld<T, VEC>(fin + idx * B + col));
                } else {
#pragma unroll
                    for (int u = 0; u < VEC; ++u) val.v[u] *= __ldg(fin + idx + fb[j][u]);
                }
            }
        }
        for (int j = kInitRegFactors; j < nf; ++j) {   // rare: many factors in one clique
            const DMsg* m = msgs + f0 + j;
            const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
            if (a.fin_batched) {
                mul(val, ld<T, VEC>(fin + idx * B + col));
            } else {
                const int* p = a.fbase ? a.fbase + (long long)m->fid * B + col : nullptr;
#pragma unroll
                for (int u = 0; u < VEC; ++u) val.v[u] *= __ldg(fin + idx + (p ? p[u] : 0));
            }
        }
        st<T, VEC>(out + (long long)s * B, val);
        s_lo += rows;
        while (s_lo >= n_slo) {
            s_lo -= n_slo;
            ++s_hi;
        }
    }
}

// ------------------------------------------------------------------------------------------
// projection task (collect E1+E2, distribute E3+E4+M1+E5, marginal E6, contract)

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) jt_project_kernel(const KArgs a) {
    typedef Pack<T, VEC> P;
    int s;
    long long bv;
    const DTask* tk = locate(a, s, bv);
    const int n_s = tk->n_s;
    if (s >= n_s || bv >= a.Bv) return;

    const int* __restrict__ tab = a.tab;
    const DMsg* __restrict__ msgs = a.msgs;
    const long long B = a.B, col = bv * VEC;
    T* work = static_cast<T*>(a.work);

    const int n_slo = tk->n_slo;
    int s_hi = 0, s_lo = s;
    if (n_slo < n_s) {
        s_hi = s / n_slo;
        s_lo = s - s_hi * n_slo;
    }

    // messages that do not depend on r, and the task's own up-message
    P sm = pack_fill<T, VEC>(T(1));
    for (int j = tk->smsg_begin; j < tk->smsg_end; ++j) {
        const DMsg* m = msgs + j;
        const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
        mul(sm, ld<T, VEC>(work + m->eoff + idx * B + col));
    }
    const bool has_own = tk->own >= 0;
    P own = pack_fill<T, VEC>(T(1));
    if (has_own) own = ld<T, VEC>(work + (tk->own + s) * B + col);

    // r-dependent messages: the first kRegMsgs are tracked in registers
    const int rm0 = tk->rmsg_begin;
    const int nr = tk->rmsg_end - rm0;
    const T* mptr[kRegMsgs];
    int mbhi[kRegMsgs], mblo[kRegMsgs];
#pragma unroll
    for (int j = 0; j < kRegMsgs; ++j) {
        mptr[j] = work;
        mbhi[j] = mblo[j] = 0;
        if (j < nr) {
            const DMsg* m = msgs + rm0 + j;
            const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
            mptr[j] = work + m->eoff + idx * B + col;
            mbhi[j] = m->b_hi;
            mblo[j] = m->b_lo;
        }
    }

    const bool has_src = tk->src >= 0;
    const long long s_off = __ldg(tab + tk->src_shi + s_hi) + __ldg(tab + tk->src_slo + s_lo);
    const T* sptr = work + ((has_src ? tk->src : 0) + s_off) * B + col;
    const bool wbeta = tk->beta >= 0;
    T* bptr = work + ((wbeta ? tk->beta : 0) + s_off) * B + col;
    P scale = sm;
    mul(scale, own);

    const int n_rlo = tk->n_rlo;
    const int n_rhi = tk->n_r / n_rlo;
    const int* __restrict__ t_rhi = tab + tk->src_rhi;
    const int* __restrict__ t_rlo = tab + tk->src_rlo;

    P acc[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) acc[u] = pack_fill<T, VEC>(T(0));

    for (int rh = 0; rh < n_rhi; ++rh) {
        const long long e_hi = __ldg(t_rhi + rh);
        long long mh[kRegMsgs];
#pragma unroll
        for (int j = 0; j < kRegMsgs; ++j) mh[j] = (j < nr) ? (long long)__ldg(tab + mbhi[j] + rh) : 0;

        int rl = 0;
        for (; rl + kUnroll <= n_rlo; rl += kUnroll) {
            long long e[kUnroll];
            P v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) e[u] = (e_hi + __ldg(t_rlo + rl + u)) * B;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                v[u] = has_src ? ld<T, VEC>(sptr + e[u]) : pack_fill<T, VEC>(T(1));
#pragma unroll
            for (int j = 0; j < kRegMsgs; ++j) {
                if (j < nr) {
                    P w[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u)
                        w[u] = ld<T, VEC>(mptr[j] + (mh[j] + __ldg(tab + mblo[j] + rl + u)) * B);
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) mul(v[u], w[u]);
                }
            }
            for (int j = kRegMsgs; j < nr; ++j) {   // rare: more than kRegMsgs r-dependent messages
                const DMsg* m = msgs + rm0 + j;
                const long long base = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo) +
                                       __ldg(tab + m->b_hi + rh);
#pragma unroll
                for (int u = 0; u < kUnroll; ++u)
                    mul(v[u], ld<T, VEC>(work + m->eoff + (base + __ldg(tab + m->b_lo + rl + u)) * B + col));
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) add(acc[u], v[u]);
            if (wbeta) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    mul(v[u], scale);
                    st<T, VEC>(bptr + e[u], v[u]);
                }
            }
        }
        for (; rl < n_rlo; ++rl) {
            const long long e = (e_hi + __ldg(t_rlo + rl)) * B;
            P v = has_src ? ld<T, VEC>(sptr + e) : pack_fill<T, VEC>(T(1));
#pragma unroll
            for (int j = 0; j < kRegMsgs; ++j)
                if (j < nr) mul(v, ld<T, VEC>(mptr[j] + (mh[j] + __ldg(tab + mblo[j] + rl)) * B));
            for (int j = kRegMsgs; j < nr; ++j) {
                const DMsg* m = msgs + rm0 + j;
                const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo) +
                                      __ldg(tab + m->b_hi + rh) + __ldg(tab + m->b_lo + rl);
                mul(v, ld<T, VEC>(work + m->eoff + idx * B + col));
            }
            add(acc[0], v);
            if (wbeta) {
                mul(v, scale);
                st<T, VEC>(bptr + e, v);
            }
        }
    }

    if (tk->out >= 0) {
        // pairwise combination of the partial sums
        add(acc[0], acc[1]);
        add(acc[2], acc[3]);
        add(acc[0], acc[2]);
        P o = acc[0];
        mul(o, sm);
        T* obase = tk->out_space ? static_cast<T*>(a.fout) : work;
        st<T, VEC>(obase + (tk->out + s) * B + col, o);
        if (tk->bel >= 0 && (a.flags & JT_SEP_BELIEFS)) {
            mul(o, own);
            st<T, VEC>(work + (tk->bel + s) * B + col, o);
        }
    }
}

// ------------------------------------------------------------------------------------------
// TMA-pipelined projection (same task semantics as jt_project_kernel).
//
// The LDG kernel above keeps every in-flight row in registers, so its memory-level parallelism
// is capped by occupancy (ncu, round 1: 22 % warps active, DRAM 33-53 %).  Here one elected
// producer thread walks the (s, r) index space and issues 1-D bulk async copies (cp.async.bulk,
// SASS UBLKCP) of whole batch-tile rows -- the clique row and one row per message -- into a
// shared-memory ring guarded by full/empty mbarriers; the consumer warps (one thread per 16-byte
// batch vector) multiply/accumulate out of shared memory and store beliefs and messages with
// coalesced 16-byte stores.  Bytes in flight = ring size (~96 KB per CTA, two CTAs per SM),
// independent of register pressure.  A stage holds the rows of one (s, r) item; the rows that
// depend on s only (s-only messages, the own up-message) ride along with the r = 0 item.

constexpr int kTmaSlots = 24;     // ring size in rows (one row = 16 bytes x consumer threads)
constexpr int kTmaMaxRows = 8;    // rows per stage supported (src + messages + own)

template <typename T>
__global__ void __launch_bounds__(kThreads + 32, 2) jt_project_tma_kernel(const KArgs a) {
    constexpr int VEC = 16 / (int)sizeof(T);
    typedef Pack<T, VEC> P;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int ct = blockDim.x - 32;                      // consumer threads = batch vectors per tile
    const int row_pitch = ct * 16;                       // bytes per ring row
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + kTmaSlots * row_pitch);
    int* e_row = reinterpret_cast<int*>(bars + 2 * kTmaSlots);
    const uint32_t slots_u32 = smem_u32(smem_raw);
    const uint32_t full_u32 = smem_u32(bars), empty_u32 = smem_u32(bars + kTmaSlots);

    int s0;
    const DTask* tk = locate_chunk(a, s0);
    const int n_s = tk->n_s;
    const int s1 = min(n_s, s0 + (1 << a.sy_log2));
    const long long col0v = (long long)blockIdx.y * ct;
    const int ncols = (int)min((long long)ct, a.Bv - col0v);
    const uint32_t row_bytes = (uint32_t)ncols * 16u;

    const bool has_src = tk->src >= 0, has_own = tk->own >= 0;
    const int m0 = tk->rmsg_begin;                       // r-dependent messages, then s-only ones
    const int nr = tk->rmsg_end - m0;
    const int nsm = tk->smsg_end - tk->smsg_begin;
    const int rows_item = (has_src ? 1 : 0) + nr;
    const int rows_extra = nsm + (has_own ? 1 : 0);
    const int rows_stage = rows_item + rows_extra;
    const int n_stage = kTmaSlots / rows_stage;
    const int n_r = tk->n_r;
    const long long B = a.B;

    const int n_cwarps = ct >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stage; ++i) {
            mbar_init(full_u32 + 8 * i, 1);
            mbar_init(empty_u32 + 8 * i, n_cwarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == n_cwarps) {
        // ---------------- producer warp: lane k owns row k of every stage ----------------
        // Row order inside a stage: [src] [r-dependent messages] [s-only messages] [own].
        const int src_rows = has_src ? 1 : 0;
        const int jm = lane - src_rows;                       // message index of this lane's row
        const bool is_src = has_src && lane == 0;
        const bool is_rmsg = jm >= 0 && jm < nr;
        const bool is_smsg = jm >= nr && jm < nr + nsm;
        const bool is_own = has_own && lane == rows_stage - 1;
        if (lane >= rows_stage) return;
        const unsigned pmask = (1u << rows_stage) - 1u;
        const bool per_item = is_src || is_rmsg;              // fetched for every (s, r), else with r = 0 only

        const int* __restrict__ tab = a.tab;
        const T* work = static_cast<const T*>(a.work);
        const int n_slo = tk->n_slo, n_rlo = tk->n_rlo, n_rhi = n_r / n_rlo;
        // per-lane row description, loaded once
        long long base = 0, eoff = 0;
        const int* t_ahi = tab;
        const int* t_alo = tab;
        const int* t_bhi = tab;
        const int* t_blo = tab;
        if (is_src) {
            base = tk->src;
            t_ahi = tab + tk->src_shi; t_alo = tab + tk->src_slo;
            t_bhi = tab + tk->src_rhi; t_blo = tab + tk->src_rlo;
        } else if (is_rmsg || is_smsg) {
            const DMsg* m = a.msgs + m0 + jm;
            base = m->off;
            eoff = m->eoff;
            t_ahi = tab + m->a_hi; t_alo = tab + m->a_lo;
            t_bhi = tab + m->b_hi; t_blo = tab + m->b_lo;
        } else {
            base = tk->own;
        }
        const T* origin = work + eoff + col0v * VEC;
        const uint32_t dst0 = slots_u32 + (uint32_t)lane * (uint32_t)row_pitch;
        const uint32_t stage_bytes = (uint32_t)rows_stage * (uint32_t)row_pitch;

        int s_hi = 0, s_lo = s0;
        if (n_slo < n_s) {
            s_hi = s0 / n_slo;
            s_lo = s0 - s_hi * n_slo;
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int s = s0; s < s1; ++s) {
            const int s_idx = is_own ? s : __ldg(t_ahi + s_hi) + __ldg(t_alo + s_lo);
            const T* rowp = origin + (base + s_idx) * B;
            bool first = true;
            for (int rh = 0; rh < n_rhi; ++rh) {
                const int h = per_item ? __ldg(t_bhi + rh) : 0;
                for (int rl = 0; rl < n_rlo; ++rl) {
                    const uint32_t full = full_u32 + 8 * stage;
                    const int e = per_item ? h + __ldg(t_blo + rl) : 0;
                    mbar_wait(empty_u32 + 8 * stage, phase ^ 1);
                    if (per_item || first) {
                        if (is_src) e_row[stage] = s_idx + e;
                        mbar_expect_tx(full, row_bytes);
                        bulk_g2s(dst0 + (uint32_t)stage * stage_bytes, rowp + (long long)e * B, row_bytes, full);
                    }
                    first = false;
                    __syncwarp(pmask);                       // every lane's expect_tx precedes the arrive
                    if (lane == 0) mbar_arrive(full);
                    if (++stage == n_stage) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (++s_lo == n_slo) {
                s_lo = 0;
                ++s_hi;
            }
        }
        return;
    }

    // ---------------- consumers: thread t owns batch vector col0v + t ----------------
    const int t = threadIdx.x;
    const bool active = t < ncols;
    T* work = static_cast<T*>(a.work);
    const long long col = (col0v + t) * VEC;
    const bool wbeta = tk->beta >= 0, wout = tk->out >= 0;
    const bool wbel = wout && tk->bel >= 0 && (a.flags & JT_SEP_BELIEFS);
    T* bptr = work + (wbeta ? tk->beta : 0) * B + col;
    T* optr = (tk->out_space ? static_cast<T*>(a.fout) : work) + (wout ? tk->out : 0) * B + col;
    T* lptr = work + (wbel ? tk->bel : 0) * B + col;
    const unsigned char* my = smem_raw + t * 16;
    const int stage_pitch = rows_stage * row_pitch;
    const int src_rows = has_src ? 1 : 0;   // consumer-side copy

    int stage = 0;
    uint32_t phase = 0;
    for (int s = s0; s < s1; ++s) {
        P sm = pack_fill<T, VEC>(T(1)), own = sm, scale = sm;
        P acc0 = pack_fill<T, VEC>(T(0)), acc1 = acc0;
        for (int r = 0; r < n_r; ++r) {
            mbar_wait(full_u32 + 8 * stage, phase);
            const unsigned char* base = my + stage * stage_pitch;
            P v = has_src ? *reinterpret_cast<const P*>(base) : pack_fill<T, VEC>(T(1));
            for (int j = 0; j < nr; ++j) mul(v, *reinterpret_cast<const P*>(base + (src_rows + j) * row_pitch));
            if (r == 0) {
                for (int j = 0; j < nsm; ++j)
                    mul(sm, *reinterpret_cast<const P*>(base + (rows_item + j) * row_pitch));
                if (has_own) own = *reinterpret_cast<const P*>(base + (rows_item + nsm) * row_pitch);
                scale = sm;
                mul(scale, own);
            }
            const int e = e_row[stage];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_u32 + 8 * stage);
            if (r & 1) add(acc1, v); else add(acc0, v);
            if (wbeta && active) {
                mul(v, scale);
                st<T, VEC>(bptr + (long long)e * B, v);
            }
            if (++stage == n_stage) {
                stage = 0;
                phase ^= 1;
            }
        }
        if (wout && active) {
            add(acc0, acc1);
            mul(acc0, sm);
            st<T, VEC>(optr + (long long)s * B, acc0);
            if (wbel) {
                mul(acc0, own);
                st<T, VEC>(lptr + (long long)s * B, acc0);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
jt_ratio_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T d = b[i];
        out[i] = d != T(0) ? a[i] / d : T(0);
    }
}

static_assert(kUnroll == 4, "the pairwise combination above assumes four partial sums");

}  // namespace

// ------------------------------------------------------------------------------------------
// host side

struct jt_plan {
    std::vector<int64_t> hdr, node_off, node_size, fin_off, fin_size, fout_off, fout_size;
    std::vector<int> ev_card, evf_ptr, evf_var, evf_stride;
    std::vector<DTask> tasks;
    std::vector<DMsg> msgs;
    std::vector<int> tab;
    struct Launch {
        int phase, begin, end, level;
        size_t prefix_off[kMaxSyLog2 + 1];
        long long blocks[kMaxSyLog2 + 1];
        bool tma_ok;          // every task fits the TMA kernel's stage (rows per stage <= kTmaMaxRows)
        int min_nr;           // smallest n_r of the launch
        long long total_s;    // sum of n_s
    };
    std::vector<Launch> launches;
    std::vector<int> prefix;

    int device = -1;
    DTask* d_tasks = nullptr;
    DMsg* d_msgs = nullptr;
    int* d_tab = nullptr;
    int* d_prefix = nullptr;
    int* d_ev = nullptr;   // ev_card | evf_ptr | evf_var | evf_stride
};

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t dtype_size(int dtype) { return dtype == JT_F64 ? 8 : 4; }

struct WorkspaceLayout {
    size_t work_bytes, fbase_off, err_off, total;
};

WorkspaceLayout workspace_layout(const jt_plan* p, int64_t B, int dtype) {
    WorkspaceLayout w;
    const int64_t entries = p->hdr[JT_H_CLIQUE_ENTRIES] + 3 * p->hdr[JT_H_SEP_ENTRIES];
    w.work_bytes = align_up((size_t)entries * (size_t)B * dtype_size(dtype), 256);
    w.fbase_off = w.work_bytes;
    const size_t fbase = p->hdr[JT_H_NEVID] > 0 ? (size_t)p->hdr[JT_H_NFACTORS] * (size_t)B * 4 : 0;
    w.err_off = w.fbase_off + align_up(fbase, 256);
    w.total = w.err_off + 256;
    return w;
}

int pick_vec(int64_t B, int dtype) {
    const int maxv = dtype == JT_F64 ? 2 : 4;
    for (int v = maxv; v > 1; v >>= 1)
        if (B % v == 0) return v;
    return 1;
}

// tile shape for a batch of Bv vectors: bx = min(256, pow2ceil(Bv)), sy = 256 / bx
void pick_tile(long long Bv, int& bx_log2, int& sy_log2) {
    bx_log2 = 0;
    while ((1LL << bx_log2) < Bv && bx_log2 < 8) ++bx_log2;
    sy_log2 = 8 - bx_log2;
}

int g_tma_enabled = -1;   // JT_DISABLE_TMA=1 forces the LDG kernel (debugging / A-B timing)

bool tma_enabled() {
    if (g_tma_enabled < 0) {
        const char* e = getenv("JT_DISABLE_TMA");
        g_tma_enabled = (e && e[0] == '1') ? 0 : 1;
    }
    return g_tma_enabled == 1;
}

template <typename T>
int launch_tma(const jt_plan* p, const jt_plan::Launch& L, KArgs a, cudaStream_t stream) {
    const int ct = a.Bv >= 256 ? 256 : (a.Bv >= 128 ? 128 : 64);
    const long long tiles = (a.Bv + ct - 1) / ct;
    // chunk of s per CTA: aim at ~8 CTAs per SM over the launch, but keep >= 32 (s, r) items per
    // CTA so the pipeline fill is amortised
    int sy_log2 = 0;
    const long long target = 148LL * 8;
    while (sy_log2 < kMaxSyLog2 && (L.total_s * tiles) >> (sy_log2 + 1) >= target) ++sy_log2;
    while (sy_log2 < kMaxSyLog2 && ((long long)L.min_nr << sy_log2) < 32) ++sy_log2;
    a.sy_log2 = sy_log2;
    a.bx_log2 = 0;
    a.tasks = p->d_tasks + L.begin;
    a.n_tasks = L.end - L.begin;
    a.prefix = p->d_prefix + L.prefix_off[sy_log2];
    const long long gx = L.blocks[sy_log2];
    if (gx <= 0) return JT_OK;
    if (gx > 2147483647LL || tiles > 65535)
        return fail(JT_ERR_INVALID, "launch grid %lld x %lld exceeds CUDA limits; split the batch", gx, tiles);
    const size_t smem = (size_t)kTmaSlots * ct * 16 + 2 * kTmaSlots * 8 + kTmaSlots * 4 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        JT_CUDA(cudaFuncSetAttribute(jt_project_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kTmaSlots * 256 * 16 + 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)gx, (unsigned)tiles, 1);
    jt_project_tma_kernel<T><<<grid, ct + 32, smem, stream>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    JT_CUDA(cudaGetLastError());
    return JT_OK;
}

template <typename T, int VEC>
int launch_tasks(const jt_plan* p, const jt_plan::Launch& L, KArgs a, cudaStream_t stream) {
    if (L.phase != JT_PHASE_INIT && VEC * sizeof(T) == 16 && L.tma_ok && a.Bv >= 64 && tma_enabled())
        return launch_tma<T>(p, L, a, stream);
    int bx_log2, sy_log2;
    pick_tile(a.Bv, bx_log2, sy_log2);
    if (L.phase == JT_PHASE_INIT) {
        // a thread walks ~32 rows of s so the per-instance factor offsets stay in registers
        sy_log2 = sy_log2 + 5 > kMaxSyLog2 ? kMaxSyLog2 : sy_log2 + 5;
        while (sy_log2 > 8 - bx_log2 && (L.total_s >> sy_log2) * ((a.Bv + (1LL << bx_log2) - 1) >> bx_log2) < 148 * 4)
            --sy_log2;
    }
    a.bx_log2 = bx_log2;
    a.sy_log2 = sy_log2;
    a.tasks = p->d_tasks + L.begin;
    a.n_tasks = L.end - L.begin;
    a.prefix = p->d_prefix + L.prefix_off[sy_log2];
    const long long gx = L.blocks[sy_log2];
    const long long gy = (a.Bv + (1LL << bx_log2) - 1) >> bx_log2;
    if (gx <= 0) return JT_OK;
    if (gx > 2147483647LL || gy > 65535)
        return fail(JT_ERR_INVALID, "launch grid %lld x %lld exceeds CUDA limits; split the batch", gx, gy);
    dim3 grid((unsigned)gx, (unsigned)gy, 1);
    if (L.phase == JT_PHASE_INIT)
        jt_init_kernel<T, VEC><<<grid, kThreads, 0, stream>>>(a);
    else
        jt_project_kernel<T, VEC><<<grid, kThreads, 0, stream>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    JT_CUDA(cudaGetLastError());
    return JT_OK;
}

int dispatch(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, int dtype, int vec,
             cudaStream_t stream) {
    if (dtype == JT_F64) {
        if (vec == 2) return launch_tasks<double, 2>(p, L, a, stream);
        return launch_tasks<double, 1>(p, L, a, stream);
    }
    if (vec == 4) return launch_tasks<float, 4>(p, L, a, stream);
    if (vec == 2) return launch_tasks<float, 2>(p, L, a, stream);
    return launch_tasks<float, 1>(p, L, a, stream);
}

int check_common(const jt_plan* p, int64_t B, int dtype, const void* workspace) {
    if (!p) return fail(JT_ERR_INVALID, "plan is null");
    if (B <= 0) return fail(JT_ERR_INVALID, "batch size must be positive");
    if (dtype != JT_F32 && dtype != JT_F64) return fail(JT_ERR_INVALID, "dtype must be JT_F32 or JT_F64");
    if (!workspace) return fail(JT_ERR_INVALID, "workspace is null");
    if (p->device < 0) return fail(JT_ERR_INVALID, "plan not uploaded: call jt_plan_upload first");
    return JT_OK;
}

KArgs base_args(const jt_plan* p, int64_t B, int dtype, void* workspace, int vec) {
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.msgs = p->d_msgs;
    a.tab = p->d_tab;
    a.work = workspace;
    a.B = B;
    a.Bv = B / vec;
    return a;
}

int run_phases(jt_plan* p, int64_t B, int dtype, void* workspace, int phase_lo, int phase_hi,
               KArgs a, int vec, cudaStream_t stream) {
    for (const auto& L : p->launches) {
        if (L.phase < phase_lo || L.phase > phase_hi) continue;
        int rc = dispatch(p, L, a, dtype, vec, stream);
        if (rc != JT_OK) return rc;
    }
    return JT_OK;
}

}  // namespace

extern "C" {

int jt_abi_version(void) { return JT_ABI_VERSION; }

const char* jt_last_error_string(void) { return g_err; }

int64_t jt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int jt_plan_create(const void* blob, size_t nbytes, jt_plan** out) {
    if (!blob || !out) return fail(JT_ERR_INVALID, "null argument");
    *out = nullptr;
    if (nbytes < JT_H_WORDS * 8 || nbytes % 8) return fail(JT_ERR_INVALID, "plan blob too short or misaligned");
    const int64_t* w = static_cast<const int64_t*>(blob);
    std::vector<int64_t> copy;
    if (reinterpret_cast<uintptr_t>(blob) % 8) {   // tolerate unaligned input
        copy.resize(nbytes / 8);
        memcpy(copy.data(), blob, nbytes);
        w = copy.data();
    }
    if (w[JT_H_MAGIC] != JT_MAGIC) return fail(JT_ERR_INVALID, "bad plan magic");
    if (w[JT_H_VERSION] != JT_ABI_VERSION)
        return fail(JT_ERR_INVALID, "plan version %lld, library expects %d", (long long)w[JT_H_VERSION], JT_ABI_VERSION);
    for (int i = 2; i < JT_H_WORDS; ++i)
        if (w[i] < 0) return fail(JT_ERR_INVALID, "negative header word %d", i);
    jt_plan* p = new (std::nothrow) jt_plan;
    if (!p) return fail(JT_ERR_NOMEM, "out of host memory");
    p->hdr.assign(w, w + JT_H_WORDS);
    const int64_t n_nodes = w[JT_H_NCLIQUES] + w[JT_H_NSEPS];
    const int64_t F = w[JT_H_NFACTORS], n_evid = w[JT_H_NEVID], n_evf = w[JT_H_NEVF];
    const int64_t n_tasks = w[JT_H_NTASKS], n_msgs = w[JT_H_NMSGS], n_launch = w[JT_H_NLAUNCHES];
    const int64_t n_tab = w[JT_H_NTAB];
    const int64_t evf_ptr_n = F > 0 ? F + 1 : 0;
    const int64_t words = JT_H_WORDS + 2 * n_nodes + 4 * F + n_evid + evf_ptr_n + 2 * n_evf +
                          n_tasks * JT_TASK_WORDS + n_msgs * JT_MSG_WORDS + n_launch * JT_LAUNCH_WORDS;
    const int64_t tab_words = (n_tab + 1) / 2;
    if ((int64_t)(nbytes / 8) != words + tab_words) {
        delete p;
        return fail(JT_ERR_INVALID, "plan blob size mismatch: %zu bytes, expected %lld", nbytes,
                    (long long)(words + tab_words) * 8);
    }
    const int64_t* q = w + JT_H_WORDS;
    auto take64 = [&](std::vector<int64_t>& v, int64_t n) { v.assign(q, q + n); q += n; };
    auto take32 = [&](std::vector<int>& v, int64_t n) {
        v.resize(n);
        for (int64_t i = 0; i < n; ++i) v[i] = (int)q[i];
        q += n;
    };
    take64(p->node_off, n_nodes);
    take64(p->node_size, n_nodes);
    take64(p->fin_off, F);
    take64(p->fin_size, F);
    take64(p->fout_off, F);
    take64(p->fout_size, F);
    take32(p->ev_card, n_evid);
    take32(p->evf_ptr, evf_ptr_n);
    take32(p->evf_var, n_evf);
    take32(p->evf_stride, n_evf);

    const int64_t work_entries = w[JT_H_CLIQUE_ENTRIES] + 3 * w[JT_H_SEP_ENTRIES];
    auto bad = [&](const char* what, int64_t i) {
        delete p;
        return fail(JT_ERR_INVALID, "malformed plan: %s (item %lld)", what, (long long)i);
    };
    for (int64_t k = 0; k < n_evf; ++k)
        if (p->evf_var[k] < 0 || p->evf_var[k] >= n_evid) return bad("evidence variable index", k);
    for (int64_t f = 0; f + 1 < evf_ptr_n; ++f)
        if (p->evf_ptr[f] > p->evf_ptr[f + 1] || p->evf_ptr[f + 1] > n_evf) return bad("evidence pointer", f);

    p->tasks.resize(n_tasks);
    for (int64_t i = 0; i < n_tasks; ++i, q += JT_TASK_WORDS) {
        DTask& t = p->tasks[i];
        t.src = q[JT_T_SRC]; t.out = q[JT_T_OUT]; t.beta = q[JT_T_BETA]; t.bel = q[JT_T_BEL]; t.own = q[JT_T_OWN];
        if (q[JT_T_NS] <= 0 || q[JT_T_NS] > 2147483647LL || q[JT_T_NR] <= 0 || q[JT_T_NR] > 2147483647LL)
            return bad("task index-space size", i);
        t.n_s = (int)q[JT_T_NS]; t.n_r = (int)q[JT_T_NR]; t.n_slo = (int)q[JT_T_NSLO]; t.n_rlo = (int)q[JT_T_NRLO];
        if (t.n_slo <= 0 || t.n_rlo <= 0 || t.n_s % t.n_slo || t.n_r % t.n_rlo) return bad("task table split", i);
        t.src_shi = (int)q[JT_T_SRC_SHI]; t.src_slo = (int)q[JT_T_SRC_SLO];
        t.src_rhi = (int)q[JT_T_SRC_RHI]; t.src_rlo = (int)q[JT_T_SRC_RLO];
        t.rmsg_begin = (int)q[JT_T_RMSG_BEGIN]; t.rmsg_end = (int)q[JT_T_RMSG_END];
        t.smsg_begin = (int)q[JT_T_SMSG_BEGIN]; t.smsg_end = (int)q[JT_T_SMSG_END];
        t.kind = (int)q[JT_T_KIND]; t.out_space = (int)q[JT_T_OUT_SPACE];
        if (t.rmsg_begin < 0 || t.rmsg_begin > t.rmsg_end || t.rmsg_end > n_msgs || t.smsg_begin < 0 ||
            t.smsg_begin > t.smsg_end || t.smsg_end > n_msgs)
            return bad("task message range", i);
        const int n_shi = t.n_s / t.n_slo, n_rhi = t.n_r / t.n_rlo;
        if (t.kind == JT_KIND_PROJECT) {
            if (t.src_shi < 0 || t.src_shi + n_shi > n_tab || t.src_slo < 0 || t.src_slo + t.n_slo > n_tab ||
                t.src_rhi < 0 || t.src_rhi + n_rhi > n_tab || t.src_rlo < 0 || t.src_rlo + t.n_rlo > n_tab)
                return bad("task table range", i);
        } else if (t.kind != JT_KIND_INIT) {
            return bad("task kind", i);
        }
        for (long long off : {t.src, t.beta, t.bel, t.own})
            if (off < -1 || off >= work_entries) return bad("task buffer offset", i);
        if (t.out < -1) return bad("task output offset", i);
        if (t.out_space == 0 && t.out >= work_entries) return bad("task output offset", i);
        if (t.out_space == 1 && t.out + t.n_s > w[JT_H_FOUT_ENTRIES]) return bad("factor output range", i);
    }
    p->msgs.resize(n_msgs);
    for (int64_t i = 0; i < n_msgs; ++i, q += JT_MSG_WORDS) {
        DMsg& m = p->msgs[i];
        m.off = q[JT_M_OFF];
        m.a_hi = (int)q[JT_M_AHI]; m.a_lo = (int)q[JT_M_ALO]; m.b_hi = (int)q[JT_M_BHI]; m.b_lo = (int)q[JT_M_BLO];
        m.fid = (int)q[JT_M_FID]; m.pad = 0; m.eoff = 0;
        if (m.off < 0 || m.a_hi < 0 || m.a_lo < 0 || m.b_hi < 0 || m.b_lo < 0 || m.a_hi >= n_tab + 1 ||
            m.a_lo >= n_tab + 1 || m.b_hi >= n_tab + 1 || m.b_lo >= n_tab + 1 || m.fid >= F)
            return bad("message descriptor", i);
    }
    p->launches.resize(n_launch);
    for (int64_t i = 0; i < n_launch; ++i, q += JT_LAUNCH_WORDS) {
        jt_plan::Launch& L = p->launches[i];
        L.phase = (int)q[JT_L_PHASE]; L.begin = (int)q[JT_L_BEGIN]; L.end = (int)q[JT_L_END]; L.level = (int)q[JT_L_LEVEL];
        if (L.phase < JT_PHASE_INIT || L.phase > JT_PHASE_MARGINAL || L.begin < 0 || L.begin >= L.end || L.end > n_tasks)
            return bad("launch descriptor", i);
        for (int t = L.begin; t < L.end; ++t)
            if ((p->tasks[t].kind == JT_KIND_INIT) != (L.phase == JT_PHASE_INIT)) return bad("task kind vs phase", i);
        L.tma_ok = true;
        L.min_nr = 2147483647;
        L.total_s = 0;
        for (int t = L.begin; t < L.end; ++t) {
            const DTask& k = p->tasks[t];
            const int rows = (k.src >= 0 ? 1 : 0) + (k.rmsg_end - k.rmsg_begin) + (k.smsg_end - k.smsg_begin) +
                             (k.own >= 0 ? 1 : 0);
            if (rows < 1 || rows > kTmaMaxRows || k.rmsg_end != k.smsg_begin) L.tma_ok = false;
            L.min_nr = k.n_r < L.min_nr ? k.n_r : L.min_nr;
            L.total_s += k.n_s;
        }
        // block prefix per tile shape: a block covers 2^sy consecutive values of s of one task
        for (int sy = 0; sy <= kMaxSyLog2; ++sy) {
            L.prefix_off[sy] = p->prefix.size();
            long long acc = 0;
            for (int t = L.begin; t < L.end; ++t) {
                p->prefix.push_back((int)acc);
                acc += ((long long)p->tasks[t].n_s + (1 << sy) - 1) >> sy;
                if (acc > 2147483647LL) return bad("launch too large", i);
            }
            p->prefix.push_back((int)acc);
            L.blocks[sy] = acc;
        }
    }
    const int32_t* tabp = reinterpret_cast<const int32_t*>(q);
    p->tab.assign(tabp, tabp + n_tab);
    *out = p;
    return JT_OK;
}

void jt_plan_destroy(jt_plan* p) {
    if (!p) return;
    if (p->device >= 0) {
        cudaFree(p->d_tasks);
        cudaFree(p->d_msgs);
        cudaFree(p->d_tab);
        cudaFree(p->d_prefix);
        cudaFree(p->d_ev);
    }
    delete p;
}

int jt_plan_query(const jt_plan* p, int what, int64_t* out) {
    if (!p || !out || what < 0 || what >= JT_H_WORDS) return fail(JT_ERR_INVALID, "bad query");
    *out = p->hdr[what];
    return JT_OK;
}

int jt_plan_node_range(const jt_plan* p, int node, int64_t* offset, int64_t* count) {
    if (!p || node < 0 || node >= (int)p->node_off.size()) return fail(JT_ERR_INVALID, "bad node index");
    if (offset) *offset = p->node_off[node];
    if (count) *count = p->node_size[node];
    return JT_OK;
}

int jt_plan_message_offsets(const jt_plan* p, int node, int64_t* up, int64_t* down) {
    if (!p) return fail(JT_ERR_INVALID, "plan is null");
    const int64_t n_c = p->hdr[JT_H_NCLIQUES];
    if (node < n_c || node >= (int64_t)p->node_off.size()) return fail(JT_ERR_INVALID, "not a separator node");
    const int64_t rel = p->node_off[node] - p->hdr[JT_H_CLIQUE_ENTRIES];
    const int64_t up_base = p->hdr[JT_H_CLIQUE_ENTRIES] + p->hdr[JT_H_SEP_ENTRIES];
    if (up) *up = up_base + rel;
    if (down) *down = up_base + p->hdr[JT_H_SEP_ENTRIES] + rel;
    return JT_OK;
}

int jt_workspace_bytes(const jt_plan* p, int64_t B, int dtype, size_t* out) {
    if (!p || !out || B <= 0 || (dtype != JT_F32 && dtype != JT_F64)) return fail(JT_ERR_INVALID, "bad argument");
    *out = workspace_layout(p, B, dtype).total;
    return JT_OK;
}

int jt_plan_upload(jt_plan* p) {
    if (!p) return fail(JT_ERR_INVALID, "plan is null");
    int dev = -1;
    JT_CUDA(cudaGetDevice(&dev));
    if (p->device == dev) return JT_OK;
    if (p->device >= 0) return fail(JT_ERR_INVALID, "plan already uploaded to device %d", p->device);
    auto up = [](auto** dst, const auto& v) -> cudaError_t {
        const size_t bytes = v.size() * sizeof(v[0]);
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dst), bytes ? bytes : 16);
        if (e != cudaSuccess || !bytes) return e;
        return cudaMemcpy(*dst, v.data(), bytes, cudaMemcpyHostToDevice);
    };
    JT_CUDA(up(&p->d_tasks, p->tasks));
    JT_CUDA(up(&p->d_msgs, p->msgs));
    JT_CUDA(up(&p->d_tab, p->tab));
    JT_CUDA(up(&p->d_prefix, p->prefix));
    std::vector<int> ev;
    ev.insert(ev.end(), p->ev_card.begin(), p->ev_card.end());
    ev.insert(ev.end(), p->evf_ptr.begin(), p->evf_ptr.end());
    ev.insert(ev.end(), p->evf_var.begin(), p->evf_var.end());
    ev.insert(ev.end(), p->evf_stride.begin(), p->evf_stride.end());
    JT_CUDA(up(&p->d_ev, ev));
    p->device = dev;
    return JT_OK;
}

int jt_init(jt_plan* p, const void* factor_tables, int factors_batched, const int32_t* evidence, int64_t B,
            int dtype, void* workspace, void* stream_) {
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    if (p->hdr[JT_H_NFACTORS] == 0) return fail(JT_ERR_INVALID, "plan has no factors: nothing to initialise");
    if (!factor_tables) return fail(JT_ERR_INVALID, "factor_tables is null");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const WorkspaceLayout wl = workspace_layout(p, B, dtype);
    const int n_evid = (int)p->hdr[JT_H_NEVID];
    const int vec = pick_vec(B, dtype);
    KArgs a = base_args(p, B, dtype, workspace, vec);
    a.fin = factor_tables;
    a.fin_batched = factors_batched ? 1 : 0;
    if (n_evid > 0) {
        if (!evidence) return fail(JT_ERR_INVALID, "plan has %d evidence variables but evidence is null", n_evid);
        if (factors_batched) return fail(JT_ERR_INVALID, "per-instance factor tables cannot be combined with evidence indices");
        int* fbase = reinterpret_cast<int*>(static_cast<char*>(workspace) + wl.fbase_off);
        unsigned long long* err = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + wl.err_off);
        const int F = (int)p->hdr[JT_H_NFACTORS];
        const int* d_card = p->d_ev;
        const int* d_ptr = d_card + p->ev_card.size();
        const int* d_var = d_ptr + p->evf_ptr.size();
        const int* d_stride = d_var + p->evf_var.size();
        const long long blocks = (B + kThreads - 1) / kThreads;
        jt_evidence_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(evidence, n_evid, d_card, d_ptr, d_var,
                                                                      d_stride, F, B, fbase, err);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        a.fbase = fbase;
    }
    return run_phases(p, B, dtype, workspace, JT_PHASE_INIT, JT_PHASE_INIT, a, vec, stream);
}

int jt_collect(jt_plan* p, int64_t B, int dtype, void* workspace, void* stream) {
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    const int vec = pick_vec(B, dtype);
    return run_phases(p, B, dtype, workspace, JT_PHASE_COLLECT, JT_PHASE_COLLECT,
                      base_args(p, B, dtype, workspace, vec), vec, static_cast<cudaStream_t>(stream));
}

int jt_distribute(jt_plan* p, int64_t B, int dtype, void* workspace, int flags, void* stream) {
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    const int vec = pick_vec(B, dtype);
    KArgs a = base_args(p, B, dtype, workspace, vec);
    a.flags = flags;
    return run_phases(p, B, dtype, workspace, JT_PHASE_DIST_PRE, JT_PHASE_DIST_MAIN, a, vec,
                      static_cast<cudaStream_t>(stream));
}

int jt_marginal(jt_plan* p, int64_t B, int dtype, void* workspace, void* factor_out, void* stream) {
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    if (!factor_out) return fail(JT_ERR_INVALID, "factor_out is null");
    const int vec = pick_vec(B, dtype);
    KArgs a = base_args(p, B, dtype, workspace, vec);
    a.fout = factor_out;
    return run_phases(p, B, dtype, workspace, JT_PHASE_MARGINAL, JT_PHASE_MARGINAL, a, vec,
                      static_cast<cudaStream_t>(stream));
}

int jt_propagate(jt_plan* p, const void* factor_tables, int factors_batched, const int32_t* evidence, int64_t B,
                 int dtype, void* workspace, void* factor_out, int flags, void* stream) {
    int rc = jt_init(p, factor_tables, factors_batched, evidence, B, dtype, workspace, stream);
    if (rc != JT_OK) return rc;
    rc = jt_collect(p, B, dtype, workspace, stream);
    if (rc != JT_OK) return rc;
    rc = jt_distribute(p, B, dtype, workspace, flags, stream);
    if (rc != JT_OK) return rc;
    if (flags & JT_SKIP_MARGINAL) return JT_OK;
    return jt_marginal(p, B, dtype, workspace, factor_out, stream);
}

int jt_evidence_errors(jt_plan* p, int64_t B, int dtype, void* workspace, void* stream, int64_t* out) {
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    if (!out) return fail(JT_ERR_INVALID, "out is null");
    const WorkspaceLayout wl = workspace_layout(p, B, dtype);
    unsigned long long v = 0;
    JT_CUDA(cudaMemcpyAsync(&v, static_cast<char*>(workspace) + wl.err_off, sizeof(v), cudaMemcpyDeviceToHost,
                            static_cast<cudaStream_t>(stream)));
    JT_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    *out = (int64_t)v;
    return JT_OK;
}

int jt_copy_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                 size_t rows, int to_host, void* stream) {
    if (!dst || !src || width_bytes > dst_pitch || width_bytes > src_pitch)
        return fail(JT_ERR_INVALID, "bad argument");
    if (!rows || !width_bytes) return JT_OK;
    JT_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows,
                              to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice,
                              static_cast<cudaStream_t>(stream)));
    return JT_OK;
}

int jt_ratio(const void* new_values, const void* old_values, void* out, int64_t n, int dtype, void* stream_) {
    if (!new_values || !old_values || !out || n < 0) return fail(JT_ERR_INVALID, "bad argument");
    if (dtype != JT_F32 && dtype != JT_F64) return fail(JT_ERR_INVALID, "dtype must be JT_F32 or JT_F64");
    if (n == 0) return JT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    long long blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (dtype == JT_F64)
        jt_ratio_kernel<double><<<(unsigned)blocks, kThreads, 0, stream>>>(
            static_cast<const double*>(new_values), static_cast<const double*>(old_values),
            static_cast<double*>(out), n);
    else
        jt_ratio_kernel<float><<<(unsigned)blocks, kThreads, 0, stream>>>(
            static_cast<const float*>(new_values), static_cast<const float*>(old_values),
            static_cast<float*>(out), n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    JT_CUDA(cudaGetLastError());
    return JT_OK;
}

int jt_contract(const void* const* ops, int n_ops, const int32_t* tables, int64_t n_tab, const int32_t* maps,
                int64_t n_s, int64_t n_r, int64_t n_slo, int64_t n_rlo, int64_t B, int dtype, void* out,
                void* stream_) {
    if (!ops || n_ops <= 0 || !tables || !maps || !out) return fail(JT_ERR_INVALID, "null argument");
    if (n_s <= 0 || n_r <= 0 || n_slo <= 0 || n_rlo <= 0 || n_s % n_slo || n_r % n_rlo || n_s > 2147483647LL ||
        n_r > 2147483647LL || B <= 0 || n_tab <= 0)
        return fail(JT_ERR_INVALID, "bad index-space sizes");
    if (dtype != JT_F32 && dtype != JT_F64) return fail(JT_ERR_INVALID, "dtype must be JT_F32 or JT_F64");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t w = dtype_size(dtype);
    const int n_shi = (int)(n_s / n_slo), n_rhi = (int)(n_r / n_rlo);
    for (int j = 0; j < n_ops; ++j) {
        const int32_t* m = maps + 4 * j;
        if (m[0] < 0 || m[0] + n_shi > n_tab || m[1] < 0 || m[1] + n_slo > n_tab || m[2] < 0 ||
            m[2] + n_rhi > n_tab || m[3] < 0 || m[3] + n_rlo > n_tab)
            return fail(JT_ERR_INVALID, "operand %d: table range", j);
    }
    int vec = pick_vec(B, dtype);
    auto misaligned = [&](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % (w * vec) != 0; };
    while (vec > 1) {
        bool bad = misaligned(out);
        for (int j = 0; j < n_ops; ++j) bad = bad || misaligned(ops[j]);
        if (!bad) break;
        vec >>= 1;
    }
    // device scratch: [DTask | DMsg x n_ops | prefix(2) | tables]
    const size_t msg_off = align_up(sizeof(DTask), 16);
    const size_t prefix_off = align_up(msg_off + sizeof(DMsg) * n_ops, 16);
    const size_t tab_off = align_up(prefix_off + 2 * sizeof(int), 16);
    const size_t total = tab_off + (size_t)n_tab * 4;
    std::vector<char> host(total, 0);
    int bx_log2, sy_log2;
    pick_tile(B / vec, bx_log2, sy_log2);
    DTask t;
    memset(&t, 0, sizeof(t));
    t.src = t.beta = t.bel = t.own = -1;
    t.out = 0;
    t.n_s = (int)n_s; t.n_r = (int)n_r; t.n_slo = (int)n_slo; t.n_rlo = (int)n_rlo;
    // no src operand: point the (unused) src maps at operand 0's tables so every read is in range
    t.src_shi = maps[0]; t.src_slo = maps[1]; t.src_rhi = maps[2]; t.src_rlo = maps[3];
    t.rmsg_begin = 0; t.rmsg_end = n_ops; t.smsg_begin = t.smsg_end = n_ops;
    t.kind = JT_KIND_PROJECT; t.out_space = 1;
    memcpy(host.data(), &t, sizeof(t));
    const char* out_c = static_cast<const char*>(out);
    for (int j = 0; j < n_ops; ++j) {
        DMsg m;
        memset(&m, 0, sizeof(m));
        const ptrdiff_t delta = static_cast<const char*>(ops[j]) - out_c;
        if (delta % (ptrdiff_t)w) return fail(JT_ERR_INVALID, "operand %d misaligned relative to out", j);
        m.eoff = delta / (ptrdiff_t)w;
        m.a_hi = maps[4 * j]; m.a_lo = maps[4 * j + 1]; m.b_hi = maps[4 * j + 2]; m.b_lo = maps[4 * j + 3];
        m.fid = -1;
        memcpy(host.data() + msg_off + sizeof(DMsg) * j, &m, sizeof(m));
    }
    const long long blocks = ((long long)n_s + (1 << sy_log2) - 1) >> sy_log2;
    int prefix[2] = {0, (int)blocks};
    memcpy(host.data() + prefix_off, prefix, sizeof(prefix));
    memcpy(host.data() + tab_off, tables, (size_t)n_tab * 4);
    char* dev = nullptr;
    JT_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dev), total, stream));
    cudaError_t e = cudaMemcpyAsync(dev, host.data(), total, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) {
        cudaFreeAsync(dev, stream);
        return fail(JT_ERR_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
    }
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.tasks = reinterpret_cast<const DTask*>(dev);
    a.msgs = reinterpret_cast<const DMsg*>(dev + msg_off);
    a.prefix = reinterpret_cast<const int*>(dev + prefix_off);
    a.tab = reinterpret_cast<const int*>(dev + tab_off);
    a.work = out;   // operands are addressed relative to `out` (DMsg::eoff)
    a.fout = out;
    a.B = B;
    a.Bv = B / vec;
    a.n_tasks = 1;
    a.bx_log2 = bx_log2;
    a.sy_log2 = sy_log2;
    const long long gy = (a.Bv + (1LL << bx_log2) - 1) >> bx_log2;
    int rc = JT_OK;
    if (blocks > 2147483647LL || gy > 65535) {
        rc = fail(JT_ERR_INVALID, "launch grid exceeds CUDA limits; split the batch");
    } else {
        dim3 grid((unsigned)blocks, (unsigned)gy, 1);
        if (dtype == JT_F64) {
            if (vec == 2) jt_project_kernel<double, 2><<<grid, kThreads, 0, stream>>>(a);
            else jt_project_kernel<double, 1><<<grid, kThreads, 0, stream>>>(a);
        } else {
            if (vec == 4) jt_project_kernel<float, 4><<<grid, kThreads, 0, stream>>>(a);
            else if (vec == 2) jt_project_kernel<float, 2><<<grid, kThreads, 0, stream>>>(a);
            else jt_project_kernel<float, 1><<<grid, kThreads, 0, stream>>>(a);
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(JT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    }
    cudaFreeAsync(dev, stream);
    return rc;
}

}  // extern "C"
