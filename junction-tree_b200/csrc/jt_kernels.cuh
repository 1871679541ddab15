// Device side of libjt_b200: mbarrier/TMA primitives, semirings and the sm_100a kernels.
// Included through jt_launch.cuh by one translation unit per semiring (jt_sr_*.cu).
// Task semantics: include/jt_b200.h; schedule: junctiontree/schedule.py.
//
// Everything on this path is HBM-bound (2 flops per 8..16 bytes), so the kernels are built
// around coalesced 16-byte accesses on the batch-innermost layout [entry][B]; all index
// arithmetic is table-driven and warp-uniform.
#pragma once

#include <type_traits>

#include "jt_host.h"

namespace {

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

// streaming store (evict-first): for data written once and not re-read by the same launch
template <typename T, int VEC>
__device__ __forceinline__ void st_stream(T* p, const Pack<T, VEC>& x) {
    if constexpr (sizeof(T) * VEC == 16) {
        __stcs(reinterpret_cast<float4*>(p), *reinterpret_cast<const float4*>(&x));
    } else if constexpr (sizeof(T) * VEC == 8) {
        __stcs(reinterpret_cast<float2*>(p), *reinterpret_cast<const float2*>(&x));
    } else {
        __stcs(reinterpret_cast<float*>(p), *reinterpret_cast<const float*>(&x));
    }
}

// ------------------------------------------------------------------------------------------
// Semirings ("distributive laws", reference sum_product.py:2-3 and junctiontree.py:300-305).
// A semiring supplies the product (x), the running reduction (+) through an accumulator type,
// and the output-stage operations.  SrSumProduct generates exactly the multiply / add code the
// kernels had before they were parameterised.
//   sum-product  (R, +, *)        marginals, partition function
//   max-product  (R>=0, max, *)   max-marginals (MAP)
//   log-sum-exp  (R, logaddexp, +) sum-product on log potentials, underflow-safe
//   max-sum      (R, max, +)      max-product on log potentials

struct SrSumProduct {
    template <typename T> struct Acc { T v; };
    template <typename T> __device__ __forceinline__ static T one() { return T(1); }
    template <typename T> __device__ __forceinline__ static T mul(T a, T b) { return a * b; }
    template <typename T> __device__ __forceinline__ static T add(T a, T b) { return a + b; }
    template <typename T> __device__ __forceinline__ static Acc<T> acc_zero() { return {T(0)}; }
    template <typename T> __device__ __forceinline__ static void accum(Acc<T>& a, T v) { a.v += v; }
    template <typename T> __device__ __forceinline__ static void merge(Acc<T>& a, const Acc<T>& b) { a.v += b.v; }
    template <typename T> __device__ __forceinline__ static T finish(const Acc<T>& a) { return a.v; }
    // output stage: x (/) z with the (+)-identity mapped to itself, and log of a total
    template <typename T> __device__ __forceinline__ static T unit(T x, T z) { return z > T(0) ? x / z : T(0); }
    template <typename T> __device__ __forceinline__ static T log_of(T z) { return log(z); }
};

struct SrMaxProduct {
    template <typename T> struct Acc { T v; };
    template <typename T> __device__ __forceinline__ static T one() { return T(1); }
    template <typename T> __device__ __forceinline__ static T mul(T a, T b) { return a * b; }
    template <typename T> __device__ __forceinline__ static T add(T a, T b) { return a > b ? a : b; }
    template <typename T> __device__ __forceinline__ static Acc<T> acc_zero() { return {T(-INFINITY)}; }
    template <typename T> __device__ __forceinline__ static void accum(Acc<T>& a, T v) { a.v = v > a.v ? v : a.v; }
    template <typename T> __device__ __forceinline__ static void merge(Acc<T>& a, const Acc<T>& b) { a.v = b.v > a.v ? b.v : a.v; }
    template <typename T> __device__ __forceinline__ static T finish(const Acc<T>& a) { return a.v; }
    template <typename T> __device__ __forceinline__ static T unit(T x, T z) { return z > T(0) ? x / z : T(0); }
    template <typename T> __device__ __forceinline__ static T log_of(T z) { return log(z); }
};

struct SrMaxSum {
    template <typename T> struct Acc { T v; };
    template <typename T> __device__ __forceinline__ static T one() { return T(0); }
    template <typename T> __device__ __forceinline__ static T mul(T a, T b) { return a + b; }
    template <typename T> __device__ __forceinline__ static T add(T a, T b) { return a > b ? a : b; }
    template <typename T> __device__ __forceinline__ static Acc<T> acc_zero() { return {T(-INFINITY)}; }
    template <typename T> __device__ __forceinline__ static void accum(Acc<T>& a, T v) { a.v = v > a.v ? v : a.v; }
    template <typename T> __device__ __forceinline__ static void merge(Acc<T>& a, const Acc<T>& b) { a.v = b.v > a.v ? b.v : a.v; }
    template <typename T> __device__ __forceinline__ static T finish(const Acc<T>& a) { return a.v; }
    template <typename T> __device__ __forceinline__ static T unit(T x, T z) { return z > T(-INFINITY) ? x - z : T(-INFINITY); }
    template <typename T> __device__ __forceinline__ static T log_of(T z) { return z; }
};

// exp(x) for x <= 0 in double precision, 2 ulp (checked against libm over [-700, 0]): table of
// 2^(j/32) times a degree-6 polynomial on |r| <= ln2/64, about half the FP64 instructions of the
// library exp -- the float64 log-sum-exp kernels are bound by the FP64 pipe, not by memory.
// Below -700 the term is < 1e-304 of the running maximum and is dropped.
__device__ const double kExp2Over32[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002,
};

__device__ __forceinline__ double exp_nonpos(double x) {
    if (x < -700.0) return 0.0;
    const double k = rint(x * 46.16624130844683);                            // 32 / ln 2
    const double r = fma(-k, 4.06140840434059e-10, fma(-k, 0.02166084898635745, x));   // x - k ln2/32 (hi, lo)
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int ki = (int)k;
    const double t = __ldg(kExp2Over32 + (ki & 31)) * p;                     // in [0.99, 2)
    return __longlong_as_double(__double_as_longlong(t) + ((long long)(ki >> 5) << 52));   // * 2^(k / 32)
}
__device__ __forceinline__ float exp_nonpos(float x) { return expf(x); }

// log-sum-exp: the running reduction keeps (max m, sum s of exp(v - m)), one exp per term
struct SrLogSumExp {
    template <typename T> struct Acc { T m, s; };
    template <typename T> __device__ __forceinline__ static T one() { return T(0); }
    template <typename T> __device__ __forceinline__ static T mul(T a, T b) { return a + b; }
    template <typename T> __device__ __forceinline__ static T add(T a, T b) {
        const T m = a > b ? a : b, n = a > b ? b : a;
        return n > T(-INFINITY) ? m + log1p(exp(n - m)) : m;
    }
    template <typename T> __device__ __forceinline__ static Acc<T> acc_zero() { return {T(-INFINITY), T(0)}; }
    // one exp per term, taken before the (predicated) case split so that lanes whose terms fall
    // on different sides of the running maximum do not serialise two exp sequences
    template <typename T> __device__ __forceinline__ static void accum(Acc<T>& a, T v) {
        const T d = v - a.m;                    // NaN only when both are -inf: nothing to add
        const T e = exp_nonpos(-fabs(d));       // in [0, 1]
        if (d > T(0)) {                         // new maximum (d = +inf when the accumulator is empty: e = 0)
            a.s = a.s * e + T(1);
            a.m = v;
        } else if (d <= T(0)) {                 // v = -inf gives e = 0
            a.s += e;
        }
    }
    template <typename T> __device__ __forceinline__ static void merge(Acc<T>& a, const Acc<T>& b) {
        const T d = b.m - a.m;
        const T e = exp_nonpos(-fabs(d));
        if (d > T(0)) {
            a.s = a.s * e + b.s;
            a.m = b.m;
        } else if (d <= T(0)) {
            a.s += b.s * e;
        }
    }
    template <typename T> __device__ __forceinline__ static T finish(const Acc<T>& a) {
        return a.s > T(0) ? a.m + log(a.s) : T(-INFINITY);
    }
    template <typename T> __device__ __forceinline__ static T unit(T x, T z) { return z > T(-INFINITY) ? x - z : T(-INFINITY); }
    template <typename T> __device__ __forceinline__ static T log_of(T z) { return z; }
};

template <typename SR, typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> pack_one() {
    return pack_fill<T, VEC>(SR::template one<T>());
}

template <typename SR, typename T, int VEC>
__device__ __forceinline__ void mul(Pack<T, VEC>& a, const Pack<T, VEC>& b) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) a.v[i] = SR::mul(a.v[i], b.v[i]);
}

// accumulator of one batch vector
template <typename SR, typename T, int VEC>
struct AccPack {
    typename SR::template Acc<T> a[VEC];
};

template <typename SR, typename T, int VEC>
__device__ __forceinline__ AccPack<SR, T, VEC> acc_zero() {
    AccPack<SR, T, VEC> x;
#pragma unroll
    for (int i = 0; i < VEC; ++i) x.a[i] = SR::template acc_zero<T>();
    return x;
}

template <typename SR, typename T, int VEC>
__device__ __forceinline__ void accum(AccPack<SR, T, VEC>& a, const Pack<T, VEC>& v) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) SR::accum(a.a[i], v.v[i]);
}

template <typename SR, typename T, int VEC>
__device__ __forceinline__ void merge(AccPack<SR, T, VEC>& a, const AccPack<SR, T, VEC>& b) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) SR::merge(a.a[i], b.a[i]);
}

template <typename SR, typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> finish(const AccPack<SR, T, VEC>& a) {
    Pack<T, VEC> p;
#pragma unroll
    for (int i = 0; i < VEC; ++i) p.v[i] = SR::finish(a.a[i]);
    return p;
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
// clique initialisation (E0 + V1): psi_C[s][b] = prod_f phi_f[ A_f(s) + fbase[f][b] ]

// A thread owns VEC batch columns and walks its rows of the block's chunk of s, so the
// per-instance factor offsets (which do not depend on s) are read once and kept in registers.
constexpr int kInitRegFactors = 6;

template <typename SR, typename T, int VEC>
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
    const T* lik = static_cast<const T*>(a.work);        // likelihood tables (fid -2) live in the workspace
    const int f0 = tk->smsg_begin, nf = tk->smsg_end - f0;
    const bool gather = !a.fin_batched && a.fbase != nullptr;

    // per-factor, per-instance base offsets (evidence slicing), resident in registers
    int fb[kInitRegFactors][VEC];
#pragma unroll
    for (int j = 0; j < kInitRegFactors; ++j) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) fb[j][u] = 0;
        if (gather && j < nf && msgs[f0 + j].fid >= 0) {
            const int* p = a.fbase + (long long)msgs[f0 + j].fid * B + col;
#pragma unroll
            for (int u = 0; u < VEC; ++u) fb[j][u] = p[u];
        }
    }

    T* out = static_cast<T*>(a.work) + tk->out * B + col;
    for (; s < s_end; s += rows) {
        P val = pack_one<SR, T, VEC>();
#pragma unroll
        for (int j = 0; j < kInitRegFactors; ++j) {
            if (j < nf) {
                const DMsg* m = msgs + f0 + j;
                const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
                if (m->fid < 0) {
                    mul<SR>(val, ld<T, VEC>(lik + idx * B + col));
                } else if (a.fin_batched) {
                    mul<SR>(val, ld<T, VEC>(fin + idx * B + col));
                } else {
#pragma unroll
                    for (int u = 0; u < VEC; ++u) val.v[u] = SR::mul(val.v[u], __ldg(fin + idx + fb[j][u]));
                }
            }
        }
        for (int j = kInitRegFactors; j < nf; ++j) {   // rare: many factors in one clique
            const DMsg* m = msgs + f0 + j;
            const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
            if (m->fid < 0) {
                mul<SR>(val, ld<T, VEC>(lik + idx * B + col));
            } else if (a.fin_batched) {
                mul<SR>(val, ld<T, VEC>(fin + idx * B + col));
            } else {
                const int* p = a.fbase ? a.fbase + (long long)m->fid * B + col : nullptr;
#pragma unroll
                for (int u = 0; u < VEC; ++u) val.v[u] = SR::mul(val.v[u], __ldg(fin + idx + (p ? p[u] : 0)));
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

// Batches of 64 vectors and more (a block spans 1-4 rows of s, 256-64 batch vectors wide): the
// table lookups of a row are the same for every thread of it, so they are done once per block -- thread t resolves row
// base + t of a 256-row sub-chunk into shared memory -- and the row loop is left with one
// shared-memory read, the gathers, the products and the store (ncu, round 1: the per-thread
// form was issue-bound at 78 % issue-active and 4.4 TB/s of a 7.4 TB/s write ceiling).
template <typename SR, typename T, int VEC>
__global__ void __launch_bounds__(kThreads) jt_init_rows_kernel(const KArgs a) {
    typedef Pack<T, VEC> P;
    __shared__ long long sidx[kInitRegFactors][kThreads];
    int s0;
    const DTask* tk = locate_chunk(a, s0);
    const int n_s = tk->n_s;
    const int s_end = min(n_s, s0 + (1 << a.sy_log2));
    const int t = threadIdx.x;
    // 2^bx_log2 batch vectors per row of the block, the remaining thread bits walk rows side by side
    const int lanes = kThreads >> a.bx_log2, tl = t >> a.bx_log2;
    const long long bv = ((long long)blockIdx.y << a.bx_log2) + (t & ((1 << a.bx_log2) - 1));
    const bool active = bv < a.Bv;
    const int* __restrict__ tab = a.tab;
    const DMsg* __restrict__ msgs = a.msgs;
    const long long B = a.B, col = (active ? bv : 0) * VEC;
    const int n_slo = tk->n_slo;
    const T* __restrict__ fin = static_cast<const T*>(a.fin);
    const T* lik = static_cast<const T*>(a.work);
    const int f0 = tk->smsg_begin, nf = tk->smsg_end - f0;
    const int nreg = nf < kInitRegFactors ? nf : kInitRegFactors;
    const bool gather = !a.fin_batched && a.fbase != nullptr;

    // per factor: how it is read (0 gather from the shared table, 1 per-instance row) and the
    // per-instance base offsets of the gathers (evidence slicing), resident in registers
    int fb[kInitRegFactors][VEC];
    bool row_op[kInitRegFactors];
    const T* row_base[kInitRegFactors];
#pragma unroll
    for (int j = 0; j < kInitRegFactors; ++j) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) fb[j][u] = 0;
        row_op[j] = false;
        row_base[j] = fin;
        if (j < nf) {
            const int fid = msgs[f0 + j].fid;
            row_op[j] = fid < 0 || a.fin_batched;
            row_base[j] = fid < 0 ? lik : fin;
            if (gather && fid >= 0) {
                const int* p = a.fbase + (long long)fid * B + col;
#pragma unroll
                for (int u = 0; u < VEC; ++u) fb[j][u] = p[u];
            }
        }
    }

    T* out = static_cast<T*>(a.work) + tk->out * B + col;
    // every factor of the clique is gathered from small shared tables (the usual case: evidence
    // slicing, no likelihood rows): a row loop without the row-operand alternative -- the compiler
    // predicates both forms otherwise and a factor costs ~30 issue slots either way -- and with
    // 32-bit gather indices (ncu r02: Ising grid, 3 factors per clique, 4.4 of ~7 TB/s)
    bool gathers_only = (a.flags & JT_X_FIN32) && nf <= kInitRegFactors;
#pragma unroll
    for (int j = 0; j < kInitRegFactors; ++j) gathers_only = gathers_only && !(j < nf && row_op[j]);
    for (int base = s0; base < s_end; base += kThreads) {
        const int nrows = min(kThreads, s_end - base);
        __syncthreads();                                  // the previous sub-chunk has been consumed
        if (t < nrows) {
            const int s = base + t;
            int s_hi = 0, s_lo = s;
            if (n_slo < n_s) {
                s_hi = s / n_slo;
                s_lo = s - s_hi * n_slo;
            }
            for (int j = 0; j < nreg; ++j) {
                const DMsg* m = msgs + f0 + j;
                sidx[j][t] = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
            }
        }
        __syncthreads();
        if (!active) continue;
        if (gathers_only) {
            // one instantiation of the row loop per factor count: no predicated-off factors in it
            auto walk = [&](auto nfc) {
                constexpr int NF = decltype(nfc)::value;
#pragma unroll 2
                for (int i = tl; i < nrows; i += lanes) {
                    P val = pack_one<SR, T, VEC>();
#pragma unroll
                    for (int j = 0; j < NF; ++j) {
                        const unsigned idx = (unsigned)sidx[j][i];
#pragma unroll
                        for (int u = 0; u < VEC; ++u) val.v[u] = SR::mul(val.v[u], __ldg(fin + (idx + (unsigned)fb[j][u])));
                    }
                    st<T, VEC>(out + (long long)(base + i) * B, val);
                }
            };
            switch (nf) {
            case 0: walk(std::integral_constant<int, 0>()); break;
            case 1: walk(std::integral_constant<int, 1>()); break;
            case 2: walk(std::integral_constant<int, 2>()); break;
            case 3: walk(std::integral_constant<int, 3>()); break;
            case 4: walk(std::integral_constant<int, 4>()); break;
            case 5: walk(std::integral_constant<int, 5>()); break;
            default: walk(std::integral_constant<int, 6>()); break;
            }
            continue;
        }
#pragma unroll 2
        for (int i = tl; i < nrows; i += lanes) {
            P val = pack_one<SR, T, VEC>();
#pragma unroll
            for (int j = 0; j < kInitRegFactors; ++j) {
                if (j < nf) {
                    const long long idx = sidx[j][i];
                    if (row_op[j]) {
                        mul<SR>(val, ld<T, VEC>(row_base[j] + idx * B + col));
                    } else {
#pragma unroll
                        for (int u = 0; u < VEC; ++u) val.v[u] = SR::mul(val.v[u], __ldg(fin + idx + fb[j][u]));
                    }
                }
            }
            if (nf > kInitRegFactors) {                   // rare: many factors in one clique
                const int s = base + i;
                int s_hi = 0, s_lo = s;
                if (n_slo < n_s) {
                    s_hi = s / n_slo;
                    s_lo = s - s_hi * n_slo;
                }
                for (int j = kInitRegFactors; j < nf; ++j) {
                    const DMsg* m = msgs + f0 + j;
                    const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
                    if (m->fid < 0) {
                        mul<SR>(val, ld<T, VEC>(lik + idx * B + col));
                    } else if (a.fin_batched) {
                        mul<SR>(val, ld<T, VEC>(fin + idx * B + col));
                    } else {
                        const int* p = a.fbase ? a.fbase + (long long)m->fid * B + col : nullptr;
#pragma unroll
                        for (int u = 0; u < VEC; ++u) val.v[u] = SR::mul(val.v[u], __ldg(fin + idx + (p ? p[u] : 0)));
                    }
                }
            }
            st<T, VEC>(out + (long long)(base + i) * B, val);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Clique beliefs in uniform mode (E5, reference computation.py:216-224, for a potential shared by
// the batch): beta_C[e][b] = U(e) * prod_j row_j[idx_j(e)][b] -- U = psi_C x the uniform messages
// (scalars from the uniform workspace), rows = the per-instance messages (at most kBetaRows).
// Nothing is reduced, so this is a streaming write like the clique initialisation and has the
// same structure as jt_init_rows_kernel: thread t resolves item base + t of a 256-item sub-chunk
// (scalar, row indices, destination row) into shared memory, then every thread owns a batch
// vector and walks the sub-chunk with a few 16-byte loads, products and one 16-byte store per
// entry.  (The projection kernels did this as a side effect of the message tasks at 3.6-5.5 TB/s
// -- issue-bound consumers, ncu r02 -- against ~7 TB/s for streaming writes.)

struct BetaArgs {
    const int* list;       // [n] task ids of this launch, then [n] offsets of their walk-order tables (jt_plan::dtab, -1: none)
    const int* perm;       // walk order of the tasks that have one: position -> (s, r) item
    const int* prefix;     // [n + 1] first block of each task for 2^ch_log2 items per block
    int n, ch_log2;
    const DTask* tasks;    // all tasks of the plan
    const DMsg* msgs;
    const int* tab;
    void* work;
    const void* uni;
    long long B, Bv;
    int bx_log2;
};

template <typename SR, typename T, int VEC>
__global__ void __launch_bounds__(kThreads, 3) jt_beta_kernel(const BetaArgs a) {
    typedef Pack<T, VEC> P;
    __shared__ T su[kThreads];
    __shared__ int se[kThreads];
    __shared__ int srow[kBetaRows][kThreads];
    int lo = 0, hi = a.n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.prefix + mid) <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    const DTask* tk = a.tasks + __ldg(a.list + lo);
    // Walk order.  The items are visited s-major, r-minor by default: the rows that depend on s only
    // (own message, s-only messages) stay put for n_r consecutive entries and the r-dependent ones
    // cycle through a handful of rows, so all of them are served by L1 and the launch is bound by
    // its scattered 16-byte-per-thread row writes (~6 TB/s measured; walking the clique in memory
    // order instead re-reads a row from L2 per entry and runs at half that).  A task without an r
    // space (a leaf clique: every entry has its own s) gets a table that visits the entries
    // sorted by their row operand, for the same effect.
    const int perm_off = __ldg(a.list + a.n + lo);
    const int* __restrict__ perm = perm_off >= 0 ? a.perm + perm_off : nullptr;
    const int n_r = tk->n_r, n_slo = tk->n_slo, n_rlo = tk->n_rlo;
    const long long n_items = (long long)tk->n_s * n_r;
    const long long i0 = (long long)((int)blockIdx.x - __ldg(a.prefix + lo)) << a.ch_log2;
    const long long i1 = i0 + (1LL << a.ch_log2) < n_items ? i0 + (1LL << a.ch_log2) : n_items;
    const int t = threadIdx.x;
    const int tx = t & ((1 << a.bx_log2) - 1), ty = t >> a.bx_log2, ry = kThreads >> a.bx_log2;
    const long long bv = ((long long)blockIdx.y << a.bx_log2) + tx;
    const bool active = bv < a.Bv;
    const long long B = a.B, col = (active ? bv : 0) * VEC;
    const int* __restrict__ tab = a.tab;
    const DMsg* __restrict__ msgs = a.msgs;
    const T* __restrict__ uni = static_cast<const T*>(a.uni);
    T* work = static_cast<T*>(a.work);
    const int tflags = tk->flags;
    const bool own_row = tk->own >= 0 && !(tflags & JT_TF_OWN_UNIFORM);
    int n_rows = own_row ? 1 : 0;
    for (int j = tk->rmsg_begin; j < tk->smsg_end; ++j) n_rows += msgs[j].uni ? 0 : 1;
    T* beta = work + tk->beta * B + col;

    // U entries per trip: all row loads first, then the products and stores -- loads and stores go
    // through the same workspace pointer, so the compiler keeps their order and only loads that
    // are issued back to back are in flight together (8 x 16 bytes per thread)
    // U entries per trip: all row loads first, then the products and stores -- loads and stores go
    // through the same workspace pointer, so the compiler keeps their order and only loads that
    // are issued back to back are in flight together (8 x 16 bytes per thread).  Consecutive
    // entries mostly reuse the same rows (see the walk order below), so the loads hit L1.
    auto rows = [&](auto n_tag, int n) {
        constexpr int N = decltype(n_tag)::value;
        constexpr int U = N <= 1 ? 8 : 4;
        int i = ty;
        for (; i + (U - 1) * ry < n; i += ry * U) {          // full trips: no bounds checks
            P x[U][N > 0 ? N : 1];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < N; ++k) x[u][k] = ld<T, VEC>(work + (long long)srow[k][i + u * ry] * B + col);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                P val = pack_fill<T, VEC>(su[i + u * ry]);
#pragma unroll
                for (int k = 0; k < N; ++k) mul<SR>(val, x[u][k]);
                st_stream<T, VEC>(beta + (long long)se[i + u * ry] * B, val);
            }
        }
        for (; i < n; i += ry) {
            P val = pack_fill<T, VEC>(su[i]);
#pragma unroll
            for (int k = 0; k < N; ++k) mul<SR>(val, ld<T, VEC>(work + (long long)srow[k][i] * B + col));
            st_stream<T, VEC>(beta + (long long)se[i] * B, val);
        }
    };

    for (long long base = i0; base < i1; base += kThreads) {
        const int n = (int)(i1 - base < kThreads ? i1 - base : kThreads);
        __syncthreads();                                  // the previous sub-chunk has been consumed
        if (t < n) {
            const long long item = perm ? (long long)__ldg(perm + base + t) : base + t;
            const int s = (int)(item / n_r), r = (int)(item - (long long)s * n_r);
            int s_hi = 0, s_lo = s;
            if (n_slo < tk->n_s) {
                s_hi = s / n_slo;
                s_lo = s - s_hi * n_slo;
            }
            const int rh = r / n_rlo, rl = r - rh * n_rlo;
            const int e = __ldg(tab + tk->src_shi + s_hi) + __ldg(tab + tk->src_slo + s_lo) +
                          __ldg(tab + tk->src_rhi + rh) + __ldg(tab + tk->src_rlo + rl);
            T u = __ldg(uni + tk->src + e);
            int k = 0;
            for (int j = tk->rmsg_begin; j < tk->smsg_end; ++j) {
                const DMsg* m = msgs + j;
                long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
                if (j < tk->rmsg_end) idx += __ldg(tab + m->b_hi + rh) + __ldg(tab + m->b_lo + rl);
                if (m->uni) u = SR::mul(u, __ldg(uni + idx));
                else srow[k++][t] = (int)idx;
            }
            if (tk->own >= 0) {
                if (own_row) srow[k++][t] = (int)(tk->own + s);
                else u = SR::mul(u, __ldg(uni + tk->own + s));
            }
            su[t] = u;
            se[t] = e;
        }
        __syncthreads();
        if (!active) continue;
        switch (n_rows) {
            case 0: rows(std::integral_constant<int, 0>{}, n); break;
            case 1: rows(std::integral_constant<int, 1>{}, n); break;
            case 2: rows(std::integral_constant<int, 2>{}, n); break;
            default: rows(std::integral_constant<int, 3>{}, n); break;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Scalar tasks (uniform mode): a projection task all of whose r-dependent inputs are uniform,
//     out[s][b] = ( sum_r U(s, r) ) * prod_j smsg_j[A_j(s)][b]        bel[s][b] = out * own[s][b]
// The total over r is the same for every instance: it is computed once, in the uniform workspace
// (an ordinary B = 1 projection launch derived when the plan is loaded, jt_dense.cu), and the
// batch launch is left with n_s rows of elementwise work instead of n_s * n_r items of scalar
// bookkeeping (config 4: 786,432 items per task, 0.41-0.44 ms, for 8,192 output rows).

struct ScalarArgs {
    const int* list;       // [n] task ids, [n] entry of each task's totals in the uniform workspace, [n + 1] block prefix
    int n;
    const DTask* tasks;    // all tasks of the plan
    const DMsg* msgs;
    const int* tab;
    void* work;
    const void* uni;
    void* fout;
    long long B, Bv;
    int bx_log2, flags;
};

template <typename SR, typename T, int VEC>
__global__ void __launch_bounds__(kThreads) jt_scalar_kernel(const ScalarArgs a) {
    typedef Pack<T, VEC> P;
    __shared__ T su[kScalarRows];
    __shared__ T sown[kScalarRows];
    __shared__ int srow[kBetaRows + 1][kScalarRows];
    const int* prefix = a.list + 2 * a.n;
    int lo = 0, hi = a.n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    const DTask* tk = a.tasks + __ldg(a.list + lo);
    const long long total0 = __ldg(a.list + a.n + lo);
    const int s0 = ((int)blockIdx.x - __ldg(prefix + lo)) * kScalarRows;
    const int n = tk->n_s - s0 < kScalarRows ? tk->n_s - s0 : kScalarRows;
    const int t = threadIdx.x;
    const int tx = t & ((1 << a.bx_log2) - 1), ty = t >> a.bx_log2, ry = kThreads >> a.bx_log2;
    const long long bv = ((long long)blockIdx.y << a.bx_log2) + tx;
    const long long B = a.B, col = bv * VEC;
    const int* __restrict__ tab = a.tab;
    const DMsg* __restrict__ msgs = a.msgs;
    const T* __restrict__ uni = static_cast<const T*>(a.uni);
    T* work = static_cast<T*>(a.work);
    const int tflags = tk->flags;
    const bool wbel = tk->bel >= 0 && (a.flags & JT_SEP_BELIEFS);
    const bool own_row = wbel && tk->own >= 0 && !(tflags & JT_TF_OWN_UNIFORM);
    int n_rows = 0;
    for (int j = tk->smsg_begin; j < tk->smsg_end; ++j) n_rows += msgs[j].uni ? 0 : 1;
    if (t < n) {
        const int s = s0 + t;
        int s_hi = 0, s_lo = s;
        if (tk->n_slo < tk->n_s) {
            s_hi = s / tk->n_slo;
            s_lo = s - s_hi * tk->n_slo;
        }
        su[t] = __ldg(uni + total0 + s);
        int k = 0;
        for (int j = tk->smsg_begin; j < tk->smsg_end; ++j) {
            const DMsg* m = msgs + j;
            if (!m->uni) srow[k++][t] = (int)(m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo));
        }
        sown[t] = SR::template one<T>();
        if (wbel && tk->own >= 0) {
            if (own_row) srow[kBetaRows][t] = (int)(tk->own + s);
            else sown[t] = __ldg(uni + tk->own + s);
        }
    }
    __syncthreads();
    if (bv >= a.Bv) return;
    T* out = (tk->out_space ? static_cast<T*>(a.fout) : work) + (tk->out + s0) * B + col;
    T* bel = work + ((wbel ? tk->bel : 0) + s0) * B + col;
    // four rows per trip: every load first (row operands, own), then the products and stores
    constexpr int U = 4;
    for (int i0 = ty; i0 < n; i0 += ry * U) {
        P x[U][kBetaRows], xo[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * ry;
            if (i < n) {
#pragma unroll
                for (int k = 0; k < kBetaRows; ++k)
                    if (k < n_rows) x[u][k] = ld<T, VEC>(work + (long long)srow[k][i] * B + col);
                if (own_row) xo[u] = ld<T, VEC>(work + (long long)srow[kBetaRows][i] * B + col);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * ry;
            if (i < n) {
                P val = pack_fill<T, VEC>(su[i]);
#pragma unroll
                for (int k = 0; k < kBetaRows; ++k)
                    if (k < n_rows) mul<SR>(val, x[u][k]);
                st<T, VEC>(out + (long long)i * B, val);
                if (wbel) {
                    if (own_row) mul<SR>(val, xo[u]);
                    else mul<SR>(val, pack_fill<T, VEC>(sown[i]));
                    st<T, VEC>(bel + (long long)i * B, val);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// projection task (collect E1+E2, distribute E3+E4+M1+E5, marginal E6, contract)
//
// Uniform operands (KArgs::uniform): an operand flagged uniform (the potential of a clique no
// evidence touches, an up-message of an evidence-free subtree) is identical for every instance;
// it lives once in the uniform workspace (same entry offsets, B = 1) and is broadcast.

template <typename SR, typename T, int VEC>
__global__ void __launch_bounds__(kThreads) jt_project_kernel(const KArgs a) {
    typedef Pack<T, VEC> P;
    typedef AccPack<SR, T, VEC> A;
    int s;
    long long bv;
    const DTask* tk = locate(a, s, bv);
    const int n_s = tk->n_s;
    if (s >= n_s || bv >= a.Bv) return;

    const int* __restrict__ tab = a.tab;
    const DMsg* __restrict__ msgs = a.msgs;
    const long long B = a.B, col = bv * VEC;
    T* work = static_cast<T*>(a.work);
    const T* uni = static_cast<const T*>(a.uni);
    const bool um = a.uniform != 0;
    const int tflags = um ? tk->flags : 0;

    const int n_slo = tk->n_slo;
    int s_hi = 0, s_lo = s;
    if (n_slo < n_s) {
        s_hi = s / n_slo;
        s_lo = s - s_hi * n_slo;
    }

    // messages that do not depend on r, and the task's own up-message
    P sm = pack_one<SR, T, VEC>();
    for (int j = tk->smsg_begin; j < tk->smsg_end; ++j) {
        const DMsg* m = msgs + j;
        const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
        if (um && m->uni) mul<SR>(sm, pack_fill<T, VEC>(__ldg(uni + idx)));
        else mul<SR>(sm, ld<T, VEC>(work + m->eoff + idx * B + col));
    }
    const bool has_own = tk->own >= 0;
    P own = pack_one<SR, T, VEC>();
    if (has_own) {
        if (tflags & JT_TF_OWN_UNIFORM) own = pack_fill<T, VEC>(__ldg(uni + tk->own + s));
        else own = ld<T, VEC>(work + (tk->own + s) * B + col);
    }

    // r-dependent messages: the first kRegMsgs are tracked in registers.  mptr/mpitch address
    // either a [n][B] row (pitch B) or the uniform copy (pitch 1, scalar broadcast).
    const int rm0 = tk->rmsg_begin;
    const int nr = tk->rmsg_end - rm0;
    const T* mptr[kRegMsgs];
    int mbhi[kRegMsgs], mblo[kRegMsgs];
    unsigned umask = 0;
#pragma unroll
    for (int j = 0; j < kRegMsgs; ++j) {
        mptr[j] = work;
        mbhi[j] = mblo[j] = 0;
        if (j < nr) {
            const DMsg* m = msgs + rm0 + j;
            const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
            if (um && m->uni) {
                umask |= 1u << j;
                mptr[j] = uni + idx;
            } else {
                mptr[j] = work + m->eoff + idx * B + col;
            }
            mbhi[j] = m->b_hi;
            mblo[j] = m->b_lo;
        }
    }

    const bool has_src = tk->src >= 0;
    const bool src_uni = (tflags & JT_TF_SRC_UNIFORM) != 0;
    const long long s_off = __ldg(tab + tk->src_shi + s_hi) + __ldg(tab + tk->src_slo + s_lo);
    const T* sptr = src_uni ? uni + tk->src + s_off : work + ((has_src ? tk->src : 0) + s_off) * B + col;
    const long long spitch = src_uni ? 1 : B;
    const bool wbeta = tk->beta >= 0 && !(a.flags & JT_NO_BELIEFS) &&
                       !((tflags & JT_TF_BETA_SPLIT) && (a.flags & JT_X_BETA_SPLIT));   // else jt_beta_kernel writes it
    T* bptr = work + ((wbeta ? tk->beta : 0) + s_off) * B + col;
    P scale = sm;
    mul<SR>(scale, own);

    const int n_rlo = tk->n_rlo;
    const int n_rhi = tk->n_r / n_rlo;
    const int* __restrict__ t_rhi = tab + tk->src_rhi;
    const int* __restrict__ t_rlo = tab + tk->src_rlo;

    A acc[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) acc[u] = acc_zero<SR, T, VEC>();

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
            for (int u = 0; u < kUnroll; ++u) e[u] = e_hi + __ldg(t_rlo + rl + u);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                if (!has_src) v[u] = pack_one<SR, T, VEC>();
                else if (src_uni) v[u] = pack_fill<T, VEC>(__ldg(sptr + e[u]));
                else v[u] = ld<T, VEC>(sptr + e[u] * B);
            }
#pragma unroll
            for (int j = 0; j < kRegMsgs; ++j) {
                if (j < nr) {
                    P w[kUnroll];
                    if ((umask >> j) & 1u) {
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u)
                            w[u] = pack_fill<T, VEC>(__ldg(mptr[j] + mh[j] + __ldg(tab + mblo[j] + rl + u)));
                    } else {
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u)
                            w[u] = ld<T, VEC>(mptr[j] + (mh[j] + __ldg(tab + mblo[j] + rl + u)) * B);
                    }
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) mul<SR>(v[u], w[u]);
                }
            }
            for (int j = kRegMsgs; j < nr; ++j) {   // rare: more than kRegMsgs r-dependent messages
                const DMsg* m = msgs + rm0 + j;
                const long long base = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo) +
                                       __ldg(tab + m->b_hi + rh);
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const long long idx = base + __ldg(tab + m->b_lo + rl + u);
                    if (um && m->uni) mul<SR>(v[u], pack_fill<T, VEC>(__ldg(uni + idx)));
                    else mul<SR>(v[u], ld<T, VEC>(work + m->eoff + idx * B + col));
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) accum<SR>(acc[u], v[u]);
            if (wbeta) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    mul<SR>(v[u], scale);
                    st<T, VEC>(bptr + e[u] * B, v[u]);
                }
            }
        }
        for (; rl < n_rlo; ++rl) {
            const long long e = e_hi + __ldg(t_rlo + rl);
            P v;
            if (!has_src) v = pack_one<SR, T, VEC>();
            else if (src_uni) v = pack_fill<T, VEC>(__ldg(sptr + e));
            else v = ld<T, VEC>(sptr + e * spitch);
#pragma unroll
            for (int j = 0; j < kRegMsgs; ++j) {
                if (j < nr) {
                    const long long d = mh[j] + __ldg(tab + mblo[j] + rl);
                    if ((umask >> j) & 1u) mul<SR>(v, pack_fill<T, VEC>(__ldg(mptr[j] + d)));
                    else mul<SR>(v, ld<T, VEC>(mptr[j] + d * B));
                }
            }
            for (int j = kRegMsgs; j < nr; ++j) {
                const DMsg* m = msgs + rm0 + j;
                const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo) +
                                      __ldg(tab + m->b_hi + rh) + __ldg(tab + m->b_lo + rl);
                if (um && m->uni) mul<SR>(v, pack_fill<T, VEC>(__ldg(uni + idx)));
                else mul<SR>(v, ld<T, VEC>(work + m->eoff + idx * B + col));
            }
            accum<SR>(acc[0], v);
            if (wbeta) {
                mul<SR>(v, scale);
                st<T, VEC>(bptr + e * B, v);
            }
        }
    }

    if (tk->out >= 0) {
        // pairwise combination of the partial sums
        merge<SR>(acc[0], acc[1]);
        merge<SR>(acc[2], acc[3]);
        merge<SR>(acc[0], acc[2]);
        P o = finish<SR>(acc[0]);
        mul<SR>(o, sm);
        T* obase = tk->out_space ? static_cast<T*>(a.fout) : work;
        st<T, VEC>(obase + (tk->out + s) * B + col, o);
        if (tk->bel >= 0 && (a.flags & JT_SEP_BELIEFS)) {
            mul<SR>(o, own);
            st<T, VEC>(work + (tk->bel + s) * B + col, o);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Split-r projection for small batches (same task semantics as jt_project_kernel).
//
// With few instances and a long reduction (a single propagation marginalising a 2^17-entry
// clique to a 2-entry factor scope) the s-parallel kernels leave all but a handful of threads
// idle.  Here a block owns one output index s; its threads are laid out as 2^bx_log2 batch lanes
// times 256 / 2^bx_log2 r-lanes, every r-lane strides over r, and the partial sums are combined
// by a fixed-order tree in shared memory (deterministic).  beta rows are written by whichever
// lane visits them, exactly once.

template <typename SR, typename T>
__global__ void __launch_bounds__(kThreads) jt_project_splitr_kernel(const KArgs a) {
    __shared__ T red[kThreads];
    const int bx_log2 = a.bx_log2;
    const int tx = threadIdx.x & ((1 << bx_log2) - 1);
    const int rz = threadIdx.x >> bx_log2, RZ = kThreads >> bx_log2;
    const long long b = ((long long)blockIdx.y << bx_log2) + tx;
    const bool valid = b < a.B;
    int s;
    const DTask* tk = locate_chunk(a, s);             // prefix for chunks of one s: s0 == s
    const int n_s = tk->n_s;

    const int* __restrict__ tab = a.tab;
    const DMsg* __restrict__ msgs = a.msgs;
    const long long B = a.B, col = valid ? b : 0;
    T* work = static_cast<T*>(a.work);
    const T* uni = static_cast<const T*>(a.uni);
    const bool um = a.uniform != 0;
    const int tflags = um ? tk->flags : 0;

    const int n_slo = tk->n_slo;
    int s_hi = 0, s_lo = s;
    if (n_slo < n_s) {
        s_hi = s / n_slo;
        s_lo = s - s_hi * n_slo;
    }
    T sm = SR::template one<T>();
    for (int j = tk->smsg_begin; j < tk->smsg_end; ++j) {
        const DMsg* m = msgs + j;
        const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
        sm = SR::mul(sm, (um && m->uni) ? __ldg(uni + idx) : work[m->eoff + idx * B + col]);
    }
    T own = SR::template one<T>();
    if (tk->own >= 0)
        own = (tflags & JT_TF_OWN_UNIFORM) ? __ldg(uni + tk->own + s) : work[(tk->own + s) * B + col];
    const T scale = SR::mul(sm, own);

    const bool has_src = tk->src >= 0, src_uni = (tflags & JT_TF_SRC_UNIFORM) != 0;
    const bool wbeta = tk->beta >= 0 && !(a.flags & JT_NO_BELIEFS) &&
                       !((tflags & JT_TF_BETA_SPLIT) && (a.flags & JT_X_BETA_SPLIT));
    const long long s_off = __ldg(tab + tk->src_shi + s_hi) + __ldg(tab + tk->src_slo + s_lo);
    const int rm0 = tk->rmsg_begin, nr = tk->rmsg_end - rm0;
    const int n_r = tk->n_r, n_rlo = tk->n_rlo;

    typename SR::template Acc<T> acc = SR::template acc_zero<T>();
    for (int r = rz; r < n_r; r += RZ) {
        const int rh = r / n_rlo, rl = r - rh * n_rlo;
        const long long e = s_off + __ldg(tab + tk->src_rhi + rh) + __ldg(tab + tk->src_rlo + rl);
        T v = SR::template one<T>();
        if (has_src) v = src_uni ? __ldg(uni + tk->src + e) : work[(tk->src + e) * B + col];
        for (int j = 0; j < nr; ++j) {
            const DMsg* m = msgs + rm0 + j;
            const long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo) +
                                  __ldg(tab + m->b_hi + rh) + __ldg(tab + m->b_lo + rl);
            v = SR::mul(v, (um && m->uni) ? __ldg(uni + idx) : work[m->eoff + idx * B + col]);
        }
        SR::accum(acc, v);
        if (wbeta && valid) work[(tk->beta + e) * B + col] = SR::mul(v, scale);
    }

    red[threadIdx.x] = SR::finish(acc);
    __syncthreads();
    for (int stride = RZ >> 1; stride > 0; stride >>= 1) {
        if (rz < stride) red[threadIdx.x] = SR::add(red[threadIdx.x], red[threadIdx.x + (stride << bx_log2)]);
        __syncthreads();
    }
    if (rz == 0 && valid && tk->out >= 0) {
        T o = SR::mul(red[threadIdx.x], sm);
        T* obase = tk->out_space ? static_cast<T*>(a.fout) : work;
        obase[(tk->out + s) * B + col] = o;
        if (tk->bel >= 0 && (a.flags & JT_SEP_BELIEFS)) work[(tk->bel + s) * B + col] = SR::mul(o, own);
    }
}

// ------------------------------------------------------------------------------------------
// TMA-pipelined projection (same task semantics as jt_project_kernel).
//
// The LDG kernel above keeps every in-flight row in registers, so its memory-level parallelism
// is capped by occupancy (ncu, round 1: 22 % warps active, DRAM 33-53 %).  Here the CTA is
// specialised into three roles:
//   * row producer warp: lane k owns the k-th streamed operand of the task (clique row,
//     message rows, own up-message), walks the (s, r) index space and issues 1-D bulk async
//     copies (cp.async.bulk, SASS UBLKCP) of whole batch-tile rows into a shared-memory ring
//     guarded by full/empty mbarriers.  A stage holds the rows of one (s, r) item; operands
//     that depend on s only ride along with the r = 0 item.  Bytes in flight = ring size
//     (~96 KB per CTA, two CTAs per SM), independent of register pressure;
//   * uniform warp: operands that are identical for every instance (uniform mode) are scalars;
//     lane l resolves item 32 b + l of a double-buffered batch, so the L2 latency of the
//     scalar reads is paid once per 32 items;
//   * consumer warps: one thread per 16-byte batch vector multiplies the operands out of
//     shared memory, accumulates over r and stores beliefs and messages with coalesced
//     16-byte stores.

constexpr int kTmaSlots = 24;     // ring size in rows (one row = 16 bytes x consumer threads)
constexpr int kUBatch = 32;       // items resolved per batch by the uniform warp

// Shared-memory bookkeeping that follows the ring rows.
template <typename T>
struct TmaAux {
    unsigned long long full[kTmaSlots];
    unsigned long long empty[kTmaSlots];
    unsigned long long u_full[2];
    unsigned long long u_empty[2];
    int e_row[kTmaSlots];                       // clique row index of the item in each stage
    int u_e[2][kUBatch];                        // same, produced by the uniform warp
    T u_val[2][kUBatch][4];                     // products of the scalar operands: item, s-only, own
};

// VPT: 16-byte batch vectors per consumer thread.  The ring holds kTmaSlots / VPT rows of
// ct * VPT vectors, so the shared-memory footprint (and bytes in flight) is the same.
template <typename SR, typename T, int VPT>
__global__ void __launch_bounds__(kThreads + 64, 2) jt_project_tma_kernel(const KArgs a) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int kSlots = kTmaSlots / VPT;
    typedef Pack<T, VEC> P;
    typedef AccPack<SR, T, VEC> A;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int ct = blockDim.x - 64;                      // consumer threads
    const int tw = ct * VPT;                             // batch vectors per tile
    const int row_pitch = tw * 16;                       // bytes per ring row
    TmaAux<T>* aux = reinterpret_cast<TmaAux<T>*>(smem_raw + kSlots * row_pitch);
    const uint32_t slots_u32 = smem_u32(smem_raw);
    const uint32_t full_u32 = smem_u32(aux->full), empty_u32 = smem_u32(aux->empty);
    const uint32_t ufull_u32 = smem_u32(aux->u_full), uempty_u32 = smem_u32(aux->u_empty);

    // block -> (task, chunk of s): per-task chunk sizes (a.prefix: [n_tasks + 1] first block of
    // each task, then [n_tasks] log2 of the task's chunk) balance the (s, r) items per CTA
    int s0, chunk_log2;
    const DTask* tk;
    {
        const int bid = blockIdx.x;
        int lo = 0, hi = a.n_tasks;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(a.prefix + mid) <= bid) lo = mid; else hi = mid;
        }
        chunk_log2 = __ldg(a.prefix + a.n_tasks + 1 + lo);
        s0 = (bid - __ldg(a.prefix + lo)) << chunk_log2;
        tk = a.tasks + lo;
    }
    const int n_s = tk->n_s;
    const int s1 = min(n_s, s0 + (1 << chunk_log2));
    const long long col0v = (long long)blockIdx.y * tw;
    const int ncols = (int)min((long long)tw, a.Bv - col0v);
    const uint32_t row_bytes = (uint32_t)ncols * 16u;

    // operands in order: [src] [r-dependent messages] [s-only messages] [own]
    const bool um = a.uniform != 0;
    const int tflags = um ? tk->flags : 0;
    const int has_src = tk->src >= 0 ? 1 : 0, has_own = tk->own >= 0 ? 1 : 0;
    const int m0 = tk->rmsg_begin;
    const int nr = tk->rmsg_end - m0;
    const int nsm = tk->smsg_end - tk->smsg_begin;
    const int n_ops = has_src + nr + nsm + has_own;
    const int n_item_ops = has_src + nr;                 // needed for every (s, r); the rest with r = 0
    unsigned umask = 0;                                  // bit k: operand k is uniform (a scalar)
    if (has_src && (tflags & JT_TF_SRC_UNIFORM)) umask |= 1u;
    if (um)
        for (int j = 0; j < nr + nsm; ++j)
            if (a.msgs[m0 + j].uni) umask |= 1u << (has_src + j);
    if (has_own && (tflags & JT_TF_OWN_UNIFORM)) umask |= 1u << (n_ops - 1);
    const unsigned rowmask = ((1u << n_ops) - 1u) & ~umask;     // operands that stream ring rows
    const int n_rows = __popc(rowmask);
    // a stage groups G consecutive items so that one barrier round trip covers G * n_rows rows
    int G = n_rows > 0 ? kSlots / (4 * n_rows) : 1;
    G = G < 1 ? 1 : (G > 4 ? 4 : G);
    const int n_plane = n_rows * G;                      // producer lanes = ring rows per stage
    const int n_stage = n_rows > 0 ? kSlots / n_plane : 1;
    const int n_r = tk->n_r;
    const int n_items = (s1 - s0) * n_r;
    const long long B = a.B;
    const bool src_uni = (umask & 1u) && has_src;

    const int n_cwarps = ct >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stage; ++i) {
            mbar_init(full_u32 + 8 * i, n_plane > 0 ? n_plane : 1);  // every producer lane arrives per stage
            mbar_init(empty_u32 + 8 * i, n_cwarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(ufull_u32 + 8 * i, 1);
            mbar_init(uempty_u32 + 8 * i, n_cwarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == n_cwarps) {
        // ---------------- row producer warp ----------------
        // A stage holds G consecutive (s, r) items; lane g * n_rows + k streams the k-th row
        // operand of sub-item g, so one trip through the barriers moves up to G * n_rows rows
        // and the G address chains run in parallel lanes.
        if (lane >= n_plane) return;
        const int g = lane / n_rows, k = lane - g * n_rows;
        int op = 0;                                       // operand index of this lane
        for (int seen = -1; op < n_ops; ++op)
            if (((rowmask >> op) & 1u) && ++seen == k) break;
        const int jm = op - has_src;
        const bool is_src = has_src && op == 0;
        const bool is_own = has_own && op == n_ops - 1;
        const bool per_item = op < n_item_ops;            // else fetched with r = 0 only

        const int* __restrict__ tab = a.tab;
        const int n_slo = tk->n_slo, n_rlo = tk->n_rlo;
        long long base = 0, eoff = 0;
        const int* t_ahi = tab;
        const int* t_alo = tab;
        const int* t_bhi = tab;
        const int* t_blo = tab;
        if (is_src) {
            base = tk->src;
            t_ahi = tab + tk->src_shi; t_alo = tab + tk->src_slo;
            t_bhi = tab + tk->src_rhi; t_blo = tab + tk->src_rlo;
        } else if (!is_own) {
            const DMsg* m = a.msgs + m0 + jm;
            base = m->off;
            eoff = m->eoff;
            t_ahi = tab + m->a_hi; t_alo = tab + m->a_lo;
            t_bhi = tab + m->b_hi; t_blo = tab + m->b_lo;
        } else {
            base = tk->own;
        }
        const T* origin = static_cast<const T*>(a.work) + eoff + col0v * VEC;
        const uint32_t dst0 = slots_u32 + (uint32_t)lane * (uint32_t)row_pitch;
        const uint32_t stage_bytes = (uint32_t)n_plane * (uint32_t)row_pitch;

        int stage = 0;
        uint32_t phase = 0;
        for (int i = g; i - g < n_items; i += G) {        // item of this lane in the current stage
            const uint32_t full = full_u32 + 8 * stage;
            const bool live = i < n_items;
            int e = 0, s_idx = 0;
            bool fetch = false;
            if (live) {
                const int ds = i / n_r;
                const int r = i - ds * n_r;
                const int s = s0 + ds;
                fetch = per_item || r == 0;
                if (fetch) {
                    if (is_own) {
                        s_idx = s;
                    } else {
                        int s_hi = 0, s_lo = s;
                        if (n_slo < n_s) {
                            s_hi = s / n_slo;
                            s_lo = s - s_hi * n_slo;
                        }
                        s_idx = __ldg(t_ahi + s_hi) + __ldg(t_alo + s_lo);
                        if (per_item) {
                            const int rh = r / n_rlo, rl = r - rh * n_rlo;
                            e = __ldg(t_bhi + rh) + __ldg(t_blo + rl);
                        }
                    }
                }
            }
            mbar_wait(empty_u32 + 8 * stage, phase ^ 1);
            if (fetch) {
                if (is_src) aux->e_row[stage * G + g] = s_idx + e;
                mbar_expect_tx(full, row_bytes);
                bulk_g2s(dst0 + (uint32_t)stage * stage_bytes, origin + (base + s_idx + e) * B, row_bytes, full);
            }
            mbar_arrive(full);
            if (++stage == n_stage) {
                stage = 0;
                phase ^= 1;
            }
        }
        return;
    }

    if (warp == n_cwarps + 1) {
        // ---------------- uniform warp: scalar operands, 32 items per batch ----------------
        // Lane l resolves item 32 b + l: its (s, r), the clique row index and the products of
        // the uniform operands, read from the uniform workspace: [0] per-item operands (src,
        // r-dependent messages), [1] s-only messages, [2] own.  One L2 round trip per batch.
        if (umask == 0) return;
        const int* __restrict__ tab = a.tab;
        const T* __restrict__ uni = static_cast<const T*>(a.uni);
        const int n_slo = tk->n_slo, n_rlo = tk->n_rlo;
        for (int b = 0; b * kUBatch < n_items; ++b) {
            const int buf = b & 1;
            mbar_wait(uempty_u32 + 8 * buf, ((b >> 1) & 1) ^ 1);
            const int i = b * kUBatch + lane;
            if (i < n_items) {
                const int ds = i / n_r;
                const int r = i - ds * n_r;
                const int s = s0 + ds;
                int s_hi = 0, s_lo = s;
                if (n_slo < n_s) {
                    s_hi = s / n_slo;
                    s_lo = s - s_hi * n_slo;
                }
                const int rh = r / n_rlo, rl = r - rh * n_rlo;
                const T one = SR::template one<T>();
                T item_prod = one, s_prod = one, own_val = one;
                if (src_uni) {
                    const int e = __ldg(tab + tk->src_shi + s_hi) + __ldg(tab + tk->src_slo + s_lo) +
                                  __ldg(tab + tk->src_rhi + rh) + __ldg(tab + tk->src_rlo + rl);
                    aux->u_e[buf][lane] = e;
                    item_prod = __ldg(uni + tk->src + e);
                }
                for (int j = 0; j < nr + nsm; ++j) {
                    if ((umask >> (has_src + j)) & 1u) {
                        const DMsg* m = a.msgs + m0 + j;
                        long long idx = m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo);
                        if (j < nr) {
                            idx += __ldg(tab + m->b_hi + rh) + __ldg(tab + m->b_lo + rl);
                            item_prod = SR::mul(item_prod, __ldg(uni + idx));
                        } else {
                            s_prod = SR::mul(s_prod, __ldg(uni + idx));
                        }
                    }
                }
                if (has_own && ((umask >> (n_ops - 1)) & 1u)) own_val = __ldg(uni + tk->own + s);
                T* out = aux->u_val[buf][lane];
                out[0] = item_prod;
                out[1] = s_prod;
                out[2] = own_val;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ufull_u32 + 8 * buf);
        }
        return;
    }

    // ---------------- consumers: thread t owns batch vectors col0v + t + v * ct, v < VPT ----------------
    const int t = threadIdx.x;
    bool active[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) active[v] = t + v * ct < ncols;
    T* work = static_cast<T*>(a.work);
    const long long col = (col0v + t) * VEC;
    const long long vstep = (long long)ct * VEC;         // elements between the vectors of a thread
    const bool wbeta = tk->beta >= 0 && !(a.flags & JT_NO_BELIEFS) &&
                       !((tflags & JT_TF_BETA_SPLIT) && (a.flags & JT_X_BETA_SPLIT));
    const bool wout = tk->out >= 0;
    const bool wbel = wout && tk->bel >= 0 && (a.flags & JT_SEP_BELIEFS);
    T* bptr = work + (wbeta ? tk->beta : 0) * B + col;
    T* optr = (tk->out_space ? static_cast<T*>(a.fout) : work) + (wout ? tk->out : 0) * B + col;
    T* lptr = work + (wbel ? tk->bel : 0) * B + col;
    const unsigned char* my = smem_raw + t * 16;
    const int vpitch = ct * 16;                          // bytes between the vectors of a thread
    const int stage_pitch = n_plane * row_pitch;
    const bool any_uni = umask != 0, any_row = n_rows > 0;
    // ring rows of a stage: the per-item operands first, then the per-s ones (own last)
    const int n_item_rows = __popc(rowmask & ((1u << n_item_ops) - 1u));
    const bool own_is_row = has_own && ((rowmask >> (n_ops - 1)) & 1u);
    const int n_sm_rows = n_rows - n_item_rows - (own_is_row ? 1 : 0);

    struct PV {
        P v[VPT];
    };
    struct AV {
        A v[VPT];
    };
    const T one = SR::template one<T>();

    // The per-item code is instantiated per number of streamed per-item operands and per
    // "writes beliefs", and the (s, r) loops are kept nested, so that an item costs a few dozen
    // instructions (ncu, round 1: a generic flattened loop made the consumers issue-bound at
    // ~100-160 instructions per item); with VPT = 2 a thread covers 32 bytes of every row, which
    // halves the control instructions per byte again.
    auto consume = [&](auto ni_tag, auto wbeta_tag) {
        constexpr int NI = decltype(ni_tag)::value;          // -1: run-time count
        constexpr bool WB = decltype(wbeta_tag)::value;
        const int ni = NI >= 0 ? NI : n_item_rows;
        const int sub_pitch = n_rows * row_pitch;            // rows of one sub-item
        int stage = 0, g = 0, ub = 0, ul = 0;
        uint32_t phase = 0, uphase = 0;
        const unsigned char* sub = my;
        const T* uv = aux->u_val[0][0];

        auto begin_item = [&]() {
            if (any_row && g == 0) mbar_wait(full_u32 + 8 * stage, phase);
            if (any_uni && ul == 0) mbar_wait(ufull_u32 + 8 * ub, uphase);
            sub = my + stage * stage_pitch + g * sub_pitch;
            uv = aux->u_val[ub][ul];
        };
        auto release_stage = [&]() {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_u32 + 8 * stage);
            g = 0;
            if (++stage == n_stage) {
                stage = 0;
                phase ^= 1;
            }
        };
        auto release_batch = [&]() {
            __syncwarp();
            if (lane == 0) mbar_arrive(uempty_u32 + 8 * ub);
            ul = 0;
            ub ^= 1;
            if (ub == 0) uphase ^= 1;
        };
        auto end_item = [&]() {
            if (any_uni && ++ul == kUBatch) release_batch();
            if (any_row && ++g == G) release_stage();
        };
        auto load_row = [&](const unsigned char* row, PV& x) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) x.v[v] = *reinterpret_cast<const P*>(row + v * vpitch);
        };
        auto item_value = [&]() {
            PV val;
            const P u = pack_fill<T, VEC>(any_uni ? uv[0] : one);
#pragma unroll
            for (int v = 0; v < VPT; ++v) val.v[v] = u;
            const int n = NI >= 0 ? NI : ni;
#pragma unroll
            for (int k = 0; k < (NI >= 0 ? NI : 8); ++k) {
                if (k < n) {
                    PV x;
                    load_row(sub + k * row_pitch, x);
#pragma unroll
                    for (int v = 0; v < VPT; ++v) mul<SR>(val.v[v], x.v[v]);
                }
            }
            return val;
        };
        auto store_beta = [&](const PV& val, const PV& scale) {
            if (WB) {
                const int e = src_uni ? aux->u_e[ub][ul] : aux->e_row[stage * G + g];
                T* dst = bptr + (long long)e * B;
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    if (active[v]) {
                        P x = val.v[v];
                        mul<SR>(x, scale.v[v]);
                        st<T, VEC>(dst + v * vstep, x);
                    }
                }
            }
        };

        for (int s = s0; s < s1; ++s) {
            // r = 0: the item that also carries the s-only operands and own
            begin_item();
            const PV first = item_value();
            PV sm, own, scale;
            AV acc0, acc1;
            const unsigned char* srow = sub + ni * row_pitch;
            const P u1 = pack_fill<T, VEC>(any_uni ? uv[1] : one), u2 = pack_fill<T, VEC>(any_uni ? uv[2] : one);
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
                acc0.v[v] = acc_zero<SR, T, VEC>();
                acc1.v[v] = acc_zero<SR, T, VEC>();
                accum<SR>(acc0.v[v], first.v[v]);
                sm.v[v] = u1;
                own.v[v] = u2;
            }
            for (int k = 0; k < n_sm_rows; ++k) {
                PV x;
                load_row(srow + k * row_pitch, x);
#pragma unroll
                for (int v = 0; v < VPT; ++v) mul<SR>(sm.v[v], x.v[v]);
            }
            if (own_is_row) load_row(srow + n_sm_rows * row_pitch, own);
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
                scale.v[v] = sm.v[v];
                mul<SR>(scale.v[v], own.v[v]);
            }
            store_beta(first, scale);
            end_item();
            for (int r = 1; r < n_r; ++r) {
                begin_item();
                const PV val = item_value();
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    if (r & 1) accum<SR>(acc1.v[v], val.v[v]); else accum<SR>(acc0.v[v], val.v[v]);
                }
                store_beta(val, scale);
                end_item();
            }
            if (wout) {
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    if (active[v]) {
                        merge<SR>(acc0.v[v], acc1.v[v]);
                        P o = finish<SR>(acc0.v[v]);
                        mul<SR>(o, sm.v[v]);
                        st<T, VEC>(optr + (long long)s * B + v * vstep, o);
                        if (wbel) {
                            mul<SR>(o, own.v[v]);
                            st<T, VEC>(lptr + (long long)s * B + v * vstep, o);
                        }
                    }
                }
            }
        }
        if (any_uni && ul != 0) release_batch();              // partially used last batch / stage
        if (any_row && g != 0) release_stage();
    };
    auto dispatch_ni = [&](auto wbeta_tag) {
        switch (n_item_rows) {
            case 0: consume(std::integral_constant<int, 0>{}, wbeta_tag); break;
            case 1: consume(std::integral_constant<int, 1>{}, wbeta_tag); break;
            case 2: consume(std::integral_constant<int, 2>{}, wbeta_tag); break;
            case 3: consume(std::integral_constant<int, 3>{}, wbeta_tag); break;
            default: consume(std::integral_constant<int, -1>{}, wbeta_tag); break;
        }
    };
    if (wbeta) dispatch_ni(std::true_type{}); else dispatch_ni(std::false_type{});
}

// ------------------------------------------------------------------------------------------
// Whole-propagation kernel for tiny workloads (a handful of instances of a small tree: the
// reference's own use, one `tree.propagate(values)` call on a network like its README example).
// There the level-ordered launches are pure launch latency (config 1: ~10 launches of a few
// microseconds each).  One CTA owns one instance and walks the whole schedule -- evidence
// offsets, init, collect, distribute, marginal -- with a block barrier between the launches of
// the plan; the task semantics are those of jt_project_kernel (general mode: every operand is
// read per instance, the uniform flags are ignored).  Workspace values written earlier in the
// same kernel are read with plain loads (never through the read-only path).

struct WalkArgs {
    const int* seq;        // [n_seq][2] task ranges, in execution order
    int n_seq;
    const int* evidence;   // [B][n_evid] or null
    int n_evid;
    const int* ev_card;
    const int* evf_ptr;
    const int* evf_var;
    const int* evf_stride;
    int n_factors;
    unsigned long long* errors;
    long long lik_base, lik_entries;   // likelihood region of the workspace (filled by the caller)
    long long work_entries;            // entries of the workspace column (shared-memory mirror)
    int n_tasks, n_msgs, n_tab;        // sizes of the plan's descriptor arrays (staged in shared memory)
    long long preload;                 // leading entries of the column that hold inputs written by the caller
};

// The instance's column of the workspace.  SM: mirrored in shared memory (small trees: every
// dependent read costs a shared-memory access instead of an L2 round trip); stores go through
// to the global workspace either way, so it ends up as the per-level kernels leave it.
template <typename T, bool SM>
struct WalkMem {
    T* work;
    T* sw;
    long long B, b;
    __device__ __forceinline__ T get(long long idx) const { return SM ? sw[idx] : work[idx * B + b]; }
    __device__ __forceinline__ void put(long long idx, T v) const {
        if (SM) sw[idx] = v;
        work[idx * B + b] = v;
    }
};

template <typename SR, typename T, bool SM>
__device__ __forceinline__ void walk_init(const KArgs& a, const DTask* tk, const WalkMem<T, SM>& mem, int first,
                                          int stride) {
    const int* tab = a.tab;             // global or shared: plain loads
    const DMsg* msgs = a.msgs;
    const long long B = a.B, b = mem.b;
    const T* __restrict__ fin = static_cast<const T*>(a.fin);
    const int n_s = tk->n_s, n_slo = tk->n_slo;
    for (int s = first; s < n_s; s += stride) {
        const int s_hi = s / n_slo, s_lo = s - s_hi * n_slo;
        T val = SR::template one<T>();
        for (int j = tk->smsg_begin; j < tk->smsg_end; ++j) {
            const DMsg* m = msgs + j;
            const long long idx = m->off + tab[m->a_hi + s_hi] + tab[m->a_lo + s_lo];
            if (m->fid < 0) val = SR::mul(val, mem.get(idx));
            else if (a.fin_batched) val = SR::mul(val, __ldg(fin + idx * B + b));
            else val = SR::mul(val, __ldg(fin + idx + (a.fbase ? a.fbase[(long long)m->fid * B + b] : 0)));
        }
        mem.put(tk->out + s, val);
    }
}

template <typename SR, typename T, bool SM>
__device__ __forceinline__ void walk_project(const KArgs& a, const DTask* tk, const WalkMem<T, SM>& mem, int first,
                                             int stride) {
    const int* tab = a.tab;             // global or shared: plain loads
    const DMsg* msgs = a.msgs;
    const int n_s = tk->n_s, n_slo = tk->n_slo, n_rlo = tk->n_rlo, n_rhi = tk->n_r / tk->n_rlo;
    const bool has_src = tk->src >= 0;
    const bool wbeta = tk->beta >= 0 && !(a.flags & JT_NO_BELIEFS);
    const int rm0 = tk->rmsg_begin, nr = tk->rmsg_end - rm0;
    for (int s = first; s < n_s; s += stride) {
        const int s_hi = s / n_slo, s_lo = s - s_hi * n_slo;
        T sm = SR::template one<T>();
        for (int j = tk->smsg_begin; j < tk->smsg_end; ++j) {
            const DMsg* m = msgs + j;
            sm = SR::mul(sm, mem.get(m->off + tab[m->a_hi + s_hi] + tab[m->a_lo + s_lo]));
        }
        T own = SR::template one<T>();
        if (tk->own >= 0) own = mem.get(tk->own + s);
        const T scale = SR::mul(sm, own);
        const long long s_off = tab[tk->src_shi + s_hi] + tab[tk->src_slo + s_lo];
        typename SR::template Acc<T> acc = SR::template acc_zero<T>();
        for (int rh = 0; rh < n_rhi; ++rh) {
            for (int rl = 0; rl < n_rlo; ++rl) {
                const long long e = s_off + tab[tk->src_rhi + rh] + tab[tk->src_rlo + rl];
                T v = has_src ? mem.get(tk->src + e) : SR::template one<T>();
                for (int j = 0; j < nr; ++j) {
                    const DMsg* m = msgs + rm0 + j;
                    v = SR::mul(v, mem.get(m->off + tab[m->a_hi + s_hi] + tab[m->a_lo + s_lo] +
                                           tab[m->b_hi + rh] + tab[m->b_lo + rl]));
                }
                SR::accum(acc, v);
                if (wbeta) mem.put(tk->beta + e, SR::mul(v, scale));
            }
        }
        if (tk->out >= 0) {
            const T o = SR::mul(SR::finish(acc), sm);
            if (tk->out_space) static_cast<T*>(a.fout)[(tk->out + s) * mem.B + mem.b] = o;
            else mem.put(tk->out + s, o);
            if (tk->bel >= 0 && (a.flags & JT_SEP_BELIEFS)) mem.put(tk->bel + s, SR::mul(o, own));
        }
    }
}

template <typename SR, typename T, bool SM>
__global__ void __launch_bounds__(kThreads) jt_walk_kernel(const KArgs a, const WalkArgs w) {
    extern __shared__ __align__(16) unsigned char walk_smem[];
    const long long b = blockIdx.x, B = a.B;
    const WalkMem<T, SM> mem = {static_cast<T*>(a.work), reinterpret_cast<T*>(walk_smem), B, b};
    KArgs la = a;
    const int* seq = w.seq;
    if (SM) {
        // the schedule itself (task and message descriptors, index tables, launch ranges) is
        // staged too: every task is visited once, so from global memory each level would pay a
        // chain of first-touch L2 misses (descriptor -> message -> table -> value)
        unsigned char* p = walk_smem + ((w.work_entries * sizeof(T) + 15) / 16) * 16;
        DTask* s_tasks = reinterpret_cast<DTask*>(p);
        DMsg* s_msgs = reinterpret_cast<DMsg*>(s_tasks + w.n_tasks);
        int* s_tab = reinterpret_cast<int*>(s_msgs + w.n_msgs);
        int* s_seq = s_tab + w.n_tab;
        const int* g_tasks = reinterpret_cast<const int*>(a.tasks);
        for (int i = threadIdx.x; i < w.n_tasks * (int)(sizeof(DTask) / 4); i += blockDim.x)
            reinterpret_cast<int*>(s_tasks)[i] = __ldg(g_tasks + i);
        const int* g_msgs = reinterpret_cast<const int*>(a.msgs);
        for (int i = threadIdx.x; i < w.n_msgs * (int)(sizeof(DMsg) / 4); i += blockDim.x)
            reinterpret_cast<int*>(s_msgs)[i] = __ldg(g_msgs + i);
        for (int i = threadIdx.x; i < w.n_tab; i += blockDim.x) s_tab[i] = __ldg(a.tab + i);
        for (int i = threadIdx.x; i < 2 * w.n_seq; i += blockDim.x) s_seq[i] = __ldg(w.seq + i);
        la.tasks = s_tasks;
        la.msgs = s_msgs;
        la.tab = s_tab;
        seq = s_seq;
        for (long long i = threadIdx.x; i < w.lik_entries; i += blockDim.x)   // soft evidence written by the caller
            mem.sw[w.lik_base + i] = mem.work[(w.lik_base + i) * B + b];
        for (long long i = threadIdx.x; i < w.preload; i += blockDim.x)       // potentials given by the caller
            mem.sw[i] = mem.work[i * B + b];
    }
    if (w.evidence) {     // V1: evidence slicing of this instance (jt_evidence_kernel)
        unsigned bad = 0;
        for (int f = threadIdx.x; f < w.n_factors; f += blockDim.x) {
            int acc = 0;
            for (int k = w.evf_ptr[f]; k < w.evf_ptr[f + 1]; ++k) {
                const int var = w.evf_var[k];
                int state = w.evidence[b * w.n_evid + var];
                const int card = w.ev_card[var];
                if (state < 0 || state >= card) {
                    ++bad;
                    state = state < 0 ? 0 : card - 1;
                }
                acc += state * w.evf_stride[k];
            }
            const_cast<int*>(a.fbase)[(long long)f * B + b] = acc;
        }
        if (bad) atomicAdd(w.errors, (unsigned long long)bad);
    }
    __syncthreads();
    // the tasks of one range are independent: small ones go to one warp each (task-level
    // parallelism; a tiny tree has a few outputs per task and a block-wide loop would leave one
    // warp walking the dependent descriptor -> table -> value loads of every task in turn)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int q = 0; q < w.n_seq; ++q) {
        int n_small = 0;
        for (int t = seq[2 * q]; t < seq[2 * q + 1]; ++t) {
            const DTask* tk = la.tasks + t;
            int first = threadIdx.x, stride = blockDim.x;
            if (tk->n_s <= 64) {
                if (n_small++ % n_warps != warp) continue;
                first = lane;
                stride = 32;
            }
            if (tk->kind == JT_KIND_INIT) walk_init<SR, T, SM>(la, tk, mem, first, stride);
            else walk_project<SR, T, SM>(la, tk, mem, first, stride);
        }
        __syncthreads();
    }
}

// Output stage: normalise every output scope per instance; log Z from scope 0.
template <typename SR, typename T>
__global__ void __launch_bounds__(kThreads)
jt_normalize_kernel(T* __restrict__ fout, const long long* __restrict__ out_off,
                    const long long* __restrict__ out_size, long long B, T* __restrict__ logz, int normalize) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int k = blockIdx.y;
    T* col = fout + out_off[k] * B + b;
    const long long n = out_size[k];
    typename SR::template Acc<T> acc = SR::template acc_zero<T>();
    for (long long e = 0; e < n; ++e) SR::accum(acc, col[e * B]);
    const T z = SR::finish(acc);
    if (normalize)
        for (long long e = 0; e < n; ++e) col[e * B] = SR::unit(col[e * B], z);
    if (k == 0 && logz) logz[b] = SR::log_of(z);
}

static_assert(kUnroll == 4, "the pairwise combination above assumes four partial sums");

}  // namespace
