// Dense contraction path of libjt_b200 (uniform mode, sum-product).
//
// The contraction is the reference's E2 / E4 einsum (junctiontree/computation.py:84-88, 205-207):
// a clique potential times the incoming messages, summed down to a separator.  With factor
// tables shared by the batch (uniform mode) the potential psi_C and most messages are the same
// for every instance; a projection task whose only per-instance input is ONE message M is then
//
//     out[s][b] = sum_r U(s, r) * M[A(s) + B(r)][b]          U = psi_C x uniform messages (scalars)
//
// The projection kernels stream one [B] row of M per (s, r) item, i.e. every row of M is re-read
// from L2 once per output row that uses it (config 4: 64-128 times, config 5: ~5 times on average).
// Grouping the output rows by the message rows they touch turns the task into a batch of small
// dense products
//
//     out[s_of[g][i]][b] = sum_k W[g][i][k] * M[mg[g] + mk[k]][b]        W = U summed over the
//                                                                        axes M does not see
//
// g = a value of the output axes M depends on, i = the other output axes, k = the summed axes M
// depends on: per group a [n_i x K] matrix of shared scalars times K message rows.  Every row of
// M and of out now moves once per batch tile; the arithmetic (2 n_i K flops per 8 (n_i + K)
// bytes) goes to the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
// Nothing of this is in the plan blob: the groups are recovered from the task's index tables
// when the plan is loaded (distinct values of A(s) and B(r)), so both plan emitters and the ABI
// are unchanged.  W lives in the *W region* of the workspace (after the uniform workspace) in
// MMA-fragment order and is rebuilt whenever the uniform workspace is (jt_dense_prepare).

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

#include "jt_kernels.cuh"

namespace {

constexpr int kDWarps = 8;                         // MMA warps per CTA
constexpr int kDNT = 2;                            // 8-column n-tiles per warp: 16 batch columns each
constexpr int kDTB = kDWarps * kDNT * 8;           // batch columns per CTA
constexpr int kDKC = 16;                           // message rows per pipeline stage
constexpr int kDStages = 4;
constexpr int kDRowPitch = kDTB * 8 + 32;          // bytes; +32: the 4 rows of a B fragment hit distinct banks
constexpr int kDWBytes = (kDKC / 4) * 4 * 32 * 8;  // W fragments of a stage: 4 k-steps x <= 4 m-tiles x 256 B
constexpr int kDStageBytes = kDKC * kDRowPitch + kDWBytes;
constexpr int kDSmem = kDStages * kDStageBytes + 2 * kDStages * 8;

struct DenseArgs {
    const DDense* dd;      // descriptors of this launch
    int n;
    const int* prefix;     // [n + 1] first block of each task, [n] units per CTA
    const int* dtab;
    const DTask* tasks;    // all tasks of the plan
    const DMsg* msgs;
    const int* tab;
    void* work;
    const void* uni;
    const void* W;
    void* fout;
    long long B;
    int flags;
};

struct PrepArgs {
    const DDense* dd;      // all descriptors of the plan
    const int* list;       // [n] descriptor ids, [n + 1] block prefix
    int n;
    const int* dtab;
    const DTask* tasks;
    const DMsg* msgs;
    const int* tab;
    const void* uni;
    void* W;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// W blocks from the uniform workspace: one thread per element of the padded fragment layout.
//   element (unit = g * n_it + it, k-step k4, m-tile mt, lane) holds W[g][i][k] with
//   i = it * 8 MT + 8 mt + lane / 4, k = 4 k4 + lane % 4 (zero outside n_i x K):
//   exactly what thread `lane` feeds to mma.m8n8k4 as the A operand.
template <typename T>
__global__ void __launch_bounds__(kThreads) jt_dense_prep_kernel(const PrepArgs a) {
    const int bid = blockIdx.x;
    const int* prefix = a.list + a.n;
    int lo = 0, hi = a.n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= bid) lo = mid; else hi = mid;
    }
    const DDense d = a.dd[__ldg(a.list + lo)];
    const long long idx = (long long)(bid - __ldg(prefix + lo)) * kThreads + threadIdx.x;
    if (idx >= d.w_size) return;
    const int lane = (int)(idx & 31);
    long long rest = idx >> 5;
    const int mt = (int)(rest % d.MT);
    rest /= d.MT;
    const int k4 = (int)(rest % d.n_k4);
    const long long unit = rest / d.n_k4;
    const int g = (int)(unit / d.n_it), it = (int)(unit % d.n_it);
    const int i = (it * d.MT + mt) * 8 + (lane >> 2);
    const int k = k4 * 4 + (lane & 3);
    double* W = static_cast<double*>(a.W) + d.w_off;       // always float64: it feeds the FP64 tensor pipe
    if (i >= d.n_i || k >= d.K) {
        W[idx] = 0.0;
        return;
    }
    const int* __restrict__ tab = a.tab;
    const int* __restrict__ dtab = a.dtab;
    const T* __restrict__ uni = static_cast<const T*>(a.uni);
    const DTask* tk = a.tasks + d.task;
    const int s = __ldg(dtab + d.s_of + g * d.n_i + i);
    const int n_slo = tk->n_slo, n_rlo = tk->n_rlo;
    const int s_hi = s / n_slo, s_lo = s - s_hi * n_slo;
    // uniform messages that depend on s only
    double scale = 1.0;
    for (int j = tk->smsg_begin; j < tk->smsg_end; ++j) {
        const DMsg* m = a.msgs + j;
        if (m->uni) scale *= (double)__ldg(uni + m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo));
    }
    const long long s_off = tk->src + __ldg(tab + tk->src_shi + s_hi) + __ldg(tab + tk->src_slo + s_lo);
    double sum = 0.0;
    for (int q = 0; q < d.n_q; ++q) {                  // r ascending: the order of the projection kernels
        const int r = __ldg(dtab + d.r_of + k * d.n_q + q);
        const int rh = r / n_rlo, rl = r - rh * n_rlo;
        double u = (double)__ldg(uni + s_off + __ldg(tab + tk->src_rhi + rh) + __ldg(tab + tk->src_rlo + rl));
        for (int j = tk->rmsg_begin; j < tk->rmsg_end; ++j) {
            const DMsg* m = a.msgs + j;
            if (m->uni)
                u *= (double)__ldg(uni + m->off + __ldg(tab + m->a_hi + s_hi) + __ldg(tab + m->a_lo + s_lo) +
                                   __ldg(tab + m->b_hi + rh) + __ldg(tab + m->b_lo + rl));
        }
        sum += u;
    }
    W[idx] = sum * scale;
}

// One MMA warp of jt_dense_kernel for a compile-time number MT of 8-row m-tiles per unit: warp w
// owns columns col0 + 16 w .. + 15 of the batch tile and keeps the [8 MT x 16] accumulator tile of
// the current unit in registers.  Everything per unit is unrolled over MT (short contractions run
// hundreds of units per CTA with 4-8 DMMAs each: the instructions around the DMMAs are what the
// kernel is bound by there, so no predicated m-tiles, 32-bit unit counters advanced without
// divisions, row pointers formed once per output row).
template <typename T, int MT>
__device__ __forceinline__ void dense_mma_warp(const DenseArgs& a, const DDense& d, const unsigned char* smem,
                                               const uint32_t full_u32, const uint32_t empty_u32, const int u0,
                                               const int u1, const long long col0, const int warp, const int lane) {
    typedef Pack<T, 2> P2;
    const long long B = a.B;
    const int* __restrict__ dtab = a.dtab;
    const DTask* tk = a.tasks + d.task;
    T* work = static_cast<T*>(a.work);
    const T* uni = static_cast<const T*>(a.uni);
    // the task's fields in registers: re-read after every store otherwise (they could alias the workspace)
    const long long t_own = tk->own, t_bel = tk->bel;
    const int t_n_slo = tk->n_slo, smsg_begin = tk->smsg_begin, smsg_end = tk->smsg_end;
    const bool wbel = t_bel >= 0 && (a.flags & JT_SEP_BELIEFS);
    const bool own_uni = wbel && t_own >= 0 && (tk->flags & JT_TF_OWN_UNIFORM);
    const bool own_rows = wbel && t_own >= 0 && !(tk->flags & JT_TF_OWN_UNIFORM);
    // does the epilogue have per-instance s-only operands?
    bool s_rows = false;
    for (int j = smsg_begin; j < smsg_end; ++j) s_rows = s_rows || !a.msgs[j].uni;
    const int bcol = warp * (kDNT * 8) + (lane >> 2);                      // B fragment: column inside the tile
    const long long ccol = col0 + warp * (kDNT * 8) + (lane & 3) * 2;      // C fragment: first of two columns, + 8 nt
    bool ok[kDNT];                                                         // B is even: both columns or none
#pragma unroll
    for (int nt = 0; nt < kDNT; ++nt) ok[nt] = ccol + nt * 8 < B;
    T* const obase = (tk->out_space ? static_cast<T*>(a.fout) : work) + tk->out * B + ccol;
    T* const wcol = work + ccol;
    const int n_i = d.n_i, n_it = d.n_it;
    const int* const s_tab = dtab + d.s_of + (lane >> 2);

    struct Acc {
        double v[MT][kDNT][2];
    };
    struct Rows {
        int s[MT];
    };
    // nk4 k-steps: rows `rows` (4 per step) times W fragments `wt` (MT per step)
    auto steps = [&](Acc& c, const unsigned char* rows, const double* wt, const int nk4) {
#pragma unroll
        for (int q = 0; q < kDKC / 4; ++q) {
            if (q >= nk4) break;
            double bf[kDNT];
#pragma unroll
            for (int nt = 0; nt < kDNT; ++nt)
                bf[nt] = (double)*reinterpret_cast<const T*>(rows + q * 4 * kDRowPitch + nt * 8 * (int)sizeof(T));
            double af[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) af[mt] = wt[(q * MT + mt) * 32];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < kDNT; ++nt) dmma(c.v[mt][nt][0], c.v[mt][nt][1], af[mt], bf[nt]);
        }
    };
    // output rows of unit (g, it) handled by this thread: i = (it MT + mt) 8 + lane / 4 (-1: padding),
    // fetched one unit ahead so the lookups are off the critical path of the epilogue
    auto lookup = [&](const int g, const int it) {
        Rows r;
        const int* t = s_tab + g * n_i + it * (8 * MT);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
            r.s[mt] = it * (8 * MT) + mt * 8 + (lane >> 2) < n_i ? __ldg(t + mt * 8) : -1;
        return r;
    };
    // Two m-tiles at a time: first every load of the pair (own rows, per-instance s-only rows), then
    // the products and stores -- loads and stores share the workspace pointer, so only loads issued
    // back to back overlap their latencies.
    auto epilogue = [&](const Acc& c, const Rows& ur) {
        if (!wbel && !s_rows) {                                  // a message and nothing else
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int s = ur.s[mt];
                if (s < 0) continue;
                T* o = obase + (long long)s * B;
#pragma unroll
                for (int nt = 0; nt < kDNT; ++nt) {
                    if (!ok[nt]) continue;
                    P2 x;
                    x.v[0] = (T)c.v[mt][nt][0];
                    x.v[1] = (T)c.v[mt][nt][1];
                    *reinterpret_cast<P2*>(o + nt * 8) = x;
                }
            }
            return;
        }
#pragma unroll
        for (int m0 = 0; m0 < MT; m0 += 2) {
            P2 ow[2][kDNT];
            double own_u[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (m0 + h >= MT) continue;
                const int s = ur.s[m0 + h];
                own_u[h] = 1.0;
#pragma unroll
                for (int nt = 0; nt < kDNT; ++nt) ow[h][nt].v[0] = ow[h][nt].v[1] = T(1);
                if (s < 0) continue;
                if (own_uni) {
                    own_u[h] = (double)__ldg(uni + t_own + s);
                } else if (own_rows) {
                    const T* r = wcol + (t_own + s) * B;
#pragma unroll
                    for (int nt = 0; nt < kDNT; ++nt)
                        if (ok[nt]) ow[h][nt] = *reinterpret_cast<const P2*>(r + nt * 8);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (m0 + h >= MT) continue;
                const int mt = m0 + h, s = ur.s[mt];
                if (s < 0) continue;
                double v[kDNT][2];
#pragma unroll
                for (int nt = 0; nt < kDNT; ++nt) {
                    v[nt][0] = c.v[mt][nt][0];
                    v[nt][1] = c.v[mt][nt][1];
                }
                if (s_rows) {
                    const int s_hi = s / t_n_slo, s_lo = s - s_hi * t_n_slo;
                    for (int j = smsg_begin; j < smsg_end; ++j) {
                        const DMsg* m = a.msgs + j;
                        if (m->uni) continue;                    // folded into W
                        const T* r = wcol + m->eoff + (m->off + __ldg(a.tab + m->a_hi + s_hi) + __ldg(a.tab + m->a_lo + s_lo)) * B;
#pragma unroll
                        for (int nt = 0; nt < kDNT; ++nt) {
                            if (!ok[nt]) continue;
                            const P2 x = *reinterpret_cast<const P2*>(r + nt * 8);
                            v[nt][0] *= (double)x.v[0];
                            v[nt][1] *= (double)x.v[1];
                        }
                    }
                }
                T* o = obase + (long long)s * B;
                T* bel = wcol + (t_bel + s) * B;
#pragma unroll
                for (int nt = 0; nt < kDNT; ++nt) {
                    if (!ok[nt]) continue;
                    P2 x;
                    x.v[0] = (T)v[nt][0];
                    x.v[1] = (T)v[nt][1];
                    *reinterpret_cast<P2*>(o + nt * 8) = x;
                    if (wbel) {
                        // from the stored (rounded) message, as the projection kernels form it
                        x.v[0] = (T)((double)x.v[0] * (double)ow[h][nt].v[0] * own_u[h]);
                        x.v[1] = (T)((double)x.v[1] * (double)ow[h][nt].v[1] * own_u[h]);
                        *reinterpret_cast<P2*>(bel + nt * 8) = x;
                    }
                }
            }
        }
    };

    int stage = 0;
    uint32_t phase = 0;
    auto release = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_u32 + 8 * stage);
        if (++stage == kDStages) {
            stage = 0;
            phase ^= 1;
        }
    };
    const int frag_off = (lane & 3) * kDRowPitch + bcol * (int)sizeof(T);   // B fragment of this thread inside a stage
    // (gn, itn): the unit after the one whose rows were looked up last
    int gn = u0 / n_it, itn = u0 - gn * n_it;
    Rows next = lookup(gn, itn);
    auto advance = [&]() {
        if (++itn == n_it) {
            itn = 0;
            ++gn;
        }
    };
    advance();
    if (d.ups > 1) {
        // short contractions: `ups` units per stage, one after the other
        const int kpad = d.n_k4 * 4, nk4 = d.n_k4, ups = d.ups;
        for (int u = u0; u < u1; u += ups) {
            const int nu = u1 - u < ups ? u1 - u : ups;
            mbar_wait(full_u32 + 8 * stage, phase);
            const unsigned char* st = smem + stage * kDStageBytes;
            const unsigned char* rows = st + frag_off;
            const double* wt = reinterpret_cast<const double*>(st + kDKC * kDRowPitch) + lane;
            for (int j = 0; j < nu; ++j) {
                Acc c;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < kDNT; ++nt) c.v[mt][nt][0] = c.v[mt][nt][1] = 0.0;
                const Rows ur = next;
                if (u + j + 1 < u1) {
                    next = lookup(gn, itn);
                    advance();
                }
                steps(c, rows, wt, nk4);
                rows += kpad * kDRowPitch;
                wt += nk4 * MT * 32;
                if (j == nu - 1) release();                       // the stage is consumed: refill during the epilogue
                epilogue(c, ur);
            }
        }
        return;
    }
    for (int u = u0; u < u1; ++u) {
        Acc c;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < kDNT; ++nt) c.v[mt][nt][0] = c.v[mt][nt][1] = 0.0;
        const Rows ur = next;
        if (u + 1 < u1) {
            next = lookup(gn, itn);
            advance();
        }
        for (int ch = 0; ch < d.n_chunks; ++ch) {
            const int nk4 = d.n_k4 - ch * 4 < 4 ? d.n_k4 - ch * 4 : 4;
            mbar_wait(full_u32 + 8 * stage, phase);
            const unsigned char* st = smem + stage * kDStageBytes;
            steps(c, st + frag_off, reinterpret_cast<const double*>(st + kDKC * kDRowPitch) + lane, nk4);
            release();
        }
        epilogue(c, ur);
    }
}

// One CTA = kDWarps MMA warps + one producer warp.  It owns a batch tile of kDTB columns and a
// run of consecutive units (group g, i-tile it) of one task.  Per unit the producer streams the K
// message rows of the group (16 per stage, one 1-D bulk copy per row, lane = row) and the
// matching W fragments (one bulk copy) through a 4-stage shared-memory ring; every MMA warp
// multiplies the [8 MT x 16] W block into its 16 columns of the rows and keeps an
// [8 MT x 16] accumulator tile in registers (MT <= 4); the epilogue multiplies the per-instance
// s-only operands in and stores out / bel rows with 16-byte stores.
template <typename T>
__global__ void __launch_bounds__((kDWarps + 1) * 32, 2) jt_dense_kernel(const DenseArgs a) {
    // T is the storage type of the workspace rows (float64 or float32); W, the fragments and the
    // accumulators are float64 either way: the float32 pipeline converts its rows on the way into
    // the B fragments and rounds once, at the store
    typedef Pack<T, 2> P2;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bid = blockIdx.x;
    int lo = 0, hi = a.n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.prefix + mid) <= bid) lo = mid; else hi = mid;
    }
    const DDense d = a.dd[lo];
    const int upc = __ldg(a.prefix + a.n + 1 + lo);
    const long long units = (long long)d.n_g * d.n_it;
    const long long u0 = (long long)(bid - __ldg(a.prefix + lo)) * upc;
    const long long u1 = u0 + upc < units ? u0 + upc : units;
    const long long B = a.B;
    const long long col0 = (long long)blockIdx.y * kDTB;
    const int ncols = (int)(B - col0 < kDTB ? B - col0 : kDTB);
    const uint32_t row_bytes = (uint32_t)ncols * (uint32_t)sizeof(T);

    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kDStages * kDStageBytes);
    const uint32_t ring_u32 = smem_u32(smem);
    const uint32_t full_u32 = smem_u32(bars), empty_u32 = smem_u32(bars + kDStages);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDStages; ++i) {
            mbar_init(full_u32 + 8 * i, 1);
            mbar_init(empty_u32 + 8 * i, kDWarps);
        }
        mbar_fence_init();
    }
    // Stale shared memory must be finite where a k-step reads it: the rows past K of the last
    // k-step are multiplied by zero fragments, and 0 * NaN would poison the accumulators.  (Stale
    // columns past the batch only reach accumulator columns that are never stored.)
    if (d.K & 3) {
        for (int i = threadIdx.x; i < kDStages * kDStageBytes / 16; i += blockDim.x)
            reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int* __restrict__ dtab = a.dtab;
    if (warp == kDWarps) {
        // ---------------- producer warp ----------------
        const DMsg* m = a.msgs + d.msg;
        const T* origin = static_cast<const T*>(a.work) + m->eoff + col0;
        const double* wbase = static_cast<const double*>(a.W) + d.w_off;
        const long long unit_w = (long long)d.n_k4 * d.MT * 32;
        int stage = 0;
        uint32_t phase = 0;
        if (d.ups > 1) {
            // short contractions: a stage holds `ups` consecutive units, unit j in rows j * 4 n_k4 ...
            const int kpad = d.n_k4 * 4;
            const int j = lane / kpad, k = lane - j * kpad;
            for (int u = (int)u0; u < (int)u1; u += d.ups) {
                const int nu = (int)u1 - u < d.ups ? (int)u1 - u : d.ups;
                const bool live = j < nu && k < d.K;
                long long row = 0;
                if (live) row = m->off + __ldg(dtab + d.mg + (u + j) / d.n_it) + __ldg(dtab + d.mk + k);
                const uint32_t full = full_u32 + 8 * stage;
                const uint32_t dst = ring_u32 + (uint32_t)stage * kDStageBytes;
                const uint32_t wbytes = (uint32_t)nu * d.n_k4 * d.MT * 256u;
                mbar_wait(empty_u32 + 8 * stage, phase ^ 1);
                if (lane == 0) mbar_expect_tx(full, (uint32_t)nu * d.K * row_bytes + wbytes);
                __syncwarp();
                if (live) bulk_g2s(dst + (uint32_t)lane * kDRowPitch, origin + row * B, row_bytes, full);
                if (lane == kDKC) bulk_g2s(dst + kDKC * kDRowPitch, wbase + u * unit_w, wbytes, full);
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
                if (++stage == kDStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            return;
        }
        for (int u = (int)u0; u < (int)u1; ++u) {
            const int g = u / d.n_it;
            const long long row_g = m->off + __ldg(dtab + d.mg + g);
            for (int c = 0; c < d.n_chunks; ++c) {
                const int k0 = c * kDKC;
                const int rows = d.K - k0 < kDKC ? d.K - k0 : kDKC;
                const int nk4 = d.n_k4 - c * 4 < 4 ? d.n_k4 - c * 4 : 4;
                long long row = 0;
                if (lane < rows) row = row_g + __ldg(dtab + d.mk + k0 + lane);
                const uint32_t full = full_u32 + 8 * stage;
                const uint32_t dst = ring_u32 + (uint32_t)stage * kDStageBytes;
                const uint32_t wbytes = (uint32_t)nk4 * d.MT * 256u;
                mbar_wait(empty_u32 + 8 * stage, phase ^ 1);
                if (lane == 0) mbar_expect_tx(full, (uint32_t)rows * row_bytes + wbytes);
                __syncwarp();
                if (lane < rows) bulk_g2s(dst + (uint32_t)lane * kDRowPitch, origin + row * B, row_bytes, full);
                if (lane == kDKC) bulk_g2s(dst + kDKC * kDRowPitch, wbase + u * unit_w + (long long)c * 4 * d.MT * 32, wbytes, full);
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
                if (++stage == kDStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    // ---------------- MMA warps ----------------
    const int iu0 = (int)u0, iu1 = (int)u1;
    switch (d.MT) {
    case 1: dense_mma_warp<T, 1>(a, d, smem, full_u32, empty_u32, iu0, iu1, col0, warp, lane); break;
    case 2: dense_mma_warp<T, 2>(a, d, smem, full_u32, empty_u32, iu0, iu1, col0, warp, lane); break;
    case 3: dense_mma_warp<T, 3>(a, d, smem, full_u32, empty_u32, iu0, iu1, col0, warp, lane); break;
    default: dense_mma_warp<T, 4>(a, d, smem, full_u32, empty_u32, iu0, iu1, col0, warp, lane); break;
    }
}

bool dense_env_enabled() {   // JT_DISABLE_DENSE=1 keeps every task on the projection kernels (A-B timing)
    static const int on = [] {
        const char* e = getenv("JT_DISABLE_DENSE");
        return (e && e[0] == '1') ? 0 : 1;
    }();
    return on == 1;
}

bool beta_walk_blocks() {   // JT_BETA_WALK=items: the plain s-major walk (A-B timing); default: blocks, see jt_dense_build
    static const int on = [] {
        const char* e = getenv("JT_BETA_WALK");
        return (e && e[0] == 'i') ? 0 : 1;
    }();
    return on == 1;
}

double dense_min_gain() {   // JT_DENSE_MIN_GAIN: items / rows moved from which a task becomes a contraction (default 2)
    static const double v = [] {
        const char* e = getenv("JT_DENSE_MIN_GAIN");
        const double x = e ? atof(e) : 0.0;
        return x >= 1.0 ? x : 2.0;
    }();
    return v;
}

int dense_waves() {          // JT_DENSE_WAVES: waves of 2 CTAs per SM a dense launch is cut into before its CTAs grow (default 4)
    static const int v = [] {
        const char* e = getenv("JT_DENSE_WAVES");
        const int x = e ? atoi(e) : 4;
        return x >= 1 && x <= 64 ? x : 4;
    }();
    return v;
}

bool dense_balance() {      // JT_DENSE_BALANCE=0: units per CTA as the power of two says (A-B timing)
    static const int on = [] {
        const char* e = getenv("JT_DENSE_BALANCE");
        return !(e && e[0] == '0');
    }();
    return on != 0;
}

int beta_block() {          // JT_BETA_BLOCK: consecutive clique entries per block of the walk (default 8)
    static const int n = [] {
        const char* e = getenv("JT_BETA_BLOCK");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= 256 ? v : 8;
    }();
    return n;
}

int phase_group(int phase) {
    if (phase == JT_PHASE_COLLECT_INSTANCE) return 0;
    if (phase == JT_PHASE_DIST_PRE_INSTANCE || phase == JT_PHASE_DIST_MAIN_MESSAGES || phase == JT_PHASE_MARGINAL_DIRECT)
        return 1;
    return -1;
}

// distinct values of a two-level additive table map in order of first appearance
struct Distinct {
    std::vector<int> value;        // value of each class
    std::vector<int> cls;          // class of every index
    std::vector<int> count;
};

Distinct classify(const jt_plan* p, int hi_off, int lo_off, int n, int n_lo) {
    Distinct d;
    d.cls.resize(n);
    if (n <= 0) return d;
    // v(x) = hi[x / n_lo] + lo[x % n_lo]; classes are numbered in the order their value first appears.
    // The values are row offsets inside one message, so their range is usually a small multiple of
    // n at most: a direct table then (plans with 2^17-entry cliques classify tens of millions of
    // entries when they are loaded); a hash map otherwise.
    const int n_hi = (n + n_lo - 1) / n_lo, lo_used = n < n_lo ? n : n_lo;
    const int* hi = p->tab.data() + hi_off;
    const int* lo = p->tab.data() + lo_off;
    long long vmin = (long long)*std::min_element(hi, hi + n_hi) + *std::min_element(lo, lo + lo_used);
    long long vmax = (long long)*std::max_element(hi, hi + n_hi) + *std::max_element(lo, lo + lo_used);
    const long long range = vmax - vmin + 1;
    auto add = [&](int x, int cls) {
        d.cls[x] = cls;
        ++d.count[cls];
    };
    if (range <= 4LL * n + 4096) {
        std::vector<int> slot((size_t)range, -1);
        int x = 0;
        for (int xh = 0; xh < n_hi; ++xh) {
            const int base = hi[xh] - (int)vmin;
            for (int xl = 0; xl < n_lo && x < n; ++xl, ++x) {
                int& c = slot[(size_t)(base + lo[xl])];
                if (c < 0) {
                    c = (int)d.value.size();
                    d.value.push_back(hi[xh] + lo[xl]);
                    d.count.push_back(0);
                }
                add(x, c);
            }
        }
        return d;
    }
    std::unordered_map<int, int> seen;
    int x = 0;
    for (int xh = 0; xh < n_hi; ++xh)
        for (int xl = 0; xl < n_lo && x < n; ++xl, ++x) {
            const int v = hi[xh] + lo[xl];
            auto it = seen.find(v);
            if (it == seen.end()) {
                it = seen.emplace(v, (int)d.value.size()).first;
                d.value.push_back(v);
                d.count.push_back(0);
            }
            add(x, it->second);
        }
    return d;
}

// Walk order of a belief task (position -> (s, r) item), block form: blocks of `beta_block()`
// consecutive clique entries -- a contiguous run of destination rows -- with the blocks that read
// the same row sets next to each other (sorted by a hash of their row indices; a collision only
// costs reuse).  Empty when the task's entries are not a permutation of 0 .. n-1.  A pure function
// of the plan's tables: jt_dense_build computes the walks of all tasks side by side.
std::vector<int> block_walk(const jt_plan* p, const DTask& k) {
    const long long n = (long long)k.n_s * k.n_r;
    std::vector<int> item_of((size_t)n, -1);
    std::vector<unsigned long long> ekey((size_t)n);
    const int n_shi = k.n_s / k.n_slo, n_rhi = k.n_r / k.n_rlo;
    long long item = 0;
    bool once = true;
    for (int sh = 0; sh < n_shi && once; ++sh)
        for (int sl = 0; sl < k.n_slo && once; ++sl) {
            const long long e_s = (long long)p->tab[k.src_shi + sh] + p->tab[k.src_slo + sl];
            unsigned long long hs = 1469598103934665603ULL;
            for (int j = k.smsg_begin; j < k.smsg_end; ++j)
                if (!p->msgs[j].uni)
                    hs = (hs ^ (unsigned long long)(p->msgs[j].off + p->tab[p->msgs[j].a_hi + sh] + p->tab[p->msgs[j].a_lo + sl])) * 1099511628211ULL;
            if (k.own >= 0 && !(k.flags & JT_TF_OWN_UNIFORM))
                hs = (hs ^ (unsigned long long)(k.own + (long long)sh * k.n_slo + sl)) * 1099511628211ULL;
            for (int rh = 0; rh < n_rhi && once; ++rh)
                for (int rl = 0; rl < k.n_rlo; ++rl, ++item) {
                    const long long e = e_s + p->tab[k.src_rhi + rh] + p->tab[k.src_rlo + rl];
                    if (e < 0 || e >= n || item_of[e] != -1) {
                        once = false;
                        break;
                    }
                    item_of[e] = (int)item;
                    unsigned long long h = hs;
                    for (int j = k.rmsg_begin; j < k.rmsg_end; ++j)
                        if (!p->msgs[j].uni)
                            h = (h ^ (unsigned long long)(p->msgs[j].off + p->tab[p->msgs[j].a_hi + sh] + p->tab[p->msgs[j].a_lo + sl] +
                                                          p->tab[p->msgs[j].b_hi + rh] + p->tab[p->msgs[j].b_lo + rl])) * 1099511628211ULL;
                    ekey[e] = h;
                }
        }
    std::vector<int> perm;
    if (!once) return perm;
    const long long W = beta_block();
    const long long nb = (n + W - 1) / W;
    std::vector<std::pair<unsigned long long, int>> blocks((size_t)nb);
    for (long long b = 0; b < nb; ++b) {
        unsigned long long h = 1469598103934665603ULL;
        for (long long e = W * b; e < std::min(n, W * b + W); ++e) h = (h ^ ekey[e]) * 1099511628211ULL;
        blocks[b] = {h, (int)b};
    }
    std::sort(blocks.begin(), blocks.end());
    perm.reserve((size_t)n);
    for (const auto& blk : blocks)
        for (long long e = W * blk.second; e < std::min(n, W * blk.second + W); ++e) perm.push_back(item_of[e]);
    return perm;
}

// ... and of a leaf clique (no r space) in the plain item walk: s sorted by its first per-instance
// row operand (counting sort, stable).  Empty when the task has no such operand.
std::vector<int> leaf_walk(const jt_plan* p, const DTask& k) {
    std::vector<int> perm;
    int hi = 0, lo = 0;
    bool found = false;
    for (int j = k.rmsg_begin; j < k.smsg_end && !found; ++j)
        if (!p->msgs[j].uni) {
            hi = p->msgs[j].a_hi;
            lo = p->msgs[j].a_lo;
            found = true;
        }
    if (!found) return perm;                              // (own has row index s: already sorted)
    std::vector<int> key(k.n_s);
    int kmax = 0;
    for (int sx = 0; sx < k.n_s; ++sx) {
        key[sx] = p->tab[hi + sx / k.n_slo] + p->tab[lo + sx % k.n_slo];
        kmax = std::max(kmax, key[sx]);
    }
    std::vector<int> start((size_t)kmax + 2, 0);
    for (int sx = 0; sx < k.n_s; ++sx) ++start[key[sx] + 1];
    for (int v = 0; v <= kmax; ++v) start[v + 1] += start[v];
    perm.resize(k.n_s);
    for (int sx = 0; sx < k.n_s; ++sx) perm[start[key[sx]]++] = sx;
    return perm;
}

}  // namespace

int jt_dense_build(jt_plan* p) {
    p->dense.clear();
    p->dtab.clear();
    p->dense_w_entries = 0;
    std::vector<int> prep_ids[2];
    if (p->hdr[JT_H_UNI_ENTRIES] > 0) {
        for (auto& L : p->launches) {
            const int group = phase_group(L.phase);
            L.dense_begin = L.dense_end = (int)p->dense.size();
            if (group < 0 || !L.tma_ok) continue;
            for (int t = L.begin; t < L.end; ++t) {
                const DTask& k = p->tasks[t];
                if (k.kind != JT_KIND_PROJECT || k.src < 0 || !(k.flags & JT_TF_SRC_UNIFORM) || k.out < 0) continue;
                if (k.beta >= 0 && L.phase != JT_PHASE_DIST_MAIN_MESSAGES) continue;   // writes the clique belief
                int msg = -1, n_rows = 0;
                for (int j = k.rmsg_begin; j < k.rmsg_end; ++j)
                    if (!p->msgs[j].uni) {
                        msg = j;
                        ++n_rows;
                    }
                if (n_rows != 1) continue;
                const DMsg& m = p->msgs[msg];
                // cheap bound before classifying: at best n_s * n_r items collapse to n_s + n_r rows
                if ((long long)k.n_s * k.n_r < 64) continue;
                Distinct gs = classify(p, m.a_hi, m.a_lo, k.n_s, k.n_slo);
                Distinct ks = classify(p, m.b_hi, m.b_lo, k.n_r, k.n_rlo);
                const int n_g = (int)gs.value.size(), K = (int)ks.value.size();
                const int n_i = k.n_s / n_g, n_q = k.n_r / K;
                bool regular = (long long)n_g * n_i == k.n_s && (long long)K * n_q == k.n_r;
                for (int c : gs.count) regular = regular && c == n_i;
                for (int c : ks.count) regular = regular && c == n_q;
                if (!regular) continue;
                DDense d;
                memset(&d, 0, sizeof(d));
                d.task = t;
                d.msg = msg;
                d.n_g = n_g; d.n_i = n_i; d.K = K; d.n_q = n_q;
                d.MT = std::min(4, (n_i + 7) / 8);
                d.n_it = (n_i + 8 * d.MT - 1) / (8 * d.MT);
                d.n_k4 = (K + 3) / 4;
                d.n_chunks = (d.n_k4 + 3) / 4;
                d.ups = d.n_chunks == 1 ? std::max(1, (kDKC / 4) / d.n_k4) : 1;
                // rows moved: the projection kernel streams one row per (s, r) item; here every
                // group loads its K rows once per i-tile and every output row is written once
                const long long items = (long long)k.n_s * k.n_r;
                const long long moved = (long long)n_g * d.n_it * K + k.n_s;
                if ((long long)n_i * K < 16 || (double)items < dense_min_gain() * (double)moved) continue;
                // a CTA contracts whole units: a long sum into a handful of rows has no parallelism
                // here (the projection kernels split r over blocks)
                if (K > 4096 && (long long)n_g * d.n_it < 64) continue;
                d.w_size = (long long)n_g * d.n_it * d.n_k4 * d.MT * 32;
                if (d.w_size > (1LL << 40) || p->dtab.size() + (size_t)k.n_s + n_g + K + k.n_r > 2000000000ULL) continue;
                d.w_off = p->dense_w_entries;
                p->dense_w_entries += d.w_size;
                d.s_of = (int)p->dtab.size();
                p->dtab.resize(p->dtab.size() + k.n_s);
                {
                    std::vector<int> fill(n_g, 0);
                    for (int s = 0; s < k.n_s; ++s) p->dtab[d.s_of + gs.cls[s] * n_i + fill[gs.cls[s]]++] = s;
                }
                d.mg = (int)p->dtab.size();
                p->dtab.insert(p->dtab.end(), gs.value.begin(), gs.value.end());
                d.mk = (int)p->dtab.size();
                p->dtab.insert(p->dtab.end(), ks.value.begin(), ks.value.end());
                d.r_of = (int)p->dtab.size();
                p->dtab.resize(p->dtab.size() + k.n_r);
                {
                    std::vector<int> fill(K, 0);
                    for (int r = 0; r < k.n_r; ++r) p->dtab[d.r_of + ks.cls[r] * n_q + fill[ks.cls[r]]++] = r;
                }
                prep_ids[group].push_back((int)p->dense.size());
                p->dense.push_back(d);
            }
            L.dense_end = (int)p->dense.size();
        }
    }
    p->dense_group_end[0] = (int)prep_ids[0].size();
    p->dense_group_end[1] = (int)(prep_ids[0].size() + prep_ids[1].size());
    // prep launches: [n ids][n + 1 block prefix]
    for (int g = 0; g < 2; ++g) {
        std::vector<int>& v = p->dense_prep_prefix[g];
        v = prep_ids[g];
        long long acc = 0;
        for (int id : prep_ids[g]) {
            v.push_back((int)acc);
            acc += (p->dense[id].w_size + kThreads - 1) / kThreads;
            if (acc > 2147483647LL) return jt_fail(JT_ERR_INVALID, "dense contraction tables too large");
        }
        v.push_back((int)acc);
    }
    // Clique beliefs of uniform cliques: written by jt_beta_kernel instead of the projection task
    // that sends the last message (DIST_MAIN launches, uniform mode).  Such a task keeps its message
    // part (which may be a dense contraction: the descriptors of the matching DIST_MAIN_MESSAGES
    // launch) and a task that only writes a belief leaves the projection launch altogether.
    const bool small_offsets = p->hdr[JT_H_CLIQUE_ENTRIES] + 3 * p->hdr[JT_H_SEP_ENTRIES] + p->hdr[JT_H_LIK_ENTRIES] < 2147483647LL;
    std::unordered_map<int, int> dense_tasks;              // task -> descriptor
    for (size_t i = 0; i < p->dense.size(); ++i) dense_tasks[p->dense[i].task] = (int)i;
    // every r-dependent input uniform, at most kBetaRows per-instance s-only rows: a scalar task
    auto scalar_ok = [&](const DTask& k) {
        if (k.kind != JT_KIND_PROJECT || k.src < 0 || !(k.flags & JT_TF_SRC_UNIFORM) || k.out < 0 || k.n_r < 2) return false;
        for (int j = k.rmsg_begin; j < k.rmsg_end; ++j)
            if (!p->msgs[j].uni) return false;
        int rows = 0;
        for (int j = k.smsg_begin; j < k.smsg_end; ++j) rows += p->msgs[j].uni ? 0 : 1;
        return rows <= kBetaRows;
    };
    // Which belief tasks are split off, and which walk table each gets (1: blocks, 2: leaf sort, 0: none).
    // Both are functions of the plan alone, so the tables -- tens of millions of entries on plans with
    // 2^17-entry cliques -- are computed side by side on the host's cores first and appended to the
    // dense table in task order below.
    auto beta_rows = [&](const DTask& k) {
        int rows = (k.own >= 0 && !(k.flags & JT_TF_OWN_UNIFORM)) ? 1 : 0;
        for (int j = k.rmsg_begin; j < k.smsg_end; ++j) rows += p->msgs[j].uni ? 0 : 1;
        return rows;
    };
    auto beta_candidate = [&](const jt_plan::Launch& L, int t) {
        if (L.phase != JT_PHASE_DIST_MAIN || p->hdr[JT_H_UNI_ENTRIES] <= 0 || !L.tma_ok || !small_offsets) return false;
        const DTask& k = p->tasks[t];
        if (k.kind != JT_KIND_PROJECT || k.beta < 0 || k.src < 0 || !(k.flags & JT_TF_SRC_UNIFORM)) return false;
        if (beta_rows(k) > kBetaRows) return false;
        return !(k.out >= 0 && !dense_tasks.count(t) && !scalar_ok(k));
    };
    auto walk_kind = [&](const DTask& k, int rows) {
        if (beta_walk_blocks() && rows > 0 && (long long)k.n_s * k.n_r > 8) return 1;
        if (k.n_r == 1 && rows > 0 && k.n_s > 1) return 2;
        return 0;
    };
    auto leaf_has_operand = [&](const DTask& k) {
        for (int j = k.rmsg_begin; j < k.smsg_end; ++j)
            if (!p->msgs[j].uni) return true;
        return false;
    };
    std::unordered_map<int, std::vector<int>> walks;       // task -> precomputed walk
    {
        std::vector<int> jobs;
        long long total = 0;
        for (const auto& L : p->launches)
            for (int t = L.begin; t < L.end; ++t)
                if (beta_candidate(L, t)) {
                    const DTask& k = p->tasks[t];
                    const int w = walk_kind(k, beta_rows(k));
                    if (w == 0 || (w == 2 && !leaf_has_operand(k))) continue;
                    jobs.push_back(t);
                    total += (long long)k.n_s * k.n_r;
                }
        const unsigned hw = std::thread::hardware_concurrency();
        const int n_threads = (int)std::min<size_t>(jobs.size(), std::min(16u, hw ? hw : 1u));
        if (n_threads > 1 && total <= (1LL << 28)) {       // at most 1 GB of tables in flight
            std::vector<std::vector<int>> out(jobs.size());
            std::atomic<size_t> next(0);
            auto work = [&]() {
                for (size_t i = next.fetch_add(1); i < jobs.size(); i = next.fetch_add(1)) {
                    const DTask& k = p->tasks[jobs[i]];
                    out[i] = walk_kind(k, beta_rows(k)) == 1 ? block_walk(p, k) : leaf_walk(p, k);
                }
            };
            std::vector<std::thread> pool;
            for (int i = 1; i < n_threads; ++i) pool.emplace_back(work);
            work();
            for (auto& th : pool) th.join();
            for (size_t i = 0; i < jobs.size(); ++i) walks[jobs[i]] = std::move(out[i]);
        }
    }
    auto take_walk = [&](int t, const DTask& k, int kind) {
        auto it = walks.find(t);
        if (it != walks.end()) {
            std::vector<int> w = std::move(it->second);
            walks.erase(it);
            return w;
        }
        return kind == 1 ? block_walk(p, k) : leaf_walk(p, k);
    };
    for (auto& L : p->launches) {
        L.beta_n = 0;
        L.beta_items = 0;
        L.beta_off = 0;
        if (L.phase != JT_PHASE_DIST_MAIN || p->hdr[JT_H_UNI_ENTRIES] <= 0 || !L.tma_ok || !small_offsets) continue;
        std::vector<int> ids, inv_off;
        for (int t = L.begin; t < L.end; ++t) {
            DTask& k = p->tasks[t];
            if (k.kind != JT_KIND_PROJECT || k.beta < 0 || k.src < 0 || !(k.flags & JT_TF_SRC_UNIFORM)) continue;
            int rows = (k.own >= 0 && !(k.flags & JT_TF_OWN_UNIFORM)) ? 1 : 0;
            for (int j = k.rmsg_begin; j < k.smsg_end; ++j) rows += p->msgs[j].uni ? 0 : 1;
            if (rows > kBetaRows) continue;
            // Only where the rest of the task gets cheap without its belief: a task that sends no
            // message, or whose message part is a dense contraction or a scalar task.  Otherwise
            // the projection kernel would walk all (s, r) items a second time just for the message
            // (measured: Ising 16x16, K = 2 contractions: 22.4 -> 25.8 ms with every writer split).
            if (k.out >= 0 && !dense_tasks.count(t) && !scalar_ok(k)) continue;
            // Walk order (position -> item table): blocks of 8 consecutive clique entries -- a
            // contiguous run of destination rows -- and blocks that read the same 8 row sets next to
            // each other, so that their rows stay in L1 (the blocks are sorted by a hash of their
            // row indices; a collision only costs reuse).  Measured (r02): config 5 at 2,048
            // instances 47.6 -> 42.8 ms against the plain s-major walk, whose r-dependent rows (K rows
            // of 16 KB cycling per s) fall out of L1; config 4 unchanged.  JT_BETA_WALK=items: the
            // s-major walk, with a leaf clique (no r space) sorted by its row operand.
            int perm_off = -1;
            const int walk = walk_kind(k, rows);
            if (walk == 1) {
                const long long n = (long long)k.n_s * k.n_r;
                if (p->dtab.size() + (size_t)n > 2000000000ULL) continue;
                std::vector<int> perm = take_walk(t, k, walk);
                if (!perm.empty()) {
                    perm_off = (int)p->dtab.size();
                    p->dtab.insert(p->dtab.end(), perm.begin(), perm.end());
                }
            } else if (walk == 2) {
                // (a leaf clique without a per-instance message operand: own has row index s, already sorted)
                if (!leaf_has_operand(k)) {
                    // nothing to sort by
                } else {
                    if (p->dtab.size() + (size_t)k.n_s > 2000000000ULL) continue;
                    std::vector<int> perm = take_walk(t, k, walk);
                    perm_off = (int)p->dtab.size();
                    p->dtab.insert(p->dtab.end(), perm.begin(), perm.end());
                }
            }
            k.flags |= JT_TF_BETA_SPLIT;
            ids.push_back(t);
            inv_off.push_back(perm_off);
            L.beta_items += (long long)k.n_s * k.n_r;
        }
        if (ids.empty()) continue;
        L.beta_n = (int)ids.size();
        L.beta_off = p->prefix.size();
        p->prefix.insert(p->prefix.end(), ids.begin(), ids.end());
        p->prefix.insert(p->prefix.end(), inv_off.begin(), inv_off.end());
        for (int ch = kBetaChMin; ch <= kBetaChMax; ++ch) {
            long long acc = 0;
            for (int t : ids) {
                p->prefix.push_back((int)acc);
                acc += ((long long)p->tasks[t].n_s * p->tasks[t].n_r + (1LL << ch) - 1) >> ch;
                if (acc > 2147483647LL) return jt_fail(JT_ERR_INVALID, "belief launch too large");
            }
            p->prefix.push_back((int)acc);
        }
        // the message parts of these tasks may run as the dense contractions of the matching
        // DIST_MAIN_MESSAGES launch (same first task), if every one of those is split
        for (const auto& M : p->launches) {
            if (M.phase != JT_PHASE_DIST_MAIN_MESSAGES || M.begin != L.begin) continue;
            bool all_split = true;
            for (int i = M.dense_begin; i < M.dense_end; ++i)
                all_split = all_split && (p->tasks[p->dense[i].task].flags & JT_TF_BETA_SPLIT);
            if (all_split) {
                L.dense_begin = M.dense_begin;
                L.dense_end = M.dense_end;
            }
        }
    }
    // Scalar tasks: every r-dependent input is uniform, so the sum over r is a per-batch total.  The
    // totals are computed by derived B = 1 projection tasks (copies of the task with its uniform
    // messages only, output behind the entries of the uniform workspace) in two derived launches:
    // after the uniform collect (collect tasks) and after the uniform distribute (the others).
    p->scalar_entries = 0;
    if (p->hdr[JT_H_UNI_ENTRIES] > 0 && small_offsets) {
        const long long work_entries = p->hdr[JT_H_CLIQUE_ENTRIES] + 3 * p->hdr[JT_H_SEP_ENTRIES] + p->hdr[JT_H_LIK_ENTRIES];
        std::unordered_map<int, long long> total_of;        // task -> entry of its totals
        std::vector<DTask> derived[2];
        const size_t n_launches = p->launches.size();
        for (size_t li = 0; li < n_launches; ++li) {
            jt_plan::Launch& L = p->launches[li];
            L.scalar_n = 0;
            L.scalar_off = 0;
            int group = phase_group(L.phase);
            if (L.phase == JT_PHASE_DIST_MAIN && L.beta_n > 0) group = 1;
            if (group < 0 || !L.tma_ok) continue;
            std::vector<int> ids, totals;
            for (int t = L.begin; t < L.end; ++t) {
                const DTask k = p->tasks[t];
                if (!scalar_ok(k)) continue;
                if (k.beta >= 0 && L.phase != JT_PHASE_DIST_MAIN_MESSAGES &&
                    !(L.phase == JT_PHASE_DIST_MAIN && (k.flags & JT_TF_BETA_SPLIT)))
                    continue;
                auto it = total_of.find(t);
                if (it == total_of.end()) {
                    if (work_entries + p->scalar_entries + k.n_s >= 2147483647LL) continue;
                    it = total_of.emplace(t, work_entries + p->scalar_entries).first;
                    p->scalar_entries += k.n_s;
                    DTask d = k;
                    d.out = it->second;
                    d.out_space = 0;
                    d.beta = d.bel = d.own = -1;
                    d.flags = 0;
                    d.rmsg_begin = (int)p->msgs.size();
                    for (int j = k.rmsg_begin; j < k.rmsg_end; ++j) p->msgs.push_back(p->msgs[j]);
                    d.rmsg_end = d.smsg_begin = (int)p->msgs.size();
                    for (int j = k.smsg_begin; j < k.smsg_end; ++j)
                        if (p->msgs[j].uni) p->msgs.push_back(p->msgs[j]);
                    d.smsg_end = (int)p->msgs.size();
                    derived[group].push_back(d);
                }
                ids.push_back(t);
                totals.push_back((int)it->second);
            }
            if (ids.empty()) continue;
            L.scalar_n = (int)ids.size();
            L.scalar_off = p->prefix.size();
            p->prefix.insert(p->prefix.end(), ids.begin(), ids.end());
            p->prefix.insert(p->prefix.end(), totals.begin(), totals.end());
            long long acc = 0;
            for (int t : ids) {
                p->prefix.push_back((int)acc);
                acc += (p->tasks[t].n_s + kScalarRows - 1) / kScalarRows;
            }
            p->prefix.push_back((int)acc);
        }
        for (int g = 0; g < 2; ++g) {
            if (derived[g].empty()) continue;
            jt_plan::Launch D;
            memset(&D, 0, sizeof(D));
            D.phase = g == 0 ? JT_PHASE_X_SCALAR0 : JT_PHASE_X_SCALAR1;
            D.begin = (int)p->tasks.size();
            p->tasks.insert(p->tasks.end(), derived[g].begin(), derived[g].end());
            D.end = (int)p->tasks.size();
            D.level = 0;
            int rc = jt_launch_tables(p, D);
            if (rc != JT_OK) return rc;
            p->launches.push_back(D);
        }
    }
    // which levels of the distribute pass may run their two launches side by side
    for (auto& L : p->launches) {
        L.pre_independent = false;
        if (L.phase != JT_PHASE_DIST_MAIN && L.phase != JT_PHASE_DIST_MAIN_MESSAGES) continue;
        L.pre_independent = true;
        if (L.phase == JT_PHASE_DIST_MAIN_MESSAGES) continue;            // writes no belief at all
        for (const auto& Q : p->launches) {
            if (Q.phase != JT_PHASE_DIST_PRE_INSTANCE || Q.level != L.level) continue;
            for (int t = L.begin; t < L.end; ++t) {
                const DTask& w = p->tasks[t];
                if (w.beta < 0 || (w.flags & JT_TF_SRC_UNIFORM)) continue;   // beliefs of shared cliques touch rows nobody reads
                for (int q = Q.begin; q < Q.end; ++q)
                    if (p->tasks[q].src == w.beta && !(p->tasks[q].flags & JT_TF_SRC_UNIFORM)) L.pre_independent = false;
            }
        }
    }
    p->accel = !p->dense.empty();
    for (const auto& L : p->launches) p->accel = p->accel || L.beta_n > 0 || L.scalar_n > 0;
    // block prefixes of the dense launches and the projection-kernel prefixes of the reduced task sets
    for (auto& L : p->launches) {
        const int n = L.dense_end - L.dense_begin;
        L.dense_stages = 0;
        for (int v = 0; v < 2; ++v) {
            L.total_items_v[v] = L.total_items;
            for (int j = 0; j <= kItemLog2Max; ++j) {
                L.item_prefix_off_v[v][j] = L.item_prefix_off[j];
                L.item_blocks_v[v][j] = L.item_blocks[j];
            }
        }
        for (int j = 0; j <= kDenseJMax; ++j) {
            L.dense_prefix_off[j] = 0;
            L.dense_blocks[j] = 0;
        }
        std::vector<char> skip(L.end - L.begin, 0);
        int v_dense = 0;                       // which reduced set excludes the dense contractions
        if (L.phase == JT_PHASE_DIST_MAIN && L.beta_n > 0) {
            // [0]: without the tasks that only write a belief
            for (int t = L.begin; t < L.end; ++t)
                if ((p->tasks[t].flags & JT_TF_BETA_SPLIT) && p->tasks[t].out < 0) {
                    skip[t - L.begin] = 1;
                    L.total_items_v[0] -= (long long)p->tasks[t].n_s * p->tasks[t].n_r;
                }
            int rc = jt_build_item_prefix(p, L, skip.data(), L.item_prefix_off_v[0], L.item_blocks_v[0]);
            if (rc != JT_OK) return rc;
            L.total_items_v[1] = L.total_items_v[0];
            v_dense = 1;
        }
        for (int i = 0; i < L.scalar_n; ++i) {
            const int t = p->prefix[L.scalar_off + i];
            if (skip[t - L.begin]) continue;
            skip[t - L.begin] = 1;
            L.total_items_v[v_dense] -= (long long)p->tasks[t].n_s * p->tasks[t].n_r;
        }
        if (n == 0) {
            if (L.scalar_n > 0) {
                int rc = jt_build_item_prefix(p, L, skip.data(), L.item_prefix_off_v[v_dense], L.item_blocks_v[v_dense]);
                if (rc != JT_OK) return rc;
            } else if (v_dense == 1) {
                for (int j = 0; j <= kItemLog2Max; ++j) {
                    L.item_prefix_off_v[1][j] = L.item_prefix_off_v[0][j];
                    L.item_blocks_v[1][j] = L.item_blocks_v[0][j];
                }
            }
            continue;
        }
        for (int i = L.dense_begin; i < L.dense_end; ++i) {
            const DDense& d = p->dense[i];
            skip[d.task - L.begin] = 1;
            L.dense_stages += ((long long)d.n_g * d.n_it + d.ups - 1) / d.ups * d.n_chunks;
            L.total_items_v[v_dense] -= (long long)p->tasks[d.task].n_s * p->tasks[d.task].n_r;
        }
        for (int j = 0; j <= kDenseJMax; ++j) {
            L.dense_prefix_off[j] = p->prefix.size();
            long long acc = 0;
            std::vector<int> upc;
            for (int i = L.dense_begin; i < L.dense_end; ++i) {
                const DDense& d = p->dense[i];
                const long long units = (long long)d.n_g * d.n_it;
                long long u = (1LL << j) / d.n_chunks * d.ups;           // a multiple of the units per stage
                u = u < d.ups ? d.ups : u;
                u = u > units ? units : u;
                if (dense_balance()) {
                    // the same number of CTAs, the units spread evenly over them (36 units at 32 per
                    // CTA would run as 32 + 4: the launch lasts as long as its longest CTA)
                    const long long nb = (units + u - 1) / u;
                    u = ((units + nb - 1) / nb + d.ups - 1) / d.ups * d.ups;
                }
                p->prefix.push_back((int)acc);
                upc.push_back((int)u);
                acc += (units + u - 1) / u;
                if (acc > 2147483647LL) return jt_fail(JT_ERR_INVALID, "dense launch too large");
            }
            p->prefix.push_back((int)acc);
            p->prefix.insert(p->prefix.end(), upc.begin(), upc.end());
            L.dense_blocks[j] = acc;
        }
        int rc = jt_build_item_prefix(p, L, skip.data(), L.item_prefix_off_v[v_dense], L.item_blocks_v[v_dense]);
        if (rc != JT_OK) return rc;
    }
    return JT_OK;
}

int jt_dense_upload(jt_plan* p) {
    if (p->dense.empty() && p->dtab.empty()) return JT_OK;
    auto up = [](auto** dst, const auto& v) -> cudaError_t {
        const size_t bytes = v.size() * sizeof(v[0]);
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dst), bytes ? bytes : 16);
        if (e != cudaSuccess || !bytes) return e;
        return cudaMemcpy(*dst, v.data(), bytes, cudaMemcpyHostToDevice);
    };
    JT_CUDA(up(&p->d_dense, p->dense));
    JT_CUDA(up(&p->d_dtab, p->dtab));
    for (int g = 0; g < 2; ++g) JT_CUDA(up(&p->d_dense_prep[g], p->dense_prep_prefix[g]));
    return JT_OK;
}

void jt_dense_free(jt_plan* p) {
    cudaFree(p->d_dense);
    cudaFree(p->d_dtab);
    for (int g = 0; g < 2; ++g) cudaFree(p->d_dense_prep[g]);
}

bool jt_tma_path(int64_t B, int dtype) {
    static const int tma_on = [] {
        const char* e = getenv("JT_DISABLE_TMA");
        return (e && e[0] == '1') ? 0 : 1;
    }();
    const int vec = dtype == JT_F64 ? 2 : 4;       // 16-byte batch vectors
    return tma_on == 1 && B % vec == 0 && B / vec >= 64;
}

bool jt_dense_enabled(const jt_plan* p, int64_t B, int dtype, int flags) {
    return p->accel && dense_env_enabled() && (flags & JT_SR_MASK) == JT_SR_SUM_PRODUCT &&
           (flags & JT_UNIFORM) && !(flags & JT_NO_DENSE) && jt_tma_path(B, dtype);
}

// uniform mode, beliefs stored, the projection tasks on the TMA kernel (whose block prefixes have
// the reduced variants); JT_DISABLE_BETA=1 / JT_NO_DENSE keep the beliefs in the projection tasks
bool jt_beta_enabled(const jt_plan* p, int64_t B, int dtype, int flags) {
    static const int on = [] {
        const char* e = getenv("JT_DISABLE_BETA");
        return (e && e[0] == '1') ? 0 : 1;
    }();
    // together with the dense contractions and scalar tasks that take over the message parts of
    // the writers (alone, the split costs a second pass over the items)
    return on == 1 && !(flags & JT_NO_BELIEFS) && jt_dense_enabled(p, B, dtype, flags);
}

int jt_dense_prepare(const jt_plan* p, int which, int dtype, const void* uni_ws, void* w_region, cudaStream_t stream) {
    const std::vector<int>& v = p->dense_prep_prefix[which];
    const int n = (int)(v.size() - 1) / 2;
    if (n <= 0) return JT_OK;
    PrepArgs a;
    a.dd = p->d_dense;
    a.list = p->d_dense_prep[which];
    a.n = n;
    a.dtab = p->d_dtab;
    a.tasks = p->d_tasks;
    a.msgs = p->d_msgs;
    a.tab = p->d_tab;
    a.uni = uni_ws;
    a.W = w_region;
    const int blocks = v.back();
    if (blocks <= 0) return JT_OK;
    if (dtype == JT_F64) jt_dense_prep_kernel<double><<<(unsigned)blocks, kThreads, 0, stream>>>(a);
    else jt_dense_prep_kernel<float><<<(unsigned)blocks, kThreads, 0, stream>>>(a);
    jt_g_launches.fetch_add(1, std::memory_order_relaxed);
    JT_CUDA(cudaGetLastError());
    return JT_OK;
}

int jt_dense_launch(const jt_plan* p, const jt_plan::Launch& L, void* work, const void* uni_ws, const void* w_region,
                    void* fout, int64_t B, int dtype, int flags, cudaStream_t stream) {
    const int n = L.dense_end - L.dense_begin;
    if (n <= 0) return JT_OK;
    const long long tiles = (B + kDTB - 1) / kDTB;
    // As many CTAs as stage-tiles until the grid covers the machine a few times over (2 CTAs per SM):
    // on deep trees most launches are small and latency-bound, so a CTA should own as little
    // serial work as possible; only large launches amortise the prologue over many stages.
    int j = 0;
    while (j < kDenseJMax && ((L.dense_stages * tiles) >> j) > 148LL * 2 * dense_waves()) ++j;
    const long long gx = L.dense_blocks[j];
    if (gx <= 0) return JT_OK;
    if (gx > 2147483647LL || tiles > 65535)
        return jt_fail(JT_ERR_INVALID, "launch grid %lld x %lld exceeds CUDA limits; split the batch", gx, tiles);
    if (!p->dense_attr_set) {
        JT_CUDA(cudaFuncSetAttribute(jt_dense_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem));
        JT_CUDA(cudaFuncSetAttribute(jt_dense_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem));
        p->dense_attr_set = true;
    }
    DenseArgs a;
    a.dd = p->d_dense + L.dense_begin;
    a.n = n;
    a.prefix = p->d_prefix + L.dense_prefix_off[j];
    a.dtab = p->d_dtab;
    a.tasks = p->d_tasks;
    a.msgs = p->d_msgs;
    a.tab = p->d_tab;
    a.work = work;
    a.uni = uni_ws;
    a.W = w_region;
    a.fout = fout;
    a.B = B;
    a.flags = flags;
    if (dtype == JT_F64)
        jt_dense_kernel<double><<<dim3((unsigned)gx, (unsigned)tiles, 1), (kDWarps + 1) * 32, kDSmem, stream>>>(a);
    else
        jt_dense_kernel<float><<<dim3((unsigned)gx, (unsigned)tiles, 1), (kDWarps + 1) * 32, kDSmem, stream>>>(a);
    jt_g_launches.fetch_add(1, std::memory_order_relaxed);
    JT_CUDA(cudaGetLastError());
    return JT_OK;
}
