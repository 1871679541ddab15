// Launch configuration of the sm_100a kernels, instantiated once per semiring.
// A translation unit jt_sr_<name>.cu includes this header and expands JT_DEFINE_SEMIRING, which
// instantiates every kernel of jt_kernels.cuh for that semiring and exports its launcher table
// (jt_host.h: jt_sr_launchers).
#pragma once

#include <cstdlib>
#include <type_traits>

#include "jt_kernels.cuh"

namespace {

// tile shape for a batch of Bv vectors: bx = min(256, pow2ceil(Bv)), sy = 256 / bx
constexpr int kInitRowsMinLog2 = 6;   // jt_init_rows_kernel from 64 batch vectors per row

inline void pick_tile(long long Bv, int& bx_log2, int& sy_log2) {
    bx_log2 = 0;
    while ((1LL << bx_log2) < Bv && bx_log2 < 8) ++bx_log2;
    sy_log2 = 8 - bx_log2;
}

inline bool tma_enabled() {   // JT_DISABLE_TMA=1 forces the LDG kernel (debugging / A-B timing)
    static const int enabled = [] {
        const char* e = getenv("JT_DISABLE_TMA");
        return (e && e[0] == '1') ? 0 : 1;
    }();
    return enabled == 1;
}

inline int tma_vpt_override() {   // JT_TMA_VPT=1|2 overrides the vectors-per-thread choice (A-B timing)
    static const int vpt = [] {
        const char* e = getenv("JT_TMA_VPT");
        return (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
    }();
    return vpt;
}

inline int tma_min_item_log2() {   // JT_TMA_MIN_ITEMS_LOG2: smallest (s, r) chunk per CTA of the TMA kernel (default 5: 32 items)
    static const int v = [] {
        const char* e = getenv("JT_TMA_MIN_ITEMS_LOG2");
        const int x = e ? atoi(e) : 5;
        return x >= 0 && x <= 12 ? x : 5;
    }();
    return v;
}

inline int tma_ctas_per_sm() {   // JT_TMA_CTAS_PER_SM overrides the CTAs per SM a TMA launch aims at before its chunks grow
    static const int v = [] {
        const char* e = getenv("JT_TMA_CTAS_PER_SM");
        const int x = e ? atoi(e) : 0;
        return x >= 1 && x <= 1024 ? x : 0;
    }();
    return v;
}

// Few instances and long reductions with too few output indices to fill the machine: split r.
inline bool use_splitr(const jt_plan::Launch& L, long long B, bool is_init) {
    if (is_init || B > 64 || L.max_nr < 128) return false;
    return L.max_nr >= 4096 || L.total_s * B < 65536;
}

template <typename SR, int SR_ID>
struct Launcher {
    template <typename T, int VPT>
    static int launch_tma_vpt(const jt_plan* p, const jt_plan::Launch& L, KArgs a, int ct, int variant,
                              cudaStream_t stream) {
        const int tw = ct * VPT;
        const long long tiles = (a.Bv + tw - 1) / tw;
        // (s, r) items per CTA: aim at a number of CTAs per SM over the launch, but keep >= 32 items per CTA
        // so the pipeline fill is amortised; every task gets its own chunk of s for that item count
        int j = tma_min_item_log2();
        // Wide batches (many batch tiles per row) take smaller chunks: the CTAs in flight then write
        // a compact range of rows (tools/micro/write_pattern.cu: 8 KB pieces of 512 KB rows, 64 rows
        // per CTA 7.0 TB/s, 1024 rows per CTA 6.1 TB/s); narrow batches keep long chunks (measured
        // on the Ising grid, one tile: 37.9 ms at 8 CTAs per SM, 39.5 at 64)
        const int per_sm = tma_ctas_per_sm() ? tma_ctas_per_sm() : (tiles >= 16 ? 64 : 8);
        const long long target = 148LL * per_sm;
        const long long total_items = variant ? L.total_items_v[variant - 1] : L.total_items;
        while (j < kItemLog2Max && (total_items * tiles) >> (j + 1) >= target) ++j;
        a.sy_log2 = 0;
        a.bx_log2 = 0;
        a.tasks = p->d_tasks + L.begin;
        a.n_tasks = L.end - L.begin;
        // variants 1, 2: the block prefix gives no blocks to the tasks that run in jt_dense_kernel or
        // only write a belief (jt_beta_kernel)
        a.prefix = p->d_prefix + (variant ? L.item_prefix_off_v[variant - 1][j] : L.item_prefix_off[j]);
        const long long gx = variant ? L.item_blocks_v[variant - 1][j] : L.item_blocks[j];
        if (gx <= 0) return JT_OK;
        if (gx > 2147483647LL || tiles > 65535)
            return jt_fail(JT_ERR_INVALID, "launch grid %lld x %lld exceeds CUDA limits; split the batch", gx, tiles);
        // ring rows, then barriers / row indices / scalar operands (TmaAux)
        const size_t smem = (size_t)(kTmaSlots / VPT) * tw * 16 + sizeof(TmaAux<T>);
        // opt in to > 48 KB of dynamic shared memory once per (plan = device, instantiation)
        bool& attr_set = p->tma_attr_set[SR_ID][sizeof(T) == 8 ? 1 : 0][VPT - 1];
        if (!attr_set) {
            JT_CUDA(cudaFuncSetAttribute(jt_project_tma_kernel<SR, T, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kTmaSlots * 256 * 16 + (int)sizeof(TmaAux<T>)));
            attr_set = true;
        }
        dim3 grid((unsigned)gx, (unsigned)tiles, 1);
        // consumer warps + row producer warp + uniform warp
        jt_project_tma_kernel<SR, T, VPT><<<grid, ct + 64, smem, stream>>>(a);
        jt_g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        return JT_OK;
    }

    template <typename T>
    static int launch_tma(const jt_plan* p, const jt_plan::Launch& L, KArgs a, int variant, cudaStream_t stream) {
        const int ct = a.Bv >= 256 ? 256 : (a.Bv >= 128 ? 128 : 64);
        // two vectors per consumer thread once the batch fills 512-vector tiles: halves the control
        // instructions per byte of the consumers
        const int forced = tma_vpt_override();
        const bool two = forced ? forced == 2 : a.Bv >= 512;
        // (measured and rejected, r02: the same tile with half the consumer threads and two vectors each
        // for batches below 512 vectors -- Ising 16x16 uniform 38.0 -> 36.9 ms, but per instance 82.4 -> 85.4
        // and config 4 per instance 7.33 -> 8.00 ms)
        if (two && ct == 256) return launch_tma_vpt<T, 2>(p, L, a, ct, variant, stream);
        return launch_tma_vpt<T, 1>(p, L, a, ct, variant, stream);
    }

    template <typename T, int VEC>
    static int launch_tasks(const jt_plan* p, const jt_plan::Launch& L, KArgs a, int variant, cudaStream_t stream) {
        const bool is_init = jt_is_init_phase(L.phase);
        if (!is_init && VEC * sizeof(T) == 16 && L.tma_ok && a.Bv >= 64 && tma_enabled())
            return launch_tma<T>(p, L, a, variant, stream);
        if (variant) return jt_fail(JT_ERR_INVALID, "internal: dense variant requested for a launch off the TMA path");
        int bx_log2, sy_log2;
        pick_tile(a.Bv, bx_log2, sy_log2);
        if (VEC == 1 && use_splitr(L, a.B, is_init)) {
            // few instances, long reductions: one block per output index, threads split r
            a.bx_log2 = bx_log2;
            a.sy_log2 = 0;
            a.tasks = p->d_tasks + L.begin;
            a.n_tasks = L.end - L.begin;
            a.prefix = p->d_prefix + L.prefix_off[0];
            const long long gx = L.blocks[0];
            const long long gy = (a.B + (1LL << bx_log2) - 1) >> bx_log2;
            if (gx <= 0) return JT_OK;
            if (gx > 2147483647LL || gy > 65535)
                return jt_fail(JT_ERR_INVALID, "launch grid %lld x %lld exceeds CUDA limits; split the batch", gx, gy);
            jt_project_splitr_kernel<SR, T><<<dim3((unsigned)gx, (unsigned)gy, 1), kThreads, 0, stream>>>(a);
            jt_g_launches.fetch_add(1, std::memory_order_relaxed);
            JT_CUDA(cudaGetLastError());
            return JT_OK;
        }
        if (is_init) {
            // a thread walks ~32 rows of s (256 when a block spans one row and shares the row
            // lookups) so the per-instance factor offsets stay in registers
            sy_log2 = bx_log2 >= kInitRowsMinLog2 ? 8 : (sy_log2 + 5 > kMaxSyLog2 ? kMaxSyLog2 : sy_log2 + 5);
            while (sy_log2 > 8 - bx_log2 &&
                   (L.total_s >> sy_log2) * ((a.Bv + (1LL << bx_log2) - 1) >> bx_log2) < 148 * 8)
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
            return jt_fail(JT_ERR_INVALID, "launch grid %lld x %lld exceeds CUDA limits; split the batch", gx, gy);
        dim3 grid((unsigned)gx, (unsigned)gy, 1);
        if (is_init && bx_log2 >= kInitRowsMinLog2)      // a block spans 1-4 whole rows: shared row lookups
            jt_init_rows_kernel<SR, T, VEC><<<grid, kThreads, 0, stream>>>(a);
        else if (is_init)
            jt_init_kernel<SR, T, VEC><<<grid, kThreads, 0, stream>>>(a);
        else
            jt_project_kernel<SR, T, VEC><<<grid, kThreads, 0, stream>>>(a);
        jt_g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        return JT_OK;
    }

    static int dispatch(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a_in, int dtype, int vec, int variant,
                        cudaStream_t stream) {
        KArgs a = a_in;
        if (use_splitr(L, a.B, jt_is_init_phase(L.phase))) {   // split-r kernel: scalar batch lanes
            vec = 1;
            a.Bv = a.B;
        }
        if (dtype == JT_F64) {
            if (vec == 2) return launch_tasks<double, 2>(p, L, a, variant, stream);
            return launch_tasks<double, 1>(p, L, a, variant, stream);
        }
        if (vec == 4) return launch_tasks<float, 4>(p, L, a, variant, stream);
        if (vec == 2) return launch_tasks<float, 2>(p, L, a, variant, stream);
        return launch_tasks<float, 1>(p, L, a, variant, stream);
    }

    template <typename T, int VEC>
    static int launch_beta(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, cudaStream_t stream) {
        BetaArgs b;
        int bx_log2, sy_log2;
        pick_tile(a.Bv, bx_log2, sy_log2);
        const long long gy = (a.Bv + (1LL << bx_log2) - 1) >> bx_log2;
        // items per block: as large as still gives ~8 blocks per SM
        int ch = kBetaChMax;
        while (ch > kBetaChMin && ((L.beta_items >> ch) + L.beta_n) * gy < 148LL * 8) --ch;
        const int* prefix = p->prefix.data() + L.beta_off + 2 * L.beta_n + (size_t)(ch - kBetaChMin) * (L.beta_n + 1);
        const long long gx = prefix[L.beta_n];
        if (gx <= 0) return JT_OK;
        if (gx > 2147483647LL || gy > 65535)
            return jt_fail(JT_ERR_INVALID, "launch grid %lld x %lld exceeds CUDA limits; split the batch", gx, gy);
        b.list = p->d_prefix + L.beta_off;
        b.perm = p->d_dtab;
        b.prefix = p->d_prefix + L.beta_off + 2 * L.beta_n + (size_t)(ch - kBetaChMin) * (L.beta_n + 1);
        b.n = L.beta_n;
        b.ch_log2 = ch;
        b.tasks = p->d_tasks;
        b.msgs = p->d_msgs;
        b.tab = p->d_tab;
        b.work = a.work;
        b.uni = a.uni;
        b.B = a.B;
        b.Bv = a.Bv;
        b.bx_log2 = bx_log2;
        jt_beta_kernel<SR, T, VEC><<<dim3((unsigned)gx, (unsigned)gy, 1), kThreads, 0, stream>>>(b);
        jt_g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        return JT_OK;
    }

    static int beta(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, int dtype, int vec, cudaStream_t stream) {
        if (L.beta_n <= 0) return JT_OK;
        if (dtype == JT_F64) {
            if (vec == 2) return launch_beta<double, 2>(p, L, a, stream);
            return launch_beta<double, 1>(p, L, a, stream);
        }
        if (vec == 4) return launch_beta<float, 4>(p, L, a, stream);
        if (vec == 2) return launch_beta<float, 2>(p, L, a, stream);
        return launch_beta<float, 1>(p, L, a, stream);
    }

    template <typename T, int VEC>
    static int launch_scalar(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, cudaStream_t stream) {
        ScalarArgs b;
        int bx_log2, sy_log2;
        pick_tile(a.Bv, bx_log2, sy_log2);
        const long long gy = (a.Bv + (1LL << bx_log2) - 1) >> bx_log2;
        const long long gx = p->prefix[L.scalar_off + 3 * (size_t)L.scalar_n];
        if (gx <= 0) return JT_OK;
        if (gx > 2147483647LL || gy > 65535)
            return jt_fail(JT_ERR_INVALID, "launch grid %lld x %lld exceeds CUDA limits; split the batch", gx, gy);
        b.list = p->d_prefix + L.scalar_off;
        b.n = L.scalar_n;
        b.tasks = p->d_tasks;
        b.msgs = p->d_msgs;
        b.tab = p->d_tab;
        b.work = a.work;
        b.uni = a.uni;
        b.fout = a.fout;
        b.B = a.B;
        b.Bv = a.Bv;
        b.bx_log2 = bx_log2;
        b.flags = a.flags;
        jt_scalar_kernel<SR, T, VEC><<<dim3((unsigned)gx, (unsigned)gy, 1), kThreads, 0, stream>>>(b);
        jt_g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        return JT_OK;
    }

    static int scalar(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, int dtype, int vec, cudaStream_t stream) {
        if (L.scalar_n <= 0) return JT_OK;
        if (dtype == JT_F64) {
            if (vec == 2) return launch_scalar<double, 2>(p, L, a, stream);
            return launch_scalar<double, 1>(p, L, a, stream);
        }
        if (vec == 4) return launch_scalar<float, 4>(p, L, a, stream);
        if (vec == 2) return launch_scalar<float, 2>(p, L, a, stream);
        return launch_scalar<float, 1>(p, L, a, stream);
    }

    static int contract(const KArgs& a, long long blocks, long long gy, int dtype, int vec, cudaStream_t stream) {
        dim3 grid((unsigned)blocks, (unsigned)gy, 1);
        if (dtype == JT_F64) {
            if (vec == 2) jt_project_kernel<SR, double, 2><<<grid, kThreads, 0, stream>>>(a);
            else jt_project_kernel<SR, double, 1><<<grid, kThreads, 0, stream>>>(a);
        } else {
            if (vec == 4) jt_project_kernel<SR, float, 4><<<grid, kThreads, 0, stream>>>(a);
            else if (vec == 2) jt_project_kernel<SR, float, 2><<<grid, kThreads, 0, stream>>>(a);
            else jt_project_kernel<SR, float, 1><<<grid, kThreads, 0, stream>>>(a);
        }
        jt_g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        return JT_OK;
    }

    // the whole propagation of a few instances of a small tree in one launch (jt_walk_kernel)
    static int walk(const KArgs& a, const jt_walk_args& w, int dtype, cudaStream_t stream) {
        WalkArgs d;
        d.seq = w.seq; d.n_seq = w.n_seq; d.evidence = w.evidence; d.n_evid = w.n_evid; d.ev_card = w.ev_card;
        d.evf_ptr = w.evf_ptr; d.evf_var = w.evf_var; d.evf_stride = w.evf_stride; d.n_factors = w.n_factors;
        d.errors = w.errors;
        d.lik_base = w.lik_base;
        d.lik_entries = w.lik_entries;
        d.preload = w.preload;
        d.work_entries = w.work_entries;
        d.n_tasks = w.n_tasks;
        d.n_msgs = w.n_msgs;
        d.n_tab = w.n_tab;
        // the instance's column of the workspace and the schedule fit the default 48 KB of shared
        // memory: mirror the column, stage the schedule
        const size_t smem = ((size_t)w.work_entries * (dtype == JT_F64 ? 8 : 4) + 15) / 16 * 16 +
                            (size_t)w.n_tasks * sizeof(DTask) + (size_t)w.n_msgs * sizeof(DMsg) +
                            (size_t)w.n_tab * 4 + (size_t)w.n_seq * 8;
        const bool sm = smem <= 48 * 1024;
        if (dtype == JT_F64) {
            if (sm) jt_walk_kernel<SR, double, true><<<(unsigned)a.B, kThreads, smem, stream>>>(a, d);
            else jt_walk_kernel<SR, double, false><<<(unsigned)a.B, kThreads, 0, stream>>>(a, d);
        } else {
            if (sm) jt_walk_kernel<SR, float, true><<<(unsigned)a.B, kThreads, smem, stream>>>(a, d);
            else jt_walk_kernel<SR, float, false><<<(unsigned)a.B, kThreads, 0, stream>>>(a, d);
        }
        jt_g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        return JT_OK;
    }

    static int normalize(const jt_plan* p, int64_t B, int dtype, void* factor_out, void* logz, int normalize,
                         cudaStream_t stream) {
        const int n_out = (int)p->fout_off.size();
        // log Z only: the total of scope 0, the outputs stay as they are
        dim3 grid((unsigned)((B + kThreads - 1) / kThreads), (unsigned)(normalize ? n_out : 1), 1);
        if (dtype == JT_F64)
            jt_normalize_kernel<SR, double><<<grid, kThreads, 0, stream>>>(
                static_cast<double*>(factor_out), p->d_out, p->d_out + n_out, B, static_cast<double*>(logz), normalize);
        else
            jt_normalize_kernel<SR, float><<<grid, kThreads, 0, stream>>>(
                static_cast<float*>(factor_out), p->d_out, p->d_out + n_out, B, static_cast<float*>(logz), normalize);
        jt_g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        return JT_OK;
    }
};

}  // namespace

#define JT_DEFINE_SEMIRING(SR, ID, NAME)                                                       \
    const jt_sr_launchers* NAME() {                                                            \
        static const jt_sr_launchers table = {&Launcher<SR, ID>::dispatch, &Launcher<SR, ID>::beta, \
                                              &Launcher<SR, ID>::scalar,                         \
                                              &Launcher<SR, ID>::contract, &Launcher<SR, ID>::normalize, \
                                              &Launcher<SR, ID>::walk};                        \
        return &table;                                                                         \
    }
