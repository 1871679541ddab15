// libjt_b200: C ABI (include/jt_b200.h) over the sm_100a kernels.
// Plan parsing and validation, workspace layout, stage sequencing.  The kernels live in
// jt_kernels.cuh and are instantiated once per semiring by jt_sr_*.cu; this file picks the
// launcher table by the JT_SR_* bits of the stage flags.

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler injects its library

#include "jt_host.h"

// ------------------------------------------------------------------------------------------
// error handling

static thread_local char g_err[512] = "";
std::atomic<int64_t> jt_g_launches{0};

int jt_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define fail jt_fail

namespace {

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

template <typename T>
__global__ void __launch_bounds__(kThreads)
jt_ratio_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T d = b[i];
        out[i] = d != T(0) ? a[i] / d : T(0);
    }
}

}  // namespace

namespace {

// NVTX range per stage and per schedule level (JT_NVTX=1): what a timeline profiler shows as
// "collect level 17" etc. around the launches of that level.
bool nvtx_enabled() {
    static const int on = [] {
        const char* e = getenv("JT_NVTX");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return on == 1;
}

struct NvtxRange {
    bool on;
    NvtxRange(const char* what, int level = -1) : on(nvtx_enabled()) {
        if (!on) return;
        char buf[64];
        if (level >= 0) snprintf(buf, sizeof(buf), "jt %s level %d", what, level);
        else snprintf(buf, sizeof(buf), "%s", what);
        nvtxRangePushA(buf);
    }
    ~NvtxRange() {
        if (on) nvtxRangePop();
    }
};

const char* phase_name(int phase) {
    switch (phase) {
        case JT_PHASE_INIT: case JT_PHASE_INIT_UNIFORM: case JT_PHASE_INIT_INSTANCE: return "init";
        case JT_PHASE_COLLECT: case JT_PHASE_COLLECT_INSTANCE: return "collect";
        case JT_PHASE_COLLECT_UNIFORM: return "collect (uniform)";
        case JT_PHASE_DIST_UNIFORM: return "distribute (uniform)";
        case JT_PHASE_MARGINAL: case JT_PHASE_MARGINAL_DIRECT: return "marginal";
        default: return "distribute";
    }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t dtype_size(int dtype) { return dtype == JT_F64 ? 8 : 4; }

// [ work: entries x B | fbase: F x B int32 | error counter | uniform workspace: entries x 1 | W region ]
// entries = cliques + 3 x separators (beliefs, up, down) + likelihood tables
struct WorkspaceLayout {
    size_t work_bytes, fbase_off, err_off, uni_off, dense_off, total;
};

WorkspaceLayout workspace_layout(const jt_plan* p, int64_t B, int dtype) {
    WorkspaceLayout w;
    const int64_t entries = p->hdr[JT_H_CLIQUE_ENTRIES] + 3 * p->hdr[JT_H_SEP_ENTRIES] + p->hdr[JT_H_LIK_ENTRIES];
    w.work_bytes = align_up((size_t)entries * (size_t)B * dtype_size(dtype), 256);
    w.fbase_off = w.work_bytes;
    const size_t fbase = p->hdr[JT_H_NEVID] > 0 ? (size_t)p->hdr[JT_H_NFACTORS] * (size_t)B * 4 : 0;
    w.err_off = w.fbase_off + align_up(fbase, 256);
    w.uni_off = w.err_off + 256;
    // uniform workspace: the same entries with B = 1, then the totals of the scalar tasks (jt_dense.cu)
    const size_t uni = p->hdr[JT_H_UNI_ENTRIES] > 0 ? (size_t)(entries + p->scalar_entries) * dtype_size(dtype) : 0;
    // W region of the dense contractions (jt_dense.cu), rebuilt with the uniform workspace
    w.dense_off = w.uni_off + align_up(uni, 256);
    w.total = w.dense_off + align_up((size_t)p->dense_w_entries * 8, 256);      // W is float64 for both dtypes
    return w;
}

int pick_vec(int64_t B, int dtype) {
    const int maxv = dtype == JT_F64 ? 2 : 4;
    for (int v = maxv; v > 1; v >>= 1)
        if (B % v == 0) return v;
    return 1;
}

const jt_sr_launchers* launchers(int flags) {
    switch (flags & JT_SR_MASK) {
        case JT_SR_MAX_PRODUCT: return jt_sr_max_product();
        case JT_SR_LOG_SUM_EXP: return jt_sr_log_sum_exp();
        case JT_SR_MAX_SUM: return jt_sr_max_sum();
        default: return jt_sr_sum_product();
    }
}

// tile shape for a batch of Bv vectors: bx = min(256, pow2ceil(Bv)), sy = 256 / bx
void pick_tile(long long Bv, int& bx_log2, int& sy_log2) {
    bx_log2 = 0;
    while ((1LL << bx_log2) < Bv && bx_log2 < 8) ++bx_log2;
    sy_log2 = 8 - bx_log2;
}

int check_common(const jt_plan* p, int64_t B, int dtype, const void* workspace) {
    if (!p) return fail(JT_ERR_INVALID, "plan is null");
    if (B <= 0) return fail(JT_ERR_INVALID, "batch size must be positive");
    if (dtype != JT_F32 && dtype != JT_F64) return fail(JT_ERR_INVALID, "dtype must be JT_F32 or JT_F64");
    if (!workspace) return fail(JT_ERR_INVALID, "workspace is null");
    if (p->device < 0) return fail(JT_ERR_INVALID, "plan not uploaded: call jt_plan_upload first");
    int dev = -1;
    JT_CUDA(cudaGetDevice(&dev));
    if (dev != p->device)
        return fail(JT_ERR_INVALID, "plan lives on device %d but the current device is %d", p->device, dev);
    return JT_OK;
}

KArgs base_args(const jt_plan* p, int64_t B, void* workspace, int vec) {
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.msgs = p->d_msgs;
    a.tab = p->d_tab;
    a.work = workspace;
    a.B = B;
    a.Bv = B / vec;
    return a;
}

// Side stream of the calling thread (per device) and a pool of events for fork / join with it.
// The clique beliefs of uniform cliques (jt_beta_kernel) are final outputs nobody reads during the
// distribute pass, so they run beside the level-by-level message chain, which alone leaves most of
// the machine idle on deep trees.  Thread-local: concurrent calls from several host threads never
// share a stream or an event; everything is joined back into the caller's stream before the stage
// call returns, so the "enqueue only, never synchronise" contract (and graph capture) holds.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaStream_t aux[4] = {nullptr, nullptr, nullptr, nullptr};   // the independent kernels of one level run side by side
    cudaStream_t pre = nullptr;                  // the message-only launch of a distribute level beside its belief-writing one
    cudaStream_t uni = nullptr;                  // the uniform (B = 1) part of the distribute pass, beside the instance collect
    // jt_collect has already run the uniform part of the distribute pass for this (plan, workspace)
    const void* uni_dist_plan = nullptr;
    const void* uni_dist_ws = nullptr;
    int device = -1;
    std::vector<cudaEvent_t> events;
    size_t used = 0;
};
thread_local SideStream g_side;

bool beta_overlap_enabled() {       // JT_BETA_STREAM=0: beliefs in the caller's stream, level by level (A-B timing)
    static const int on = [] {
        const char* e = getenv("JT_BETA_STREAM");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on == 1;
}

cudaError_t side_acquire(cudaStream_t* out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!g_side.stream || g_side.device != dev) {
        g_side.events.clear();                 // events of another device are left to the driver
        e = cudaStreamCreateWithFlags(&g_side.stream, cudaStreamNonBlocking);
        for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&g_side.aux[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g_side.pre, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g_side.uni, cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        g_side.device = dev;
    }
    if (out) *out = g_side.stream;
    return cudaSuccess;
}

bool uniform_overlap_enabled() {    // JT_UNIFORM_OVERLAP=0: the uniform part of the distribute pass inside jt_distribute (A-B timing)
    static const int on = [] {
        const char* e = getenv("JT_UNIFORM_OVERLAP");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on == 1;
}

double beta_min_bytes() {           // JT_BETA_MIN_MB: belief bytes of a launch from which they are split off (default 1024)
    static const double v = [] {
        const char* e = getenv("JT_BETA_MIN_MB");
        return (e && e[0]) ? atof(e) * 1e6 : 1024e6;
    }();
    return v;
}

long long beta_max_batch() {        // JT_BETA_MAX_B: largest batch whose beliefs are split off (default 4096: rows of 32 KB)
    static const long long v = [] {
        const char* e = getenv("JT_BETA_MAX_B");
        return (e && e[0]) ? atoll(e) : 4096LL;
    }();
    return v;
}

bool level_fork_enabled() {         // JT_LEVEL_STREAMS=0: the kernels of a level one after the other (A-B timing)
    static const int on = [] {
        const char* e = getenv("JT_LEVEL_STREAMS");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on == 1;
}

cudaError_t side_event(cudaEvent_t* out) {
    if (g_side.used == g_side.events.size()) {
        cudaEvent_t ev;
        cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        g_side.events.push_back(ev);
    }
    *out = g_side.events[g_side.used++];
    return cudaSuccess;
}

// One launch of the plan.  With `w_region` set (uniform mode, dense contractions enabled) the
// tasks that are dense contractions run in jt_dense_kernel and the projection kernel gets the
// block prefix without them.  A DIST_MAIN launch in uniform mode (`split`) leaves the clique
// beliefs of the uniform cliques to jt_beta_kernel: the projection kernel (and the dense
// contractions of the matching message-only launch) compute the messages, tasks that only write
// a belief drop out of it.
int run_launch(jt_plan* p, const jt_plan::Launch& L, const KArgs& a_in, int dtype, int vec, void* w_region,
               cudaStream_t stream, bool split = false, cudaStream_t beta_stream = nullptr, int aux_set = 0) {
    NvtxRange range(phase_name(L.phase), L.level);
    KArgs a = a_in;
    // Splitting the beliefs off pays for the extra kernels (a second latency-bound launch in the
    // chain, the rows read once more) only when the launch writes a lot: jt_beta_kernel streams at
    // ~5.9 TB/s against 3.6-5.4 TB/s for the same rows written from inside the projection tasks, i.e.
    // some tens of microseconds per GB.  Measured (r02): config 4 (3.2-6.4 GB per launch) 2.75 -> 2.06 ms,
    // Ising 16x16 at B = 128 (0.13 GB per launch) 22.2 -> 22.9 ms.  With batch rows of 64 KB and more the
    // projection tasks already write at that rate (two vectors per thread), so nothing is gained.
    split = split && L.phase == JT_PHASE_DIST_MAIN && L.beta_n > 0 &&
            (double)L.beta_items * (double)a.B * (dtype == JT_F64 ? 8.0 : 4.0) >= beta_min_bytes() && a.B <= beta_max_batch();
    // dense contractions and scalar tasks leave the projection launch together (one reduced task set)
    const bool accel = w_region && (L.dense_end > L.dense_begin || L.scalar_n > 0) &&
                       (L.phase != JT_PHASE_DIST_MAIN || split);
    int variant = 0;
    if (split) {
        a.flags |= JT_X_BETA_SPLIT;
        variant = 1;
    }
    if (accel) variant = split ? 2 : 1;
    const bool has_dense = accel && L.dense_end > L.dense_begin, has_scalar = accel && L.scalar_n > 0;
    const bool has_proj = (variant ? L.total_items_v[variant - 1] : L.total_items) > 0;
    // The dense contractions, the scalar tasks and the projection tasks of a launch are independent
    // of each other: on deep, narrow trees each of them is a small latency-bound kernel, so they
    // run side by side (fork after everything enqueued so far, join before the next launch).
    cudaStream_t s_dense = stream, s_scalar = stream;
    cudaEvent_t ev;
    const int n_kernels = (has_dense ? 1 : 0) + (has_scalar ? 1 : 0) + (has_proj ? 1 : 0);
    const bool fork = n_kernels >= 2 && level_fork_enabled();
    if (fork) {
        JT_CUDA(side_acquire(nullptr));
        JT_CUDA(side_event(&ev));
        JT_CUDA(cudaEventRecord(ev, stream));
        if (has_dense && (has_proj || has_scalar)) {
            s_dense = g_side.aux[2 * aux_set];
            JT_CUDA(cudaStreamWaitEvent(s_dense, ev, 0));
        }
        if (has_scalar && has_proj) {
            s_scalar = g_side.aux[2 * aux_set + 1];
            JT_CUDA(cudaStreamWaitEvent(s_scalar, ev, 0));
        }
    }
    int rc = JT_OK;
    if (has_dense) rc = jt_dense_launch(p, L, a.work, a.uni, w_region, a.fout, a.B, dtype, a.flags, s_dense);
    if (rc == JT_OK && has_scalar) rc = launchers(a.flags)->scalar(p, L, a, dtype, vec, s_scalar);
    if (rc == JT_OK && (has_proj || !accel)) rc = launchers(a.flags)->dispatch(p, L, a, dtype, vec, variant, stream);
    if (rc != JT_OK) return rc;
    if (fork) {
        if (s_dense != stream) {
            JT_CUDA(side_event(&ev));
            JT_CUDA(cudaEventRecord(ev, s_dense));
            JT_CUDA(cudaStreamWaitEvent(stream, ev, 0));
        }
        if (s_scalar != stream) {
            JT_CUDA(side_event(&ev));
            JT_CUDA(cudaEventRecord(ev, s_scalar));
            JT_CUDA(cudaStreamWaitEvent(stream, ev, 0));
        }
    }
    if (!split) return JT_OK;
    return launchers(a.flags)->beta(p, L, a, dtype, vec, beta_stream ? beta_stream : stream);
}

// Launches of one phase, in plan order.
int run_phase(jt_plan* p, int phase, const KArgs& a, int dtype, int vec, cudaStream_t stream, void* w_region = nullptr) {
    for (const auto& L : p->launches) {
        if (L.phase != phase) continue;
        int rc = run_launch(p, L, a, dtype, vec, w_region, stream);
        if (rc != JT_OK) return rc;
    }
    return JT_OK;
}

// A phase executed once for the whole batch: an ordinary B = 1 launch on the uniform workspace.
int run_phase_uniform(jt_plan* p, int phase, KArgs a, int dtype, void* uni_ws, cudaStream_t stream) {
    a.work = uni_ws;
    a.uni = nullptr;
    a.uniform = 0;
    a.fbase = nullptr;
    a.B = 1;
    a.Bv = 1;
    return run_phase(p, phase, a, dtype, 1, stream);
}

bool uniform_mode(const jt_plan* p, int flags) {
    return (flags & JT_UNIFORM) && p->hdr[JT_H_UNI_ENTRIES] > 0;
}

void* uniform_ws(const jt_plan* p, int64_t B, int dtype, void* workspace) {
    return static_cast<char*>(workspace) + workspace_layout(p, B, dtype).uni_off;
}

// W region of the workspace when this call runs its dense contractions in jt_dense_kernel, else null
void* dense_region(const jt_plan* p, int64_t B, int dtype, void* workspace, int flags) {
    if (!jt_dense_enabled(p, B, dtype, flags)) return nullptr;
    return static_cast<char*>(workspace) + workspace_layout(p, B, dtype).dense_off;
}

// The uniform part of the distribute pass: down-messages with an evidence-free source side, once,
// top level first (B = 1 launches in the uniform workspace), then the W blocks and totals of the
// distribute and marginal tasks, which also need those messages.
int uniform_distribute(jt_plan* p, const KArgs& a, int dtype, void* uni, void* w_region, cudaStream_t stream) {
    int rc = run_phase_uniform(p, JT_PHASE_DIST_UNIFORM, a, dtype, uni, stream);
    if (rc != JT_OK || !w_region) return rc;
    rc = jt_dense_prepare(p, 1, dtype, uni, w_region, stream);
    if (rc == JT_OK) rc = run_phase_uniform(p, JT_PHASE_X_SCALAR1, a, dtype, uni, stream);
    return rc;
}

}  // namespace

// Block prefix of a launch of the TMA projection kernel for every item-count target 2^j: per task a
// chunk of s sized for ~2^j (s, r) items per CTA; layout per j: [n_tasks + 1] first block of each
// task, [n_tasks] log2 of the task's chunk.  Tasks flagged in `skip` get no blocks.
int jt_build_item_prefix(jt_plan* p, const jt_plan::Launch& L, const char* skip, size_t* off_out, long long* blocks_out) {
    for (int j = 0; j <= kItemLog2Max; ++j) {
        off_out[j] = p->prefix.size();
        long long acc = 0;
        std::vector<int> chunk;
        for (int t = L.begin; t < L.end; ++t) {
            const DTask& k = p->tasks[t];
            int nr_log2 = 0;
            while ((1LL << nr_log2) < k.n_r) ++nr_log2;
            int c = j - nr_log2;
            c = c < 0 ? 0 : (c > 30 ? 30 : c);
            while (c > 0 && (1LL << c) >= 2LL * k.n_s) --c;      // no larger than the task
            p->prefix.push_back((int)acc);
            chunk.push_back(c);
            if (!(skip && skip[t - L.begin])) acc += ((long long)k.n_s + (1LL << c) - 1) >> c;
            if (acc > 2147483647LL) return fail(JT_ERR_INVALID, "launch too large");
        }
        p->prefix.push_back((int)acc);
        p->prefix.insert(p->prefix.end(), chunk.begin(), chunk.end());
        blocks_out[j] = acc;
    }
    return JT_OK;
}

// Everything a launch needs besides its task range: item totals, block prefixes per tile shape
// and per item-count target, which kernels it may use.
int jt_launch_tables(jt_plan* p, jt_plan::Launch& L) {
    L.tma_ok = true;
    L.min_nr = 2147483647;
    L.max_nr = 0;
    L.total_s = 0;
    L.total_items = 0;
    for (int t = L.begin; t < L.end; ++t) L.total_items += (long long)p->tasks[t].n_s * p->tasks[t].n_r;
    int rc = jt_build_item_prefix(p, L, nullptr, L.item_prefix_off, L.item_blocks);
    if (rc != JT_OK) return rc;
    for (int t = L.begin; t < L.end; ++t) {
        const DTask& k = p->tasks[t];
        const int rows = (k.src >= 0 ? 1 : 0) + (k.rmsg_end - k.rmsg_begin) + (k.smsg_end - k.smsg_begin) +
                         (k.own >= 0 ? 1 : 0);
        if (rows < 1 || rows > kTmaMaxRows || k.rmsg_end != k.smsg_begin) L.tma_ok = false;
        L.min_nr = k.n_r < L.min_nr ? k.n_r : L.min_nr;
        L.max_nr = k.n_r > L.max_nr ? k.n_r : L.max_nr;
        L.total_s += k.n_s;
    }
    // block prefix per tile shape: a block covers 2^sy consecutive values of s of one task
    for (int sy = 0; sy <= kMaxSyLog2; ++sy) {
        L.prefix_off[sy] = p->prefix.size();
        long long acc = 0;
        for (int t = L.begin; t < L.end; ++t) {
            p->prefix.push_back((int)acc);
            acc += ((long long)p->tasks[t].n_s + (1 << sy) - 1) >> sy;
            if (acc > 2147483647LL) return fail(JT_ERR_INVALID, "launch too large");
        }
        p->prefix.push_back((int)acc);
        L.blocks[sy] = acc;
    }
    return JT_OK;
}

extern "C" {

int jt_abi_version(void) { return JT_ABI_VERSION; }

const char* jt_last_error_string(void) { return g_err; }

int64_t jt_launch_count(void) { return jt_g_launches.load(std::memory_order_relaxed); }

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
    const int64_t F = w[JT_H_NFACTORS], n_evid = w[JT_H_NEVID], n_evf = w[JT_H_NEVF], n_out = w[JT_H_NOUT];
    const int64_t n_tasks = w[JT_H_NTASKS], n_msgs = w[JT_H_NMSGS], n_launch = w[JT_H_NLAUNCHES];
    const int64_t n_tab = w[JT_H_NTAB];
    const int64_t evf_ptr_n = F > 0 ? F + 1 : 0;
    const int64_t words = JT_H_WORDS + 2 * n_nodes + 2 * F + 2 * n_out + n_evid + evf_ptr_n + 2 * n_evf +
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
    take64(p->fout_off, n_out);
    take64(p->fout_size, n_out);
    for (int64_t k = 0; k < n_out; ++k)
        if (p->fout_off[k] < 0 || p->fout_size[k] <= 0 || p->fout_off[k] + p->fout_size[k] > w[JT_H_FOUT_ENTRIES]) {
            delete p;
            return fail(JT_ERR_INVALID, "malformed plan: output scope %lld", (long long)k);
        }
    take32(p->ev_card, n_evid);
    take32(p->evf_ptr, evf_ptr_n);
    take32(p->evf_var, n_evf);
    take32(p->evf_stride, n_evf);

    const int64_t lik_base = w[JT_H_CLIQUE_ENTRIES] + 3 * w[JT_H_SEP_ENTRIES];
    const int64_t work_entries = lik_base + w[JT_H_LIK_ENTRIES];
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
        t.flags = (int)q[JT_T_FLAGS]; t.pad = 0;
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
        m.fid = (int)q[JT_M_FID]; m.uni = q[JT_M_UNI] ? 1 : 0; m.eoff = 0;
        if (m.off < 0 || m.a_hi < 0 || m.a_lo < 0 || m.b_hi < 0 || m.b_lo < 0 || m.a_hi >= n_tab + 1 ||
            m.a_lo >= n_tab + 1 || m.b_hi >= n_tab + 1 || m.b_lo >= n_tab + 1 || m.fid >= F || m.fid < -2)
            return bad("message descriptor", i);
        // fid -2: a likelihood table in the workspace (soft evidence operand of an init task)
        if (m.fid == -2 && (m.off < lik_base || m.off >= work_entries)) return bad("likelihood operand", i);
    }
    p->launches.resize(n_launch);
    for (int64_t i = 0; i < n_launch; ++i, q += JT_LAUNCH_WORDS) {
        jt_plan::Launch& L = p->launches[i];
        L.phase = (int)q[JT_L_PHASE]; L.begin = (int)q[JT_L_BEGIN]; L.end = (int)q[JT_L_END]; L.level = (int)q[JT_L_LEVEL];
        if (L.phase < JT_PHASE_INIT || L.phase > JT_PHASE_DIST_PRE_INSTANCE || L.begin < 0 || L.begin >= L.end || L.end > n_tasks)
            return bad("launch descriptor", i);
        for (int t = L.begin; t < L.end; ++t)
            if ((p->tasks[t].kind == JT_KIND_INIT) != (L.phase == JT_PHASE_INIT || L.phase == JT_PHASE_INIT_UNIFORM ||
                                                        L.phase == JT_PHASE_INIT_INSTANCE))
                return bad("task kind vs phase", i);
        if (jt_launch_tables(p, L) != JT_OK) return bad("launch too large", i);
    }
    const int32_t* tabp = reinterpret_cast<const int32_t*>(q);
    p->tab.assign(tabp, tabp + n_tab);
    // task ranges of the whole-propagation kernel, general mode, in execution order
    for (int k = 0; k < 3; ++k) {       // 2: collect + distribute on given potentials (compute_beliefs)
        const int main_phase = k == 1 ? JT_PHASE_DIST_MAIN_MESSAGES : JT_PHASE_DIST_MAIN;
        const int marg_phase = k == 1 ? JT_PHASE_MARGINAL_DIRECT : JT_PHASE_MARGINAL;
        auto add = [&](const jt_plan::Launch& L) {
            p->walk_seq[k].push_back(L.begin);
            p->walk_seq[k].push_back(L.end);
            p->walk_items[k] += L.total_items;
        };
        if (k < 2)
            for (const auto& L : p->launches) if (L.phase == JT_PHASE_INIT) add(L);
        for (const auto& L : p->launches) if (L.phase == JT_PHASE_COLLECT) add(L);
        for (const auto& L : p->launches) if (L.phase == JT_PHASE_DIST_PRE || L.phase == main_phase) add(L);
        p->walk_marginal[k] = (int)p->walk_seq[k].size() / 2;
        if (k < 2)
            for (const auto& L : p->launches) if (L.phase == marg_phase) add(L);
    }
    if (jt_dense_build(p) != JT_OK) {
        delete p;
        return JT_ERR_INVALID;
    }
    *out = p;
    return JT_OK;
}

void jt_plan_destroy(jt_plan* p) {
    if (!p) return;
    if (p->device >= 0) {
        jt_dense_free(p);
        cudaFree(p->d_tasks);
        cudaFree(p->d_msgs);
        cudaFree(p->d_tab);
        cudaFree(p->d_prefix);
        cudaFree(p->d_ev);
        cudaFree(p->d_out);
        for (int k = 0; k < 3; ++k) cudaFree(p->d_walk[k]);
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

int jt_plan_dense_count(const jt_plan* p) { return p ? (int)p->dense.size() : 0; }

int jt_plan_dense_get(const jt_plan* p, int index, int64_t* out16) {
    if (!p || !out16 || index < 0 || index >= (int)p->dense.size()) return fail(JT_ERR_INVALID, "bad dense index");
    const DDense& d = p->dense[index];
    int launch = -1;
    for (size_t i = 0; i < p->launches.size(); ++i)
        if (index >= p->launches[i].dense_begin && index < p->launches[i].dense_end) launch = (int)i;
    const int64_t v[16] = {d.task, d.msg, d.n_g, d.n_i, d.K, d.n_q, d.MT, d.n_it, d.n_k4, d.s_of, d.mg, d.mk,
                           d.r_of, d.w_off, d.w_size, launch};
    memcpy(out16, v, sizeof(v));
    return JT_OK;
}

const int32_t* jt_plan_dense_table(const jt_plan* p, int64_t* count) {
    if (count) *count = p ? (int64_t)p->dtab.size() : 0;
    return p && !p->dtab.empty() ? p->dtab.data() : nullptr;
}

int jt_workspace_bytes(const jt_plan* p, int64_t B, int dtype, size_t* out) {
    if (!p || !out || B <= 0 || (dtype != JT_F32 && dtype != JT_F64)) return fail(JT_ERR_INVALID, "bad argument");
    *out = workspace_layout(p, B, dtype).total;
    return JT_OK;
}

int jt_workspace_layout(const jt_plan* p, int64_t B, int dtype, int64_t* out4) {
    if (!p || !out4 || B <= 0 || (dtype != JT_F32 && dtype != JT_F64)) return fail(JT_ERR_INVALID, "bad argument");
    const WorkspaceLayout w = workspace_layout(p, B, dtype);
    out4[0] = (int64_t)w.fbase_off;
    out4[1] = (int64_t)w.err_off;
    out4[2] = (int64_t)w.uni_off;
    out4[3] = (int64_t)w.total;
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
    std::vector<long long> outs(p->fout_off.begin(), p->fout_off.end());
    outs.insert(outs.end(), p->fout_size.begin(), p->fout_size.end());
    JT_CUDA(up(&p->d_out, outs));
    for (int k = 0; k < 3; ++k) JT_CUDA(up(&p->d_walk[k], p->walk_seq[k]));
    if (jt_dense_upload(p) != JT_OK) return JT_ERR_CUDA;
    p->device = dev;
    return JT_OK;
}

int jt_init(jt_plan* p, const void* factor_tables, int factors_batched, const int32_t* evidence, int64_t B,
            int dtype, void* workspace, int flags, void* stream_) {
    NvtxRange range("jt_init");
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    if ((flags & JT_UNIFORM) && factors_batched)
        return fail(JT_ERR_INVALID, "JT_UNIFORM requires factor tables shared by the batch");
    if (p->hdr[JT_H_NFACTORS] == 0) return fail(JT_ERR_INVALID, "plan has no factors: nothing to initialise");
    if (!factor_tables) return fail(JT_ERR_INVALID, "factor_tables is null");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const WorkspaceLayout wl = workspace_layout(p, B, dtype);
    const int n_evid = (int)p->hdr[JT_H_NEVID];
    const int vec = pick_vec(B, dtype);
    KArgs a = base_args(p, B, workspace, vec);
    a.fin = factor_tables;
    a.fin_batched = factors_batched ? 1 : 0;
    a.flags = flags;
    if (!factors_batched) {
        long long fin_end = 0;
        for (size_t f = 0; f < p->fin_off.size(); ++f) fin_end = std::max<long long>(fin_end, p->fin_off[f] + p->fin_size[f]);
        if (fin_end < 2147483647LL) a.flags |= JT_X_FIN32;
    }
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
        jt_g_launches.fetch_add(1, std::memory_order_relaxed);
        JT_CUDA(cudaGetLastError());
        a.fbase = fbase;
    }
    if (!uniform_mode(p, flags)) return run_phase(p, JT_PHASE_INIT, a, dtype, vec, stream);
    // uniform mode: potentials no evidence touches are written once, the others per instance
    if (!(flags & JT_UNIFORM_VALID)) {
        rc = run_phase_uniform(p, JT_PHASE_INIT_UNIFORM, a, dtype, uniform_ws(p, B, dtype, workspace), stream);
        if (rc != JT_OK) return rc;
    }
    return run_phase(p, JT_PHASE_INIT_INSTANCE, a, dtype, vec, stream);
}

int jt_collect(jt_plan* p, int64_t B, int dtype, void* workspace, int flags, void* stream_) {
    NvtxRange range("jt_collect");
    g_side.used = 0;                      // the event pool of this thread starts over with every stage call
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int vec = pick_vec(B, dtype);
    KArgs a = base_args(p, B, workspace, vec);
    a.flags = flags;
    if (!uniform_mode(p, flags)) return run_phase(p, JT_PHASE_COLLECT, a, dtype, vec, stream);
    // uniform mode: evidence-free subtrees are collected once (B = 1), then the rest per instance
    void* uni = uniform_ws(p, B, dtype, workspace);
    void* w_region = dense_region(p, B, dtype, workspace, flags);
    if (!(flags & JT_UNIFORM_VALID)) {
        rc = run_phase_uniform(p, JT_PHASE_COLLECT_UNIFORM, a, dtype, uni, stream);
        if (rc != JT_OK) return rc;
        if (w_region) {      // W blocks of the collect contractions, totals of the scalar tasks
            rc = jt_dense_prepare(p, 0, dtype, uni, w_region, stream);
            if (rc == JT_OK) rc = run_phase_uniform(p, JT_PHASE_X_SCALAR0, a, dtype, uni, stream);
            if (rc != JT_OK) return rc;
        }
    }
    // The uniform part of the distribute pass depends on the uniform collect only: it runs now, on
    // a stream of its own beside the instance collect (a chain of one tiny launch per level that
    // would otherwise sit in front of jt_distribute), and jt_distribute finds it done.
    cudaEvent_t uni_done = nullptr;
    g_side.uni_dist_plan = g_side.uni_dist_ws = nullptr;
    if (!(flags & JT_UNIFORM_VALID) && uniform_overlap_enabled()) {
        cudaEvent_t ev;
        JT_CUDA(side_acquire(nullptr));
        JT_CUDA(side_event(&ev));
        JT_CUDA(cudaEventRecord(ev, stream));
        JT_CUDA(cudaStreamWaitEvent(g_side.uni, ev, 0));
        rc = uniform_distribute(p, a, dtype, uni, w_region, g_side.uni);
        if (rc != JT_OK) return rc;
        JT_CUDA(side_event(&uni_done));
        JT_CUDA(cudaEventRecord(uni_done, g_side.uni));
    }
    a.uni = uni;
    a.uniform = 1;
    rc = run_phase(p, JT_PHASE_COLLECT_INSTANCE, a, dtype, vec, stream, w_region);
    if (uni_done) {                     // join: the stage is complete when the caller's stream is
        JT_CUDA(cudaStreamWaitEvent(stream, uni_done, 0));
        if (rc == JT_OK) {
            g_side.uni_dist_plan = p;
            g_side.uni_dist_ws = workspace;
        }
    }
    return rc;
}

int jt_distribute(jt_plan* p, int64_t B, int dtype, void* workspace, int flags, void* stream_) {
    NvtxRange range("jt_distribute");
    g_side.used = 0;                      // the event pool of this thread starts over with every stage call
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int vec = pick_vec(B, dtype);
    KArgs a = base_args(p, B, workspace, vec);
    a.flags = flags;
    int pre_phase = JT_PHASE_DIST_PRE;
    void* w_region = nullptr;
    bool split = false;
    if (uniform_mode(p, flags)) {
        split = jt_beta_enabled(p, B, dtype, flags) && vec * dtype_size(dtype) == 16;
        // down-messages with an evidence-free source side: once, top level first (B = 1)
        void* uni = uniform_ws(p, B, dtype, workspace);
        w_region = dense_region(p, B, dtype, workspace, flags);
        const bool done_by_collect = g_side.uni_dist_plan == p && g_side.uni_dist_ws == workspace;
        g_side.uni_dist_plan = g_side.uni_dist_ws = nullptr;
        if (!(flags & JT_UNIFORM_VALID) && !done_by_collect) {
            rc = uniform_distribute(p, a, dtype, uni, w_region, stream);
            if (rc != JT_OK) return rc;
        }
        a.uni = uni;
        a.uniform = 1;
        pre_phase = JT_PHASE_DIST_PRE_INSTANCE;
    }
    // per level: the tasks that only read psi_C, then the task that overwrites it with beta_C
    // (JT_NO_BELIEFS: only the message-sending tasks, and the kernels skip the belief stores)
    const int main_phase = (flags & JT_NO_BELIEFS) ? JT_PHASE_DIST_MAIN_MESSAGES : JT_PHASE_DIST_MAIN;
    cudaStream_t side = nullptr;
    bool forked = false;
    if (split && beta_overlap_enabled()) JT_CUDA(side_acquire(&side));
    const jt_plan::Launch* pending_pre = nullptr;       // the DIST_PRE launch of the current level, not yet enqueued
    auto run_pre = [&](const jt_plan::Launch& Q, cudaStream_t on, int aux_set) {
        return run_launch(p, Q, a, dtype, vec, w_region, on, false, nullptr, aux_set);
    };
    for (const auto& L : p->launches) {
        if (L.phase != pre_phase && L.phase != main_phase) continue;
        if (L.phase == pre_phase) {
            if (pending_pre) {                                    // a level without a main launch
                rc = run_pre(*pending_pre, stream, 0);
                if (rc != JT_OK) return rc;
            }
            pending_pre = &L;
            continue;
        }
        // the two launches of a level side by side when neither touches what the other reads
        const bool pair = pending_pre && pending_pre->level == L.level && L.pre_independent && uniform_mode(p, flags) &&
                          level_fork_enabled();
        cudaEvent_t pre_done = nullptr;
        if (pending_pre && !pair) {
            rc = run_pre(*pending_pre, stream, 0);
            if (rc != JT_OK) return rc;
        } else if (pair) {
            cudaEvent_t ev;
            JT_CUDA(side_acquire(nullptr));
            JT_CUDA(side_event(&ev));
            JT_CUDA(cudaEventRecord(ev, stream));
            JT_CUDA(cudaStreamWaitEvent(g_side.pre, ev, 0));
            rc = run_pre(*pending_pre, g_side.pre, 1);
            if (rc != JT_OK) return rc;
            JT_CUDA(side_event(&pre_done));
            JT_CUDA(cudaEventRecord(pre_done, g_side.pre));
        }
        pending_pre = nullptr;
        cudaStream_t beta_stream = nullptr;
        if (side && L.phase == JT_PHASE_DIST_MAIN && L.beta_n > 0) {
            // the beliefs of this level need the messages of the levels above (and the collect
            // pass): everything enqueued so far
            cudaEvent_t ev;
            JT_CUDA(side_event(&ev));
            JT_CUDA(cudaEventRecord(ev, stream));
            JT_CUDA(cudaStreamWaitEvent(side, ev, 0));
            beta_stream = side;
            forked = true;
        }
        rc = run_launch(p, L, a, dtype, vec, w_region, stream, split, beta_stream);
        if (rc != JT_OK) return rc;
        if (pre_done) JT_CUDA(cudaStreamWaitEvent(stream, pre_done, 0));
    }
    if (pending_pre) {
        rc = run_pre(*pending_pre, stream, 0);
        if (rc != JT_OK) return rc;
    }
    if (forked) {                               // join: the stage is complete when the caller's stream is
        cudaEvent_t ev;
        JT_CUDA(side_event(&ev));
        JT_CUDA(cudaEventRecord(ev, side));
        JT_CUDA(cudaStreamWaitEvent(stream, ev, 0));
    }
    return JT_OK;
}

int jt_marginal(jt_plan* p, int64_t B, int dtype, void* workspace, void* factor_out, int flags, void* stream) {
    NvtxRange range("jt_marginal");
    g_side.used = 0;
    int rc = check_common(p, B, dtype, workspace);
    if (rc != JT_OK) return rc;
    if (!factor_out) return fail(JT_ERR_INVALID, "factor_out is null");
    const int vec = pick_vec(B, dtype);
    KArgs a = base_args(p, B, workspace, vec);
    a.fout = factor_out;
    a.flags = flags;
    if (!(flags & JT_NO_BELIEFS)) return run_phase(p, JT_PHASE_MARGINAL, a, dtype, vec, static_cast<cudaStream_t>(stream));
    // outputs straight from psi_C and the incoming messages (the beliefs were not written)
    void* w_region = nullptr;
    if (uniform_mode(p, flags)) {
        a.uni = uniform_ws(p, B, dtype, workspace);
        a.uniform = 1;
        w_region = dense_region(p, B, dtype, workspace, flags);     // prepared by jt_distribute
    }
    return run_phase(p, JT_PHASE_MARGINAL_DIRECT, a, dtype, vec, static_cast<cudaStream_t>(stream), w_region);
}

namespace {

// A few instances of a small tree: the level-ordered launches are pure launch latency, so the
// whole propagation runs as one launch, one CTA per instance (jt_walk_kernel).  JT_DISABLE_WALK=1
// keeps the per-level launches (A-B timing, and the tests of the per-level kernels at tiny B).
constexpr int64_t kWalkMaxBatch = 16;
constexpr long long kWalkMaxItems = 1 << 16;

bool walk_enabled() {
    static const int on = [] {
        const char* e = getenv("JT_DISABLE_WALK");
        return (e && e[0] == '1') ? 0 : 1;
    }();
    return on == 1;
}

}  // namespace

int jt_plan_single_launch(const jt_plan* p, int64_t B, int flags) {
    if (!p || B <= 0 || B > kWalkMaxBatch || !walk_enabled() || p->hdr[JT_H_NFACTORS] <= 0) return 0;
    const int k = (flags & JT_NO_BELIEFS) ? 1 : 0;
    return !p->walk_seq[k].empty() && p->walk_items[k] <= kWalkMaxItems ? 1 : 0;
}

int jt_propagate(jt_plan* p, const void* factor_tables, int factors_batched, const int32_t* evidence, int64_t B,
                 int dtype, void* workspace, void* factor_out, int flags, void* stream) {
    if (jt_plan_single_launch(p, B, flags)) {
        const int k = (flags & JT_NO_BELIEFS) ? 1 : 0;
        {
            int rc = check_common(p, B, dtype, workspace);
            if (rc != JT_OK) return rc;
            if (!factor_tables) return fail(JT_ERR_INVALID, "factor_tables is null");
            const bool marginal = !(flags & JT_SKIP_MARGINAL);
            if (marginal && !factor_out) return fail(JT_ERR_INVALID, "factor_out is null");
            const int n_evid = (int)p->hdr[JT_H_NEVID];
            if (n_evid > 0 && !evidence)
                return fail(JT_ERR_INVALID, "plan has %d evidence variables but evidence is null", n_evid);
            if (n_evid > 0 && factors_batched)
                return fail(JT_ERR_INVALID, "per-instance factor tables cannot be combined with evidence indices");
            const WorkspaceLayout wl = workspace_layout(p, B, dtype);
            KArgs a = base_args(p, B, workspace, 1);
            a.tasks = p->d_tasks;
            a.fin = factor_tables;
            a.fin_batched = factors_batched ? 1 : 0;
            a.fout = factor_out;
            a.flags = flags;
            jt_walk_args w;
            memset(&w, 0, sizeof(w));
            w.seq = p->d_walk[k];
            w.lik_base = p->hdr[JT_H_CLIQUE_ENTRIES] + 3 * p->hdr[JT_H_SEP_ENTRIES];
            w.lik_entries = p->hdr[JT_H_LIK_ENTRIES];
            w.work_entries = w.lik_base + w.lik_entries;
            w.n_tasks = (int)p->tasks.size();
            w.n_msgs = (int)p->msgs.size();
            w.n_tab = (int)p->tab.size();
            w.n_seq = marginal ? (int)p->walk_seq[k].size() / 2 : p->walk_marginal[k];
            if (n_evid > 0) {
                a.fbase = reinterpret_cast<int*>(static_cast<char*>(workspace) + wl.fbase_off);
                w.evidence = evidence;
                w.n_evid = n_evid;
                w.ev_card = p->d_ev;
                w.evf_ptr = w.ev_card + p->ev_card.size();
                w.evf_var = w.evf_ptr + p->evf_ptr.size();
                w.evf_stride = w.evf_var + p->evf_var.size();
                w.n_factors = (int)p->hdr[JT_H_NFACTORS];
                w.errors = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + wl.err_off);
            }
            return launchers(flags)->walk(a, w, dtype, static_cast<cudaStream_t>(stream));
        }
    }
    // shared factor tables: potentials and messages no evidence reaches are computed once
    if (factors_batched) flags &= ~JT_UNIFORM;
    else if (!(flags & JT_NO_UNIFORM)) flags |= JT_UNIFORM;
    int rc = jt_init(p, factor_tables, factors_batched, evidence, B, dtype, workspace, flags, stream);
    if (rc != JT_OK) return rc;
    rc = jt_collect(p, B, dtype, workspace, flags, stream);
    if (rc != JT_OK) return rc;
    rc = jt_distribute(p, B, dtype, workspace, flags, stream);
    if (rc != JT_OK) return rc;
    if (flags & JT_SKIP_MARGINAL) return JT_OK;
    return jt_marginal(p, B, dtype, workspace, factor_out, flags, stream);
}

int jt_propagate_host(jt_plan* p, const void* host_factors, size_t factor_bytes, const int32_t* host_evidence,
                      int64_t B, int dtype, void* dev_factors, int32_t* dev_evidence, void* workspace,
                      void* dev_out, void* host_out, size_t out_bytes, int flags, void* stream_) {
    if (!p || !host_factors || !dev_factors || !host_out || !dev_out) return fail(JT_ERR_INVALID, "null argument");
    if (B <= 0) return fail(JT_ERR_INVALID, "batch size must be positive");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n_evid = (size_t)p->hdr[JT_H_NEVID];
    if (n_evid > 0 && (!host_evidence || !dev_evidence))
        return fail(JT_ERR_INVALID, "plan has %zu evidence variables but evidence is null", n_evid);
    JT_CUDA(cudaMemcpyAsync(dev_factors, host_factors, factor_bytes, cudaMemcpyHostToDevice, stream));
    if (n_evid > 0)
        JT_CUDA(cudaMemcpyAsync(dev_evidence, host_evidence, (size_t)B * n_evid * sizeof(int32_t),
                                cudaMemcpyHostToDevice, stream));
    int rc = jt_propagate(p, dev_factors, 0, n_evid > 0 ? dev_evidence : nullptr, B, dtype, workspace, dev_out, flags,
                          stream_);
    if (rc != JT_OK) return rc;
    JT_CUDA(cudaMemcpyAsync(host_out, dev_out, out_bytes, cudaMemcpyDeviceToHost, stream));
    JT_CUDA(cudaStreamSynchronize(stream));
    return JT_OK;
}

int jt_beliefs_host(jt_plan* p, const void* host_potentials, int dtype, void* workspace, void* host_beliefs,
                    int flags, void* stream_) {
    int rc = check_common(p, 1, dtype, workspace);
    if (rc != JT_OK) return rc;
    if (!host_potentials || !host_beliefs) return fail(JT_ERR_INVALID, "null argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t w = dtype_size(dtype);
    const size_t n_c = (size_t)p->hdr[JT_H_CLIQUE_ENTRIES], n_s = (size_t)p->hdr[JT_H_SEP_ENTRIES];
    flags = (flags & JT_SR_MASK) | JT_SEP_BELIEFS;
    // B = 1: the [entries][1] block is a plain vector, cliques first, then separator beliefs
    JT_CUDA(cudaMemcpyAsync(workspace, host_potentials, n_c * w, cudaMemcpyHostToDevice, stream));
    if (walk_enabled() && !p->walk_seq[2].empty() && p->walk_items[2] <= kWalkMaxItems) {
        KArgs a = base_args(p, 1, workspace, 1);
        a.tasks = p->d_tasks;
        a.flags = flags;
        jt_walk_args wa;
        memset(&wa, 0, sizeof(wa));
        wa.seq = p->d_walk[2];
        wa.n_seq = (int)p->walk_seq[2].size() / 2;
        wa.lik_base = p->hdr[JT_H_CLIQUE_ENTRIES] + 3 * p->hdr[JT_H_SEP_ENTRIES];
        wa.work_entries = wa.lik_base + p->hdr[JT_H_LIK_ENTRIES];
        wa.n_tasks = (int)p->tasks.size();
        wa.n_msgs = (int)p->msgs.size();
        wa.n_tab = (int)p->tab.size();
        wa.preload = (long long)n_c;           // the potentials just copied in
        rc = launchers(flags)->walk(a, wa, dtype, stream);
    } else {
        rc = jt_collect(p, 1, dtype, workspace, flags, stream_);
        if (rc == JT_OK) rc = jt_distribute(p, 1, dtype, workspace, flags, stream_);
    }
    if (rc != JT_OK) return rc;
    JT_CUDA(cudaMemcpyAsync(host_beliefs, workspace, (n_c + n_s) * w, cudaMemcpyDeviceToHost, stream));
    JT_CUDA(cudaStreamSynchronize(stream));
    return JT_OK;
}

int jt_normalize(jt_plan* p, int64_t B, int dtype, void* factor_out, void* logz, int flags, void* stream) {
    int rc = check_common(p, B, dtype, factor_out);
    if (rc != JT_OK) return rc;
    const int n_out = (int)p->fout_off.size();
    if (n_out == 0) return JT_OK;
    if (n_out > 65535) return fail(JT_ERR_INVALID, "too many output scopes for one launch");
    if ((flags & JT_LOGZ_ONLY) && !logz) return fail(JT_ERR_INVALID, "JT_LOGZ_ONLY needs a logz buffer");
    return launchers(flags)->normalize(p, B, dtype, factor_out, logz, (flags & JT_LOGZ_ONLY) ? 0 : 1,
                                       static_cast<cudaStream_t>(stream));
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
    jt_g_launches.fetch_add(1, std::memory_order_relaxed);
    JT_CUDA(cudaGetLastError());
    return JT_OK;
}

int jt_contract(const void* const* ops, int n_ops, const int32_t* tables, int64_t n_tab, const int32_t* maps,
                int64_t n_s, int64_t n_r, int64_t n_slo, int64_t n_rlo, int64_t B, int dtype, void* out,
                int flags, void* stream_) {
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
        a.flags = flags;
        rc = launchers(flags)->contract(a, blocks, gy, dtype, vec, stream);
    }
    cudaFreeAsync(dev, stream);
    return rc;
}

}  // extern "C"
