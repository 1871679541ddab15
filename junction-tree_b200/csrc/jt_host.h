// Host-side types shared by the translation units of libjt_b200: the parsed plan, the kernel
// argument block, error reporting and the per-semiring launcher table.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../include/jt_b200.h"

// ------------------------------------------------------------------------------------------
// error handling (defined in jt_abi.cu)

int jt_fail(int code, const char* fmt, ...);
extern std::atomic<int64_t> jt_g_launches;

#define JT_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return jt_fail(JT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));      \
    } while (0)

// ------------------------------------------------------------------------------------------
// device-side descriptors

struct DTask {
    long long src, out, beta, bel, own;  // entry offsets, -1 = absent
    int n_s, n_r, n_slo, n_rlo;
    int src_shi, src_slo, src_rhi, src_rlo;
    int rmsg_begin, rmsg_end, smsg_begin, smsg_end;
    int kind, out_space;
    int flags, pad;       // JT_TF_* (honoured in uniform mode only)
};

struct DMsg {
    long long off;    // entry offset (multiplied by B on the device)
    long long eoff;   // element offset added as is (jt_contract operands; 0 inside a plan)
    int a_hi, a_lo, b_hi, b_lo;
    int fid;          // init: factor index
    int uni;          // message buffer is uniform (read from the uniform workspace in uniform mode)
};

struct KArgs {
    const DTask* tasks;   // first task of this launch
    const DMsg* msgs;     // all messages of the plan
    const int* tab;       // all index tables
    const int* prefix;    // [n_tasks + 1] first block of each task for this launch / tile shape
    void* work;
    const void* uni;      // uniform workspace (same entry offsets, B = 1), or null
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
    int uniform;          // honour the uniform-operand flags of tasks and messages
};

constexpr int kThreads = 256;
constexpr int kMaxSyLog2 = 12;    // largest chunk of s per block: 4096
constexpr int kItemLog2Max = 24;
constexpr int kTmaMaxRows = 8;    // operands per task supported by the TMA kernel (src + messages + own)
constexpr int kNumSemirings = 4;

// ------------------------------------------------------------------------------------------
// the parsed plan

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
        // TMA kernel: per-task chunks sized for ~2^j (s, r) items per CTA, j = 0..kItemLog2Max;
        // layout per j: [n_tasks + 1] block prefix, [n_tasks] log2 chunk
        size_t item_prefix_off[kItemLog2Max + 1];
        long long item_blocks[kItemLog2Max + 1];
        long long total_items;
        bool tma_ok;          // every task fits the TMA kernel's stage (rows per stage <= kTmaMaxRows)
        int min_nr;           // smallest n_r of the launch
        int max_nr;           // largest n_r of the launch
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
    long long* d_out = nullptr;   // fout_off | fout_size
    // whole-propagation kernel: task ranges in execution order, general mode; [0] with clique
    // beliefs (INIT, COLLECT, DIST_PRE/DIST_MAIN, MARGINAL), [1] without (.., DIST_MAIN_MESSAGES,
    // MARGINAL_DIRECT); walk_marginal[k] = number of ranges before the marginal stage
    // [2] = COLLECT, DIST_PRE/DIST_MAIN only (collect + distribute on given potentials)
    std::vector<int> walk_seq[3];
    int walk_marginal[3] = {0, 0, 0};
    long long walk_items[3] = {0, 0, 0};  // sum of n_s * n_r over the tasks of each sequence
    int* d_walk[3] = {nullptr, nullptr, nullptr};
    // > 48 KB dynamic shared memory opted in per [semiring][f32|f64][VPT-1]
    mutable bool tma_attr_set[kNumSemirings][2][2] = {};
};

inline bool jt_is_init_phase(int phase) {
    return phase == JT_PHASE_INIT || phase == JT_PHASE_INIT_UNIFORM || phase == JT_PHASE_INIT_INSTANCE;
}

// ------------------------------------------------------------------------------------------
// Per-semiring launchers: one translation unit per semiring (jt_sr_*.cu) instantiates the
// kernels of jt_kernels.cuh and exports this table; jt_abi.cu picks one by the JT_SR_* flag.

// arguments of the whole-propagation kernel besides KArgs (device pointers)
struct jt_walk_args {
    const int* seq;        // [n_seq][2] task ranges in execution order
    int n_seq;
    const int* evidence;   // [B][n_evid] or null
    int n_evid;
    const int* ev_card;
    const int* evf_ptr;
    const int* evf_var;
    const int* evf_stride;
    int n_factors;
    unsigned long long* errors;
    long long work_entries;            // entries of the [entries][B] block of the workspace
    long long lik_base, lik_entries;   // its likelihood region
    int n_tasks, n_msgs, n_tab;        // sizes of the plan's descriptor arrays
    long long preload;                 // leading entries of the column that already hold inputs (compute_beliefs)
};

struct jt_sr_launchers {
    // all tasks of one launch of the plan (init or projection), kernel chosen by batch shape
    int (*dispatch)(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, int dtype, int vec,
                    cudaStream_t stream);
    // one projection task outside a plan (jt_contract), LDG kernel
    int (*contract)(const KArgs& a, long long blocks, long long gy, int dtype, int vec, cudaStream_t stream);
    // output stage
    int (*normalize)(const jt_plan* p, int64_t B, int dtype, void* factor_out, void* logz, cudaStream_t stream);
    // init + collect + distribute (+ marginal) of a few instances of a small tree in one launch
    int (*walk)(const KArgs& a, const jt_walk_args& w, int dtype, cudaStream_t stream);
};

const jt_sr_launchers* jt_sr_sum_product();
const jt_sr_launchers* jt_sr_max_product();
const jt_sr_launchers* jt_sr_log_sum_exp();
const jt_sr_launchers* jt_sr_max_sum();
