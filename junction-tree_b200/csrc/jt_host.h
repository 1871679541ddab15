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

// Dense contraction of a batch-shared potential with one per-instance message (uniform mode,
// sum-product): derived from a projection task when the plan is loaded (jt_dense.cu).
//   out[s_of[g][i]][b] = sum_k W[g][i][k] * msg[mg[g] + mk[k]][b]
// W = the uniform operands (potential, uniform messages) summed over the axes the message does
// not see; it lives in the W region of the workspace in MMA-fragment order (see jt_dense.cu).
struct DDense {
    int task;             // index into the plan's tasks: out / bel / own / flags / s-only messages
    int msg;              // index of the per-instance r-message (DMsg)
    int n_g, n_i, K;      // groups (distinct message rows bases), output rows per group, contraction length
    int MT, n_it, n_k4;   // 8-row m-tiles per i-tile, i-tiles per group, k-steps of 4
    int n_q;              // clique entries summed into one W entry
    int s_of, mg, mk;     // offsets into the dense int table: s_of[n_g * n_i], mg[n_g], mk[K]
    int r_of;             // prep: r_of[K * n_q] = the r values summed into column k, ascending
    int n_chunks;         // stages per unit: ceil(n_k4 / 4)
    int ups;              // units per stage (short contractions, K <= 8, share a stage), else 1
    long long w_off;      // element offset of the task's W block inside the W region
    long long w_size;     // elements
};

// internal bits (never in a plan blob or in the public flags)
constexpr int JT_TF_BETA_SPLIT = 0x100;   // DTask::flags: in uniform mode the clique belief of this task is written by jt_beta_kernel
constexpr int JT_X_BETA_SPLIT = 0x10000;  // KArgs::flags: ... and this launch runs that way
constexpr int JT_X_FIN32 = 0x20000;       // KArgs::flags (init): the shared factor tables hold < 2^31 entries, gathers index them in 32 bits
// internal phases (launches derived at plan load, never in a blob): the totals of the scalar tasks,
// ordinary B = 1 projection launches in the uniform workspace after the uniform collect / distribute
constexpr int JT_PHASE_X_SCALAR0 = 64, JT_PHASE_X_SCALAR1 = 65;
constexpr int kBetaRows = 3;
constexpr int kScalarRows = 32;           // output rows per block of jt_scalar_kernel              // per-instance row operands jt_beta_kernel multiplies per entry
constexpr int kBetaChMin = 8, kBetaChMax = 14;   // log2 of the items per block of a jt_beta_kernel launch

constexpr int kThreads = 256;
constexpr int kDenseJMax = 16;    // dense launches: per-task units per CTA for ~2^j stages per CTA
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
        // reduced task sets (variants 1 and 2 of dispatch).  Launches other than DIST_MAIN: [0] =
        // without the tasks that run as dense contractions.  DIST_MAIN (uniform mode, beliefs
        // written by jt_beta_kernel): [0] = without the tasks that only write a belief, [1] = also
        // without the dense contractions.
        size_t item_prefix_off_v[2][kItemLog2Max + 1];
        long long item_blocks_v[2][kItemLog2Max + 1];
        long long total_items_v[2];
        // DIST_MAIN: tasks whose clique belief jt_beta_kernel writes ([n] task ids, then per
        // chunk size 2^kBetaChMin .. 2^kBetaChMax a block prefix [n + 1]), in jt_plan::prefix
        size_t beta_off = 0;
        int beta_n = 0;
        long long beta_items = 0;
        // scalar tasks of this launch (uniform mode): [n] task ids, [n] entry of their totals in the
        // uniform workspace, [n + 1] block prefix (kScalarRows values of s per block), in jt_plan::prefix
        size_t scalar_off = 0;
        int scalar_n = 0;
        // DIST_MAIN / DIST_MAIN_MESSAGES: no task of this launch overwrites (in place, with its clique
        // belief) a per-instance potential that a task of the level's DIST_PRE_INSTANCE launch reads,
        // so the two launches of the level may run side by side
        bool pre_independent = false;
        // dense contractions of this launch: range in jt_plan::dense, block prefix per j
        // (layout per j: [n + 1] first block of each task, [n] units per CTA)
        int dense_begin = 0, dense_end = 0;
        size_t dense_prefix_off[kDenseJMax + 1];
        long long dense_blocks[kDenseJMax + 1];
        long long dense_stages;   // sum over tasks of units * stages per unit
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
    // dense contractions (jt_dense.cu): descriptors sorted by launch, their int tables, the size
    // of the W region in elements; [0] tasks prepared after the uniform collect, [1] after the
    // uniform distribute (ranges in `dense`, W element ranges)
    bool accel = false;                       // the plan has dense contractions, scalar tasks or split beliefs
    long long scalar_entries = 0;             // totals of the scalar tasks, after the entries of the uniform workspace
    std::vector<DDense> dense;
    std::vector<int> dtab;
    long long dense_w_entries = 0;
    int dense_group_end[2] = {0, 0};          // dense[0 .. end[0]) collect, [end[0] .. end[1]) distribute + marginal
    std::vector<int> dense_prep_prefix[2];    // block prefix of the prep launches
    DDense* d_dense = nullptr;
    int* d_dtab = nullptr;
    int* d_dense_prep[2] = {nullptr, nullptr};
    mutable bool dense_attr_set = false;
};

// jt_abi.cu: the tables of a launch (prefixes per tile shape, totals, kernel eligibility)
int jt_launch_tables(jt_plan* p, jt_plan::Launch& L);
// jt_abi.cu: block prefix tables of a TMA projection launch (tasks flagged in `skip` get no blocks)
int jt_build_item_prefix(jt_plan* p, const jt_plan::Launch& L, const char* skip, size_t* off_out, long long* blocks_out);

// jt_dense.cu
int jt_dense_build(jt_plan* p);                       // derive the dense contractions (plan load)
int jt_dense_upload(jt_plan* p);
void jt_dense_free(jt_plan* p);
bool jt_dense_enabled(const jt_plan* p, int64_t B, int dtype, int flags);
bool jt_tma_path(int64_t B, int dtype);               // batch shape that runs the projection tasks in jt_project_tma_kernel
bool jt_beta_enabled(const jt_plan* p, int64_t B, int dtype, int flags);
// W blocks of group `which` (0 collect, 1 distribute + marginal) from the uniform workspace
int jt_dense_prepare(const jt_plan* p, int which, int dtype, const void* uni_ws, void* w_region, cudaStream_t stream);
// all dense contractions of one launch
int jt_dense_launch(const jt_plan* p, const jt_plan::Launch& L, void* work, const void* uni_ws, const void* w_region,
                    void* fout, int64_t B, int dtype, int flags, cudaStream_t stream);

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
    // variant 1, 2: reduced task sets (jt_plan::Launch::item_prefix_off_v), the rest of the launch
    // runs in jt_dense_kernel / jt_beta_kernel
    int (*dispatch)(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, int dtype, int vec, int variant,
                    cudaStream_t stream);
    // clique beliefs of the launch's JT_TF_BETA_SPLIT tasks (uniform mode): beta = scalars x rows
    int (*beta)(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, int dtype, int vec, cudaStream_t stream);
    // scalar tasks of the launch (uniform mode): out[s] = total[s] x the per-instance s-only rows
    int (*scalar)(const jt_plan* p, const jt_plan::Launch& L, const KArgs& a, int dtype, int vec, cudaStream_t stream);
    // one projection task outside a plan (jt_contract), LDG kernel
    int (*contract)(const KArgs& a, long long blocks, long long gy, int dtype, int vec, cudaStream_t stream);
    // output stage
    int (*normalize)(const jt_plan* p, int64_t B, int dtype, void* factor_out, void* logz, int normalize,
                     cudaStream_t stream);
    // init + collect + distribute (+ marginal) of a few instances of a small tree in one launch
    int (*walk)(const KArgs& a, const jt_walk_args& w, int dtype, cudaStream_t stream);
};

const jt_sr_launchers* jt_sr_sum_product();
const jt_sr_launchers* jt_sr_max_product();
const jt_sr_launchers* jt_sr_log_sum_exp();
const jt_sr_launchers* jt_sr_max_sum();
