/*
 * jt_b200.h -- C ABI of the B200 (sm_100a) sum-product propagation library, libjt_b200.so.
 *
 * The reference (jluttine/junction-tree, pure Python + NumPy) has no native boundary; its
 * boundary for this path is the Python call surface.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference checkout):
 *
 *   jt_init        CliqueGraph.evaluate                junctiontree/junctiontree.py:203-226
 *                  + evidence slicing apply_evidence   junctiontree/computation.py:11-34
 *   jt_collect     compute_beliefs / get_message       junctiontree/computation.py:47-96
 *   jt_distribute  compute_beliefs / send_message      junctiontree/computation.py:140-224
 *                  (remove_message :99-136 is replaced by a division-free exclude-one product)
 *   jt_marginal    CliqueGraph.marginalize             junctiontree/junctiontree.py:229-274
 *   jt_propagate   JunctionTree.propagate              junctiontree/junctiontree.py:297-331
 *   jt_propagate_host  the same with host buffers (copies, propagate, synchronise in one call)
 *   jt_beliefs_host    compute_beliefs with host buffers  junctiontree/computation.py:37-246
 *   jt_contract    SumProduct.einsum                   junctiontree/sum_product.py:14-35
 *   jt_normalize   the partition function the reference discards  junctiontree/computation.py:90-96
 *   jt_triangulate / jt_junction_tree   find_triangulation / construct_junction_tree
 *                                                      junctiontree/construction.py:176-353, 522-601
 *   jt_plan_build  the per-call Python bookkeeping of the passes above, compiled once
 *
 * Conventions: plain C types only; every function returns JT_OK or an error code and never
 * throws; jt_last_error_string() describes the last error of the calling thread.  The caller
 * owns every device buffer.  All device work is enqueued on the caller's stream and the library
 * never synchronises (except jt_evidence_errors), so calls can be captured into a CUDA graph
 * after jt_plan_upload.  A plan is immutable after jt_plan_upload and may be shared; a
 * workspace belongs to one stream at a time.  There is no CPU fallback: without a CUDA device
 * the compute entry points fail.
 *
 * Device memory layout: batch-innermost.  A node (clique or separator) with n entries and a
 * batch of B independent propagations is stored as [n][B]; element (entry e, instance b) of the
 * node at entry offset `off` lives at workspace[(off + e) * B + b].  Workspace regions, in
 * entries: [ cliques | separator beliefs | up-messages | down-messages | likelihoods ], i.e. the
 * first (clique_entries + sep_entries) rows are the reference's node order `maxcliques +
 * separators` (junctiontree.py:317-323).  The likelihood region (soft evidence; JT_H_LIK_ENTRIES
 * rows starting at entry clique_entries + 3 * sep_entries, empty for most plans) holds one
 * [size_v][B] table per soft-evidence variable, in the order given to jt_plan_build; the caller
 * fills it before jt_init, which multiplies each table into a clique containing the variable --
 * exactly one more single-variable factor per instance.  Entry offsets of nodes come from jt_plan_node_range().  After the
 * [entries][B] block the workspace holds the per-instance factor offsets (int32 [F][B]), an
 * error counter and the *uniform workspace*: one more copy of the same regions with B = 1
 * (byte offsets: jt_workspace_layout).
 *
 * Uniform mode (JT_UNIFORM): with factor tables shared by the batch, a clique potential whose
 * factors contain no observed variable, and an up-message of a subtree that contains none, are
 * the same for every instance.  They are computed once into the uniform workspace and the batch
 * kernels broadcast them instead of streaming [n][B] rows; every belief and every down-message
 * is still computed and stored per instance, so the results are identical.  The same holds for
 * a down-message whose whole source side (clique potential, message from above, up-messages of
 * the other children) is evidence-free: it is also computed once (JT_PHASE_DIST_UNIFORM) and its
 * consumers read the uniform copy.  jt_propagate enables uniform mode automatically when
 * factors_batched = 0.
 *
 * Plan blob (produced by junctiontree/schedule.py, Plan.to_blob): little-endian int64 words
 *   header[JT_H_WORDS], node_off[n_nodes], node_size[n_nodes], fin_off[F], fin_size[F],
 *   fout_off[n_out], fout_size[n_out], ev_card[n_evid], evf_ptr[F+1 or 0], evf_var[n_evf],
 *   evf_stride[n_evf], tasks[n_tasks][JT_TASK_WORDS], msgs[n_msgs][JT_MSG_WORDS],
 *   launches[n_launches][JT_LAUNCH_WORDS], then int32 tables[n_tab (padded to even)].
 *
 * Projection task semantics (s = output index, r = index over the remaining clique axes):
 *   term(s,r) = src[S(s)+R(r)] * prod_j rmsg_j[A_j(s)+B_j(r)]      acc(s) = sum_r term(s,r)
 *   sm(s)     = prod_j smsg_j[A_j(s)]
 *   out[s]    = acc(s)*sm(s);   bel[s] = out[s]*own[s];   beta[S(s)+R(r)] = term*sm(s)*own[s]
 * where every index map X(i) = tab[hi + i / n_lo] + tab[lo + i % n_lo].
 */
#ifndef JT_B200_H
#define JT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JT_ABI_VERSION 9

/* status codes */
#define JT_OK 0
#define JT_ERR_INVALID 1   /* bad argument or malformed plan blob */
#define JT_ERR_CUDA 2      /* CUDA runtime error (no device, launch failure, ...) */
#define JT_ERR_NOMEM 3

/* dtype */
#define JT_F32 0
#define JT_F64 1

/* flags of the stage calls */
#define JT_SEP_BELIEFS 1   /* distribute: also write separator beliefs (up*down), computation.py:210 */
#define JT_SKIP_MARGINAL 2 /* jt_propagate: stop after distribute */
#define JT_UNIFORM 4       /* init/collect/distribute: uniform mode (pass the same value to all three) */
#define JT_NO_UNIFORM 8    /* jt_propagate: do not enable uniform mode automatically */
#define JT_NO_BELIEFS 32    /* distribute/marginal/propagate: only messages and outputs are wanted -- clique
                              beliefs are not written and jt_marginal computes the outputs directly from
                              psi_C and the incoming messages (pass the same value to both stages) */
#define JT_NO_DENSE 64      /* uniform mode: keep every task on the projection kernels (no dense contractions);
                              pass the same value to every stage */
#define JT_LOGZ_ONLY 128    /* jt_normalize: only write log Z (the log of the total of output scope 0); the outputs
                              stay unnormalised */
#define JT_UNIFORM_VALID 16 /* uniform mode: the uniform workspace of this workspace already holds the
                              potentials and up-messages of these factor tables (an earlier call with
                              the same tables and the same workspace): skip recomputing them */

/* Semiring ("distributive law", reference sum_product.py:2-3, junctiontree.py:300-305) of the stage
 * calls, jt_normalize and jt_contract: which (+, x) pair the kernels use.  Pass the same value to
 * every stage of one propagation.  Log-domain semirings take log potentials as factor tables.
 *   sum-product  (+, *)           marginals and the partition function (the reference's only law)
 *   max-product  (max, *)         max-marginals: the value of the best joint state per entry (MAP)
 *   log-sum-exp  (logaddexp, +)   sum-product on log potentials, safe against underflow
 *   max-sum      (max, +)         max-product on log potentials */
#define JT_SR_SUM_PRODUCT 0x000
#define JT_SR_MAX_PRODUCT 0x100
#define JT_SR_LOG_SUM_EXP 0x200
#define JT_SR_MAX_SUM 0x300
#define JT_SR_MASK 0x300

/* plan blob header words */
#define JT_MAGIC 0x324E4C5042544ALL
enum {
    JT_H_MAGIC, JT_H_VERSION, JT_H_NCLIQUES, JT_H_NSEPS, JT_H_NFACTORS, JT_H_NEVID,
    JT_H_CLIQUE_ENTRIES, JT_H_SEP_ENTRIES, JT_H_FIN_ENTRIES, JT_H_FOUT_ENTRIES, JT_H_NTAB,
    JT_H_NTASKS, JT_H_NMSGS, JT_H_NLAUNCHES, JT_H_MAXDEPTH, JT_H_NEVF, JT_H_ROOT_ENTRIES,
    JT_H_UNI_ENTRIES, JT_H_NOUT, JT_H_LIK_ENTRIES, JT_H_WORDS
};
#define JT_TASK_WORDS 24
enum {
    JT_T_KIND, JT_T_SRC, JT_T_OUT, JT_T_BETA, JT_T_BEL, JT_T_OWN, JT_T_NS, JT_T_NR, JT_T_NSLO,
    JT_T_NRLO, JT_T_SRC_SHI, JT_T_SRC_SLO, JT_T_SRC_RHI, JT_T_SRC_RLO, JT_T_RMSG_BEGIN,
    JT_T_RMSG_END, JT_T_SMSG_BEGIN, JT_T_SMSG_END, JT_T_OUT_SPACE, JT_T_NODE, JT_T_AUX, JT_T_FLAGS
};
/* task flags, honoured in uniform mode only */
#define JT_TF_SRC_UNIFORM 1   /* src is the potential of a clique no evidence touches */
#define JT_TF_OWN_UNIFORM 2   /* the own up-message is uniform */
#define JT_TF_TASK_UNIFORM 4  /* every input is uniform: the task runs in the uniform workspace */
#define JT_MSG_WORDS 8
enum { JT_M_OFF, JT_M_AHI, JT_M_ALO, JT_M_BHI, JT_M_BLO, JT_M_FID, JT_M_UNI };
#define JT_LAUNCH_WORDS 4
enum { JT_L_PHASE, JT_L_BEGIN, JT_L_END, JT_L_LEVEL };
enum { JT_KIND_PROJECT = 0, JT_KIND_INIT = 1 };
enum {
    JT_PHASE_INIT, JT_PHASE_COLLECT, JT_PHASE_DIST_PRE, JT_PHASE_DIST_MAIN, JT_PHASE_MARGINAL,
    JT_PHASE_INIT_UNIFORM, JT_PHASE_INIT_INSTANCE, JT_PHASE_COLLECT_UNIFORM, JT_PHASE_COLLECT_INSTANCE,
    JT_PHASE_DIST_MAIN_MESSAGES, JT_PHASE_MARGINAL_DIRECT, JT_PHASE_DIST_UNIFORM, JT_PHASE_DIST_PRE_INSTANCE
};

typedef struct jt_plan jt_plan;

/* ---- library ---- */
int jt_abi_version(void);
const char* jt_last_error_string(void);
/* number of kernel launches issued by this library in this process (all threads) */
int64_t jt_launch_count(void);

/* ---- plan (host only until jt_plan_upload; no CUDA call is made by create/query/destroy) ---- */
int jt_plan_create(const void* blob, size_t nbytes, jt_plan** out);
void jt_plan_destroy(jt_plan* plan);
/* `what` is a JT_H_* header index */
int jt_plan_query(const jt_plan* plan, int what, int64_t* out);
/* entry offset and entry count of node k (cliques 0..N-1 then separators) in the workspace */
int jt_plan_node_range(const jt_plan* plan, int node, int64_t* offset, int64_t* count);
/* entry offset of separator node k's up / down message buffers */
int jt_plan_message_offsets(const jt_plan* plan, int sep_node, int64_t* up, int64_t* down);
/* bytes of workspace for a batch of B */
int jt_workspace_bytes(const jt_plan* plan, int64_t B, int dtype, size_t* out);
/* byte offsets inside the workspace: out4 = { factor offsets, error counter, uniform workspace,
 * total size }; the [entries][B] block starts at 0.  The W region of the dense contractions
 * (below) follows the uniform workspace and is included in the total. */
int jt_workspace_layout(const jt_plan* plan, int64_t B, int dtype, int64_t* out4);

/* Dense contractions (uniform mode, sum-product, float64, B >= 128; JT_DISABLE_DENSE=1 switches
 * them off).  A projection task -- the reference's E2 / E4 einsum, computation.py:84-88, 205-207 --
 * whose potential is shared by the batch and whose only per-instance input is one message M is a
 * batch of small dense products
 *     out[s_of[g][i]][b] = sum_k W[g][i][k] * M[mg[g] + mk[k]][b]
 * (g: the output axes M depends on, i: the other output axes, k: the summed axes M depends on;
 * W = potential x uniform messages, summed over the axes M does not see).  The groups are derived
 * from the task's index tables when the plan is loaded; such tasks then run on the FP64 tensor pipe
 * (mma.sync.m8n8k4.f64) with every row of M and of out moved once per batch tile, instead of one
 * row of M per (s, r) item.  These three calls expose the derivation (tests, tools):
 * out16 = { task, message, n_g, n_i, K, n_q, MT, n_it, n_k4, s_of, mg, mk, r_of, w_off, w_size,
 * launch }, the offsets s_of / mg / mk / r_of index the int32 table of jt_plan_dense_table. */
int jt_plan_dense_count(const jt_plan* plan);
int jt_plan_dense_get(const jt_plan* plan, int index, int64_t* out16);
const int32_t* jt_plan_dense_table(const jt_plan* plan, int64_t* count);
/* copy the schedule tables to the current CUDA device (idempotent) */
int jt_plan_upload(jt_plan* plan);

/*
 * ---- sparse workspaces (CUDA virtual memory management) ----
 * The same address layout as a dense workspace of jt_workspace_bytes(), but only the rows that
 * the stage calls touch when run with `flags` (e.g. JT_UNIFORM | JT_NO_BELIEFS: the streaming
 * pipelines) are backed by device memory; with most potentials uniform this is a small fraction
 * (config 5: 11 of 80 MB per instance), so chunks can be several times larger on the same GPU.
 * Use the pointer exactly like a dense workspace, with the same plan, B, dtype and a subset of
 * the behaviour implied by `flags` (a stage that touches other rows faults).  Batches of at most
 * 16 instances of small trees run as one launch in general mode and need a dense workspace.
 */
typedef struct jt_sparse_ws jt_sparse_ws;
/* the rows (entries of the [entries][B] block) the stages touch when run with `flags`, as merged
 * half-open intervals [begin, end): up to `capacity` pairs are written, *count receives their number */
int jt_workspace_sparse_rows(const jt_plan* plan, int flags, int64_t* intervals, int64_t capacity, int64_t* count);
/* bytes that would be backed by memory (2 MB granularity) and the dense size, without allocating */
int jt_workspace_sparse_bytes(const jt_plan* plan, int64_t B, int dtype, int flags, size_t* mapped, size_t* dense);
int jt_workspace_sparse_create(const jt_plan* plan, int64_t B, int dtype, int flags, jt_sparse_ws** out);
void* jt_workspace_sparse_ptr(const jt_sparse_ws* ws);
size_t jt_workspace_sparse_mapped(const jt_sparse_ws* ws);
void jt_workspace_sparse_destroy(jt_sparse_ws* ws);

/*
 * ---- stages; all pointers are device pointers, stream is a cudaStream_t ----
 *
 * factor_tables : concatenated factor tables in plan order, fin_entries values of `dtype`
 *                 (factors_batched = 0, shared by the batch), or [fin_entries][B] per-instance
 *                 tables (factors_batched = 1).
 * evidence      : int32 [B][n_evid] observed states (row-major), or NULL when the plan has no
 *                 per-instance evidence variables.  States outside [0, card) are clamped and
 *                 counted; see jt_evidence_errors.
 * flags         : JT_UNIFORM must be passed consistently to init, collect and distribute.
 *
 * Streams: every stage call enqueues on `stream` and returns without synchronising.  In uniform
 * mode a call may run independent kernels of one schedule level side by side, the uniform (B = 1)
 * half of the distribute pass beside the instance collect (jt_collect enqueues it; jt_distribute on
 * the same thread, plan and workspace finds it done, any other caller recomputes it) and the clique
 * beliefs of shared cliques beside the message chain -- on auxiliary streams that belong to the
 * calling thread and are joined back into `stream` (event wait) before the call returns.  So a stage
 * is complete exactly when `stream` has drained, calls from several host threads never share a
 * stream or an event, and the calls stay capturable into a CUDA graph (fork / join pattern).
 */
int jt_init(jt_plan* plan, const void* factor_tables, int factors_batched, const int32_t* evidence,
            int64_t B, int dtype, void* workspace, int flags, void* stream);
int jt_collect(jt_plan* plan, int64_t B, int dtype, void* workspace, int flags, void* stream);
int jt_distribute(jt_plan* plan, int64_t B, int dtype, void* workspace, int flags, void* stream);
/* factor_out: [fout_entries][B] values of `dtype` */
int jt_marginal(jt_plan* plan, int64_t B, int dtype, void* workspace, void* factor_out, int flags, void* stream);
/* init + collect + distribute (+ marginal unless JT_SKIP_MARGINAL) */
int jt_propagate(jt_plan* plan, const void* factor_tables, int factors_batched,
                 const int32_t* evidence, int64_t B, int dtype, void* workspace, void* factor_out,
                 int flags, void* stream);
/* 1 when jt_propagate(plan, ..., B, ..., flags) runs as a single launch of the whole-propagation
 * kernel (a few instances of a small tree), else 0 */
int jt_plan_single_launch(const jt_plan* plan, int64_t B, int flags);
/* JunctionTree.propagate (junctiontree.py:297-331) with HOST buffers in one call: factor tables
 * (and evidence) host -> device staging buffers, jt_propagate, factor_out device -> host, then
 * the stream is synchronised -- the results are in host_out on return.  host_* should be pinned;
 * dev_factors / dev_evidence / workspace / dev_out are caller-owned device buffers of the sizes
 * jt_propagate needs (dev_evidence and host_evidence may be NULL for plans without evidence
 * variables).  For small trees this is one kernel between two copies, ~25 us host to host. */
int jt_propagate_host(jt_plan* plan, const void* host_factors, size_t factor_bytes, const int32_t* host_evidence,
                      int64_t B, int dtype, void* dev_factors, int32_t* dev_evidence, void* workspace,
                      void* dev_out, void* host_out, size_t out_bytes, int flags, void* stream);

/* compute_beliefs (junctiontree/computation.py:37-246) of ONE instance with host buffers in one
 * call: clique potentials (clique_entries values, node order) host -> the clique rows of the
 * workspace, jt_collect + jt_distribute with separator beliefs (a single launch for small
 * trees), the clique and separator beliefs (clique_entries + sep_entries values, the reference's
 * node order) device -> host, stream synchronised. */
int jt_beliefs_host(jt_plan* plan, const void* host_potentials, int dtype, void* workspace, void* host_beliefs,
                    int flags, void* stream);

/* Output stage: divide every output scope of factor_out ([fout_entries][B], as written by
 * jt_marginal) by its sum over the scope, per instance; the sum of scope 0 -- the partition
 * function Z = P(evidence) that the reference computes at the root and discards,
 * computation.py:90-96 -- is written as log Z to logz[B] when logz is not NULL.  A scope whose sum
 * is 0 (impossible evidence) becomes all zeros and log Z = -inf.  `flags` selects the semiring
 * (JT_SR_*): the total is the semiring's reduction (sum / max / logsumexp) over the scope, the
 * entries are divided by it (log domain: it is subtracted) and logz receives its logarithm (log
 * domain: the total itself). */
int jt_normalize(jt_plan* plan, int64_t B, int dtype, void* factor_out, void* logz, int flags, void* stream);
/* number of out-of-range evidence states seen by jt_init calls on this workspace since it was
 * last zeroed (synchronises the stream) */
int jt_evidence_errors(jt_plan* plan, int64_t B, int dtype, void* workspace, void* stream, int64_t* out);

/*
 * ---- single operator: the SumProduct surface (sum_product.py:14-35) ----
 * out[s] = sum_r prod_j op_j[ A_j(s) + B_j(r) ], every operand and the output batch-innermost
 * ([n][B]).  `tables` (host, int32) holds all index tables; per operand j the four table
 * offsets are maps[4*j .. 4*j+3] = (a_hi, a_lo, b_hi, b_lo).  ops[j] are device pointers.
 * "project" is the case of one operand, "absorb" the case n_r = 1.  `flags`: JT_SR_* semiring
 * (sum and product above become the semiring's reduction and product).
 */
int jt_contract(const void* const* ops, int n_ops, const int32_t* tables, int64_t n_tab,
                const int32_t* maps, int64_t n_s, int64_t n_r, int64_t n_slo, int64_t n_rlo,
                int64_t B, int dtype, void* out, int flags, void* stream);

/* Strided row copy between a device buffer and a (pinned) host buffer on `stream`
 * (cudaMemcpy2DAsync): `rows` rows of `width_bytes`; to_host = 1 for device -> host.  Used to
 * move a batch chunk [rows][chunk] into / out of a [rows][B] buffer without a host transpose. */
int jt_copy_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                 size_t rows, int to_host, void* stream);

/* Hugin separator ratio: out[i] = old[i] != 0 ? new[i] / old[i] : 0, n contiguous elements.
 * (The propagation schedule itself is division-free; this serves SumProduct.absorb(old=...).) */
int jt_ratio(const void* new_values, const void* old_values, void* out, int64_t n, int dtype, void* stream);

/*
 * ---- compile phase on the host, in C++ (no CUDA call; works without a device) ----
 *
 * Replaces find_triangulation (junctiontree/construction.py:176-353) and construct_junction_tree
 * (:522-601) of the reference -- with valid trees (the reference's are not) --
 * and emits the plan blob that the reference's per-call Python bookkeeping corresponds to
 * (computation.py:47-96,140-224; junctiontree.py:203-274).  Variables are integers 0..n_vars-1;
 * the integer is also the tie-break rank.  Lists of variable lists are CSR pairs (ptr[n+1], data).
 * Results of the two graph calls come back as a list of int32 arrays (jt_ibuf).
 */
typedef struct jt_ibuf jt_ibuf;
int jt_ibuf_count(const jt_ibuf* buf);
int64_t jt_ibuf_size(const jt_ibuf* buf, int k);           /* -1 when k is out of range */
const int32_t* jt_ibuf_data(const jt_ibuf* buf, int k);
void jt_ibuf_destroy(jt_ibuf* buf);
/* frees a blob returned by jt_plan_build */
void jt_free(void* blob);

/* Min-fill elimination (ties: cluster weight, then rank; scores always those of the current
 * graph) or the given elimination `order` (n_order entries, a permutation of the variables that
 * occur in a factor; NULL for min-fill).  Every elimination cluster not contained in an earlier
 * one is a maximal clique.  Output arrays: 0 clique_ptr, 1 clique_vars (ascending), 2
 * factor_to_clique, 3 fill-in edges (pairs), 4 elimination order. */
int jt_triangulate(int32_t n_vars, const int64_t* var_sizes, int32_t n_factors, const int32_t* factor_ptr,
                   const int32_t* factor_vars, const int32_t* order, int32_t n_order, jt_ibuf** out);

/* Maximum spanning tree over the clique graph (Kruskal: most shared variables first, then the
 * lighter clique pair, then the pair index), unconnected components joined by empty separators,
 * rooted at `root` or, with root = -1, at the centre of the tree (fewest levels).  Output
 * arrays: 0 sep_ptr, 1 sep_vars (ascending), 2 parent[n_cliques] (-1 for the root), 3
 * parent_sep[n_cliques] (node id n_cliques + k, -1 for the root), 4 cliques in breadth-first
 * order.  The children of a clique are its successors in that order, in order of appearance. */
int jt_junction_tree(int32_t n_vars, const int64_t* var_sizes, int32_t n_cliques, const int32_t* clique_ptr,
                     const int32_t* clique_vars, int32_t root, jt_ibuf** out);

/* Emit the plan blob (the input of jt_plan_create) for a tree.
 *   sizes / full_sizes : effective size per variable (1 for observed ones) / size of the stored
 *                        factor axes (NULL: same as sizes)
 *   nodes              : cliques 0..n_cliques-1 then separators, axis order as listed
 *   has_tree = 0       : bare clique graph (init and marginal stages only; n_seps ignored)
 *   order/parent/parent_sep : the tree as produced by jt_junction_tree
 *   n_factors = -1     : no factors (collect/distribute on given potentials: compute_beliefs)
 *   n_outputs = -1     : marginalise to the factor scopes (CliqueGraph.marginalize); otherwise to
 *                        the listed scopes, each from the smallest clique containing it
 *   likelihood_vars    : variables with soft evidence (a per-instance likelihood vector in the
 *                        workspace's likelihood region, see the layout above)
 * Byte-identical to junctiontree/schedule.py Plan.to_blob() of this package for the same input. */
int jt_plan_build(int32_t n_vars, const int64_t* sizes, const int64_t* full_sizes, int32_t n_cliques,
                  int32_t n_seps, const int32_t* node_ptr, const int32_t* node_vars, int32_t has_tree,
                  const int32_t* order, const int32_t* parent, const int32_t* parent_sep, int32_t n_factors,
                  const int32_t* factor_ptr, const int32_t* factor_vars, const int32_t* factor_to_clique,
                  int32_t n_evidence, const int32_t* evidence_vars, int32_t n_outputs,
                  const int32_t* output_ptr, const int32_t* output_vars, int32_t n_likelihood,
                  const int32_t* likelihood_vars, void** blob, size_t* nbytes);

#ifdef __cplusplus
}
#endif
#endif /* JT_B200_H */
