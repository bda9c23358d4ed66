/*
 * tob200.h — C ABI of the B200-native contraction executor for TensorOrder.
 *
 * This is the drop-in boundary for ONE path of vardigroup/TensorOrder: the execution stage
 * behind `--tensor_library` (reference: src/tensor_network/tensor_apis/base_api.py:8-28,
 * called from src/execution.py:136).  The reference has no FFI of its own for this path
 * (its backends are Python classes calling numpy / TensorFlow / JAX); the binding a
 * maintainer adds is the ctypes stub shown in INTEGRATION.md (shipped as
 * tensororder_b200/cabi.py), called only from the `B200API` backend class.
 *
 * Conventions: plain C, caller-owned input buffers, callee-owned plan; every function
 * returns TOB_OK (0) or an error code and leaves a message in tob_last_error().  No CPU
 * fallback exists anywhere behind this interface: without a CUDA device the run/upload
 * entry points fail with TOB_E_NODEVICE.
 *
 * Data model: every index has extent 2, so a rank-r tensor is 2^r doubles
 * and numpy axis j of a C-ordered array is address bit (r-1-j).
 */
#ifndef TOB200_H
#define TOB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOB_OK 0
#define TOB_E_INVALID 1  /* malformed plan / arguments                              */
#define TOB_E_OOM 2      /* arena does not fit: caller raises OutOfMemoryError,      */
                         /* the reference then slices once more (execution.py:140)  */
#define TOB_E_CUDA 3     /* CUDA runtime error (message has the cudaError string)    */
#define TOB_E_NODEVICE 4 /* no usable CUDA device                                    */

typedef struct tob_plan tob_plan;

/*
 * Flat form of a sliced execution plan.  Replaces the object graph the reference walks in
 * TensorNetwork.identify (src/tensor_network/tensor_network.pyx:142-156) and
 * JaxAPI.contract_sliced_base (src/tensor_network/tensor_apis/jax_apis.py:253-277):
 *   - nodes are in post-order (src/contraction_methods/contraction_tree.pyx:174-183);
 *   - a leaf names an entry of the leaf table; its axes carry the network edge id, or
 *     -(g+1) when the axis is fixed by slice group g (sliced axes are removed from the tree
 *     exactly like remove_sliced_indices_from, tensor_network.pyx:444-468);
 *   - contracted axes of a join are the edge ids its two children share
 *     (compute_join_properties, contraction_tree.pyx:248-288).
 * Slice ids enumerate itertools.product over the groups, first group most significant
 * (tensor_network.pyx:396): bit (n_slice_groups-1-g) of the slice id is the value of group g.
 */
typedef struct {
    int32_t n_nodes;
    const int32_t* node_left;        /* [n_nodes] post-order position of the left child, -1 for a leaf  */
    const int32_t* node_right;       /* [n_nodes] same for the right child                                */
    const int32_t* node_leaf;        /* [n_nodes] leaf-table index, -1 for a join                         */
    int32_t n_leaves;
    const int32_t* leaf_rank;        /* [n_leaves] rank of the stored leaf tensor (all axes)              */
    const int64_t* leaf_data_offset; /* [n_leaves] offset in doubles into the leaf data buffer            */
    const int32_t* leaf_axis_start;  /* [n_leaves+1] CSR offsets into leaf_axis_edge                      */
    const int32_t* leaf_axis_edge;   /* per axis, numpy order (axis 0 = most significant address bit)     */
    int32_t n_slice_groups;
    int64_t leaf_data_len;           /* doubles in the leaf data buffer passed to tob_plan_upload         */
} tob_plan_desc;

typedef struct {
    int32_t device;          /* CUDA device ordinal                                                        */
    int32_t use_graph;       /* 2 (default): CUDA graph per slice when the slice is launch-bound, plain     */
                             /* stream launches (with per-GEMM CUDA events) otherwise; 1 always; 0 never   */
    int32_t kernel_policy;   /* 0 auto, 1 force the generic kernel for every join (debug / parity)         */
    int32_t hoist_invariant; /* 1: compute slice-invariant subtrees once (default), 0: per slice           */
    int64_t mem_limit_bytes; /* 0: use free device memory; else refuse (TOB_E_OOM) plans needing more      */
    int32_t use_microtree;   /* 1 (default): subtrees made only of tiny joins run inside ONE kernel launch  */
                             /* (one CTA per subtree); 0: one launch per join                              */
    int32_t slice_lanes;     /* slices in flight at once, each with its own arena/workspace/stream:        */
                             /* 0 (default) = 2 when the second arena costs <= 2 GiB, else 1; 1; 2          */
    int32_t dag_branches;    /* joins of independent subtrees run concurrently on up to this many streams   */
                             /* per lane (fork/join by events; captured into the slice graph as a DAG):     */
                             /* 0 (default) = 16; 1 = one stream, strict post-order                          */
} tob_options;

void tob_default_options(tob_options* opt);

/* Host-only: analyses the plan, chooses layouts/kernels per join, plans the arena.  No GPU needed. */
int tob_plan_create(const tob_plan_desc* desc, const tob_options* opt, tob_plan** out);

/* Bytes of device memory the plan needs (leaves + arena + workspaces).  Replaces the entry count the
 * reference derives from estimate_cost (contraction_tree.pyx:382-446) when checking --mem_limit. */
int64_t tob_plan_peak_bytes(const tob_plan* plan);

/* Number of slices 2^n_slice_groups. */
uint64_t tob_plan_num_slices(const tob_plan* plan);

/* JSON description of the compiled program (one op per join: kernel, operand layouts, arena offsets).
 * Writes at most cap bytes (NUL-terminated) and returns the full length needed (excluding NUL). */
int64_t tob_plan_describe(const tob_plan* plan, char* buf, int64_t cap);

/* Allocates the device arena (TOB_E_OOM if it does not fit) and copies the leaf tensors host->device.
 * leaf_data: leaf_data_len doubles, each leaf C-ordered as Tensor.build() returns it
 * (src/tensor_network/tensor_network_constructions.py:69-99,144-152). */
int tob_plan_upload(tob_plan* plan, const double* leaf_data, int64_t n_doubles);

/* Plan cache support (SURVEY.md §8b "Ownership": the backend owns all device memory and must release it
 * per call or cache it keyed by plan identity).  tob_plan_update_leaves re-reads the caller's leaf values
 * into an already uploaded plan (one pinned H2D copy of the leaf region; arena, tables and captured graphs
 * are kept), or behaves like tob_plan_upload when the plan holds no device memory.  tob_plan_release
 * returns the plan's device and pinned blocks to the process-wide pool and keeps the compiled program, so a
 * later tob_plan_upload costs no recompilation. */
int tob_plan_update_leaves(tob_plan* plan, const double* leaf_data, int64_t n_doubles);
int tob_plan_release(tob_plan* plan);

/* Frees every cached (idle) device and pinned block of the process-wide pool; device < 0 = all devices.
 * For hosts that share the GPU with other allocators (e.g. torch). */
int tob_pool_trim(int32_t device);

/* Contracts slices first, first+stride, ... (count of them) and writes the float64 sum of their
 * rank-0 results to *result (host).  Replaces the loop of BaseTensorAPI.contract_sliced
 * (base_api.py:21-28); count = min(num_slice_limit, 2^s) with first=0, stride=1 reproduces it;
 * first=rank, stride=world gives the per-GPU partial sum of the multi-GPU partition. */
int tob_plan_run(tob_plan* plan, uint64_t first, uint64_t count, uint64_t stride, double* result);

/* Continuation form: the device accumulator starts at `initial` (so chunked calls reproduce the
 * sequential slice-order sum bit for bit) and TOB_RUN_SKIP_INVARIANT skips the slice-invariant
 * prologue already computed by an earlier call on this plan.  Lets the Python host return to the
 * interpreter between chunks of slices, where the reference's SIGALRM timeout (util.py:32-39) fires. */
#define TOB_RUN_SKIP_INVARIANT 1
int tob_plan_run_ex(tob_plan* plan, uint64_t first, uint64_t count, uint64_t stride, double initial,
                    int32_t flags, double* result);

/* Asynchronous form, for hosts that keep several independent plans in flight on one GPU (the launch-bound
 * stretches of one contraction then overlap the GEMMs of another): tob_plan_run_async issues the run on the
 * plan's stream(s) and returns; when after_stream (a cudaStream_t) is not NULL the run starts behind the work
 * already enqueued on it.  tob_plan_join makes `stream` wait (on the device) for the run's last kernel;
 * tob_plan_wait blocks the host until the run is complete and returns its result like tob_plan_run_ex.
 * One run in flight per plan; at most 4096 slices per asynchronous call. */
int tob_plan_run_async(tob_plan* plan, uint64_t first, uint64_t count, uint64_t stride, double initial,
                       int32_t flags, void* after_stream);
int tob_plan_join(tob_plan* plan, void* stream);
int tob_plan_wait(tob_plan* plan, double* result);

/* Device time (CUDA events on the plan's stream) of the last tob_plan_run, in milliseconds. */
double tob_plan_last_ms(const tob_plan* plan);

/* Host time (ms) the last tob_plan_run spent issuing work, i.e. before it started waiting for the device:
 * tells a launch-bound run that is limited by the host's launch rate from one limited by the device. */
double tob_plan_last_issue_ms(const tob_plan* plan);

/* Number of kernels launched by the last tob_plan_run. */
int64_t tob_plan_last_launches(const tob_plan* plan);

/* Exact mode (the reference's `--entry_type=bigint`, numpy_apis.py:15-22): with a prime modulus p < 2^23
 * every kernel reduces its sums modulo p (tensors hold residues as doubles; sums of <= 128 products stay
 * below 2^53, so all arithmetic is exact, on the same DMMA path).  The leaves passed to tob_plan_upload
 * must already be residues in [0, p).  The host combines the residues of several primes by CRT.
 * modulus = 0 restores float64 arithmetic. */
int tob_plan_set_modulus(tob_plan* plan, double modulus);

/* Per-kernel timing of the DMMA GEMM launches (off by default).  When on, every GEMM of a stream-mode run
 * is bracketed by a CUDA-event pair and GEMMs of the two slice lanes are chained so each pair brackets
 * one GEMM (costs ~3 % on mid-size sliced plans, where overlapping GEMM tails otherwise helps).
 * tob_plan_last_gemm: summed event duration, algorithmic flops (2*2^(fL+fR+k) per join)
 * and launch count of the last run; all zero when timing is off or the run was replayed as a graph. */
int tob_plan_set_gemm_timing(tob_plan* plan, int32_t on);
int tob_plan_last_gemm(const tob_plan* plan, double* ms, double* flops, int64_t* launches);

/* Run on a caller-owned CUDA stream (cudaStream_t) instead of the plan's own; call after
 * tob_plan_upload.  The caller keeps the stream alive for the life of the plan. */
int tob_plan_set_stream(tob_plan* plan, void* stream);

/* Runs one slice op by op with CUDA events around every op; ms_per_op has n_ops entries
 * (tob_plan_num_ops).  Used for the per-node roofline report. */
int64_t tob_plan_num_ops(const tob_plan* plan);
/* Work of the compiled program (the SURVEY.md §8d model: 2*2^(fL+fR+k) flops per join): per slice and of the
 * slice-invariant prologue, and the launches one slice issues.  A host that must return to its interpreter at
 * intervals (the reference's SIGALRM TimeoutTimer, src/util/util.py:32-39) sizes its first chunk of slices from it. */
int tob_plan_work(const tob_plan* plan, double* slice_flops, double* invariant_flops, int64_t* slice_launches);
int tob_plan_profile(tob_plan* plan, uint64_t slice, float* ms_per_op, int64_t n_ops, double* result);

/* Verification aids: run the first n_ops ops (slice-invariant list, then the per-slice list) of one slice
 * sequentially on one stream and wait; read doubles back from the leaf region (space 0) or the arena (1).  With
 * them every join's output is compared with the numpy interpreter of the same program (tools/verify_ops.py). */
int tob_plan_debug_run(tob_plan* plan, uint64_t slice, int64_t n_ops);
int tob_plan_debug_read(tob_plan* plan, int32_t space, int64_t offset, int64_t n, double* out);

void tob_plan_destroy(tob_plan* plan);

/*
 * Single pairwise contraction, numpy.tensordot semantics (reference call site
 * NumpyAPI.tensordot, src/tensor_network/tensor_apis/numpy_apis.py:42-43): contracts axes_a[i] of a
 * with axes_b[i] of b; the result has a's free axes in order followed by b's free axes in order.
 * All buffers are DEVICE pointers (a: 2^rank_a doubles, b: 2^rank_b, c: 2^(rank_a+rank_b-2n)).
 * workspace may be NULL when both operands are already GEMM-ready (contracted axes trailing, in pair
 * order); it holds the permuted operands and split-K partials otherwise.  ms (optional) points to 3
 * floats: {operand-permutation ms, contraction ms, kernel kind (0 generic, 1 DMMA GEMM)}.
 * kernel_policy as in tob_options.
 */
int tob_tensordot_device(const double* a, int32_t rank_a, const double* b, int32_t rank_b,
                         const int32_t* axes_a, const int32_t* axes_b, int32_t n_axes, double* c,
                         double* workspace, int64_t workspace_bytes, int32_t kernel_policy,
                         void* stream, float* ms);

/* Same with HOST buffers (copies in, contracts, copies out) — the `api.tensordot` secondary entry. */
int tob_tensordot_host(const double* a, int32_t rank_a, const double* b, int32_t rank_b,
                       const int32_t* axes_a, const int32_t* axes_b, int32_t n_axes, double* c);

/* Index permutation of a rank-r binary tensor, numpy.transpose semantics: out = in.transpose(perm).
 * DEVICE pointers.  (Stand-alone kernel of north_star item (1); 16 B moved per element.) */
int tob_permute_device(const double* in, double* out, int32_t rank, const int32_t* perm, void* stream, float* ms);

/* Per-join dispatch parameters (generic <-> GEMM crossover, kernel classes, the split-K time model).  The
 * defaults are the measured table tensororder_b200/csrc/tob_dispatch_table.h, generated by
 * tools/fit_dispatch.py; these two calls read / override single values (experiments and the fit itself).
 * They affect plans created afterwards.  Keys: see `kTuneFields` in csrc/tob_compile.cpp. */
int tob_tuning_set(const char* key, double value);
int tob_tuning_get(const char* key, double* value);
/* The split-K time model the plan compiler minimises (microseconds) for a DMMA GEMM join with K split 2^c ways. */
double tob_gemm_time_model_us(int32_t m, int32_t n, int32_t k, int32_t tm_log2, int32_t tn_log2, int32_t c);

/* Creates the CUDA context on `device` and primes the block / stream pools, so the first contraction does not pay
 * for it (TOB_E_NODEVICE without a GPU).  Optional: every entry point initialises the device on demand. */
int tob_warm(int32_t device);

int tob_device_count(void);
const char* tob_version(void);
const char* tob_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* TOB200_H */
