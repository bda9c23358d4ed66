// Internal types of the B200 contraction executor (host side).  See DESIGN.md §3-§5.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tob200.h"

namespace tob {

// ------------------------------------------------------------------------------------------------
// Canonical layouts.  Every tensor is dense; address bit p holds one network edge.
// A tensor that is an operand of a join with sibling S is stored as
//        [ K' (edges shared with S, ascending edge id) | F' (its other edges, in the parent's
//          output order) ]            low bits  ------------------------------------>  high bits
// so every join is   C[pdep(mi,maskM)|pdep(ni,maskN)] = sum_kk A[mi<<k | kk] * B[ni<<k | kk].
// ------------------------------------------------------------------------------------------------
enum OpKind : int32_t {
    OP_GENERIC = 0,  // thread / warp / CTA per output element, bandwidth-bound
    OP_GEMM = 1,     // DMMA tile kernel, compute-bound
    OP_ACCUM = 2,    // acc += root scalar
    OP_MICRO = 3,    // one launch executing every mini join of one stage, one CTA per tree fragment
};

struct OperandRef {
    int32_t space = 0;       // 0: leaf region, 1: the running lane's arena, 2: lane 0's arena (slice-invariant tensor)
    int64_t offset = 0;      // doubles from the start of that region
    int32_t leaf = -1;       // leaf-table index when space == 0 (slice offset lookup), else -1
    int32_t node = -1;       // post-order position of the producing node
};

struct Op {
    int32_t kind = OP_GENERIC;
    int32_t node = -1;          // post-order position of the join this op computes
    OperandRef a, b;            // a = "M side", b = "N side" (may be swapped w.r.t. left/right)
    int64_t c_offset = 0;       // arena offset (doubles) of the result
    int32_t m = 0, n = 0, k = 0;
    uint64_t mask_m = 0;        // C address bits fed by mi (ascending); the rest of the low m+n bits by ni
    // generic kernel configuration
    int32_t threads_per_out = 1;  // 1, 32 or 256
    int32_t ksplit_log2 = 0;      // K split across CTAs (partials in the workspace, reduced deterministically)
    int32_t streamk = 0;          // > 0: stream-K GEMM on this many CTAs (the K steps of all tiles cut into equal ranges;
                                  // partial tiles in the workspace, summed by the tile's owner in CTA order)
    // gemm kernel configuration
    int32_t tm_log2 = 7, tn_log2 = 7;
    int64_t ws_offset = -1;     // workspace offset (doubles) for split-K partials
    int32_t invariant = 0;      // 1: independent of the slice id (hoisted)
    int32_t micro_which = -1;   // OP_MICRO: index into Program::micro
    double flops = 0, bytes = 0;
    // DAG schedule inside the op's list (schedule_branches): ops of one branch run in list order on one
    // stream; `waits` names ops of OTHER branches (list indices) that must have finished first — the
    // producers of the operands and the last users of the arena space the result overwrites.
    int32_t branch = 0;
    int32_t signal = 0;           // 1: an op of another branch waits for this one (record an event)
    std::vector<int32_t> waits;
};

// One micro stage (tob_compile.cpp "micro stages"): every mini join of one level of one phase, grouped by
// CTA; inside a CTA the joins run one after the other with a CTA barrier in between.
struct MicroProgram {
    std::vector<Op> ops;              // grouped by CTA, post-order inside a fragment
    std::vector<int32_t> cta_start;   // [n_ctas + 1]
    int32_t threads = 256;            // CTA size: 1024 when the stage has results above 2^12 doubles
};

struct LeafInfo {
    int32_t rank = 0;                 // stored rank (live + sliced axes)
    int32_t live_rank = 0;
    int64_t src_offset = 0;           // doubles, in the caller's leaf buffer (numpy order)
    int64_t dev_offset = 0;           // doubles, in the device leaf region (canonical order)
    std::vector<int32_t> axis_edge;   // numpy axis -> edge id or -(g+1)
    // permutation applied at upload: device address bit p takes source address bit src_bit[p]
    std::vector<int32_t> src_bit;     // size rank; first live_rank entries are the canonical live layout
    // slice offset = sum over terms of ((slice_id >> id_bit) & 1) << addr_bit   (doubles)
    std::vector<int32_t> slice_id_bit, slice_addr_bit;
};

struct NodeInfo {
    int32_t left = -1, right = -1, leaf = -1, parent = -1;
    std::vector<int32_t> edges;       // sorted edge ids
    std::vector<int32_t> layout;      // canonical layout: layout[p] = edge id at address bit p
    int32_t k_with_sibling = 0;       // how many low bits of `layout` are contracted at the parent
    bool slice_dependent = false;
    OperandRef where;                 // where the node's tensor lives when it is consumed
};

struct Program {
    tob_options opt;
    int32_t n_slice_groups = 0;
    std::vector<NodeInfo> nodes;
    std::vector<LeafInfo> leaves;
    std::vector<Op> invariant_ops;    // run once per tob_plan_run
    std::vector<Op> slice_ops;        // run once per slice
    std::vector<MicroProgram> micro;  // micro stages, in execution order (Op::micro_which indexes this)
    int64_t leaf_doubles = 0;         // device leaf region
    int64_t arena_doubles = 0;        // intermediates
    int64_t ws_doubles = 0;           // split-K workspace
    int32_t lanes = 1;                // slices in flight at once (each lane has its own arena + workspace)
    int32_t branches = 1;             // streams per lane used by the DAG schedule of the op lists
    int64_t src_leaf_len = 0;
    double total_flops = 0, total_bytes = 0;
    OperandRef root;                  // rank-0 result of one slice
};

int compile(const tob_plan_desc* desc, const tob_options* opt, Program* out, std::string* err);
std::string describe(const Program& p);

// Shared by compile and the stand-alone tensordot: pick kernel + configuration for a canonical join.
void choose_kernel(Op* op, int32_t kernel_policy, bool allow_splitk);

// Assigns Op::branch / waits / signal for one op list; returns the number of branches used.
int schedule_branches(std::vector<Op>* list, int max_branches);

void set_error(const std::string& msg);

// SM count used for grid sizing and the split-K wave model (default 148; queried from the device when one exists)
int num_sms();
void set_num_sms(int n);

// stream-K GEMM (k_gemm_dmma_sk): one partial 128x64 tile per CTA in the workspace, a 4 KB counter + flag block per lane
constexpr int kSkFlagBytes = 4096;
constexpr int64_t kSkSlotDoubles = 128 * 64;
constexpr int kSkMaxSegs = 16;  // tiles a CTA's range may touch

// Per-join dispatch parameters: defaults come from the measured table tob_dispatch_table.h (generated by
// tools/fit_dispatch.py); tob_tuning_set overrides single values at run time (experiments, the fit itself).
struct Tuning {
    int gemm_min_free, gemm_min_k, gemm_smallk_min_free;
    int gemm_min_out[17];   // by k (k > 16 uses [16]): a join runs on the DMMA GEMM kernel when m + n >= gemm_min_out[k]
    int t1_max_k, t1_small_out, t1_small_max_k, t32_max_k, t32_min_out;
    int persist_max_k;
    double sm_gflops, alone_frac, gemm_fix_us, reduce_gbs, reduce_fix_us;
    int max_ksplit_log2, min_k_per_split_log2;
    int force_ksplit_log2;  // >= 0: experiments only — every split-capable GEMM uses this split
    int streamk;            // 0: never, 1: where the time model says so, 2: experiments only — every eligible GEMM
    int streamk_min_tiles_log2;  // stream-K needs at least this many tiles (few tiles => many partials per owner)
    int streamk_max_tiles_log2;  // ... and at most this many (ranges over several tiles lose the L2 locality of the raster)
    int streamk_max_steps;       // ... and at most this many K steps per CTA (long K: the one-tile-per-CTA grid is as fast)
    double streamk_fix_us;  // modelled cost of the partial-tile exchange
    int store_group_log2;   // persistent short-K kernel: M-tiles per raster group (0: N-tiles fastest)
    int permute_low_bits, permute_ctas_per_sm;  // stand-alone permutation kernel: tile shape / grid (0 = defaults)
    int t256_ctas_log2;     // CTA-per-output kernel: K is split until outputs * splits reach 2^this CTAs (chunks stay >= 2^12 elements)
    int ws_min_k;           // log2 K per split from which the warp-specialised kernels run (below: k_gemm_dmma, one CTA barrier per K step)
    int store_tile;         // row-streamed persistent kernels: 0 = never, 1 = K = 16 (k_gemm_dmma_p1), 2 = also K = 32 from 2048 tiles on (k_gemm_dmma_wp)
};
Tuning& tuning();
bool tuning_set(const char* key, double value);
bool tuning_get(const char* key, double* value);
// Modelled duration (microseconds) of a GEMM-kernel join with K split 2^c ways (choose_kernel minimises it).
double gemm_time_model_us(int m, int n, int k, int tm_log2, int tn_log2, int c);
// Modelled duration of the same join on the stream-K kernel; `ctas` receives its grid.  < 0: not eligible.
double streamk_time_model_us(int m, int n, int k, int* ctas);

}  // namespace tob
