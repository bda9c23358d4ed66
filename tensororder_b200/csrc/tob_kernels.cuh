// Device-side parameter blocks and launchers (sm_100a).  See tob_kernels.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "tob_internal.h"

namespace tob {

// A mask split into runs of adjacent set bits: pdep/pext in a handful of shift+and ops.
struct BitRuns {
    int32_t n;
    uint8_t dst[24];  // position of the run in the scattered word
    uint8_t src[24];  // position of the run in the compact word
    uint8_t len[24];
};
BitRuns make_runs(uint64_t mask);

struct KParams {
    const double* a;
    const double* b;
    double* c;
    double* ws;                 // split-K partials / stream-K partial tiles (or nullptr)
    unsigned* sk_flags;         // stream-K: [0] start-order counter, [1 + rank] partial-ready flags (zero between launches)
    const long long* leaf_off;  // per-leaf slice offsets in doubles (device), may be nullptr
    int32_t a_leaf, b_leaf;     // index into leaf_off, or -1
    int32_t m, n, k;
    int32_t ksplit_log2;
    int32_t streamk;            // > 0: stream-K launch with this many CTAs (k_gemm_dmma_sk)
    int32_t raster_group_log2;  // persistent short-K kernel: M-tiles per raster group (set by the launcher)
    uint64_t mask_m, mask_n;
    BitRuns runs_m, runs_n;
    // exact mode (entry type "bigint"): every tensor holds residues modulo the prime `modp` < 2^23 as
    // doubles; sums of up to 128 products stay below 2^53 and are exact, then get reduced.  0 = float64.
    double modp, inv_modp;
};

// One join of a micro stage (device copy).  m + n <= 14 (mask_m has 16 bits), operands <= 2^14 doubles.
struct MicroOpDev {
    long long a_off, b_off, c_off;  // doubles: a/b inside their space, c inside the arena
    int32_t a_leaf, b_leaf;         // leaf_off index or -1
    uint16_t mask_m;
    uint8_t a_space, b_space;       // 0 leaves, 1 this lane's arena, 2 lane 0's arena (slice-invariant tensor)
    uint8_t m, n, k;
    uint8_t fwd_out;                // 1: the next join of this CTA reads this result from the shared forward buffer
    // operand source inside k_microtree: 0 global memory, 1 the previous join's result (shared forward
    // buffer), 2 leaf staged into the CTA's shared leaf cache at a_soff / b_soff (doubles)
    uint8_t a_src, b_src;
    uint16_t a_soff, b_soff;
    uint16_t pad2;                  // 48 bytes: descriptors are staged into shared memory with 16-byte copies
};
constexpr int kMicroLeafCache = 2048;   // doubles of staged leaves per CTA
constexpr int kMicroFwdMax = 2048;      // largest result (doubles) forwarded through shared memory (256-thread stages)
constexpr int kMicroFwdMaxBig = 8192;   // same for the 1024-thread stages (2 x 64 KB of forward buffers)
constexpr int kMicroDescBytes = 48 * 1024;  // shared memory for staged join descriptors (1024 joins per CTA)
constexpr int kMicroStageMax = 64;      // largest leaf operand (doubles) staged
static_assert(sizeof(MicroOpDev) % 16 == 0, "MicroOpDev is staged with 16-byte copies");

struct DevState {                  // one per lane
    unsigned long long next_slice;  // slice id the lane's next slice uses
    unsigned long long stride;
    unsigned long long slot;        // index into the per-slice result buffer for the lane's current slice
    unsigned long long slot_stride;
};

struct SliceTables {            // device pointers
    const int32_t* term_start;  // [n_leaves+1]
    const uint8_t* id_bit;      // per term: bit of the slice id
    const uint8_t* addr_bit;    // per term: address bit it selects
    long long* leaf_off;        // [n_leaves] out
    int32_t n_leaves;
};

struct PermuteParams {
    const double* in;
    double* out;
    int32_t rank;
    int32_t tbits;              // tile bits
    uint8_t in_pos[16];         // tile bit j (input order) -> input address bit
    uint8_t in_to_tile[16];     // tile bit j (input order) -> bit index in the output-ordered tile index
    uint8_t out_pos[16];        // tile bit j (output order) -> output address bit
    int32_t nrest;              // non-tile bits
    uint8_t rest_out[40];       // rest bit j -> output address bit
    uint8_t rest_in[40];        // rest bit j -> input address bit
};

cudaError_t launch_contract(const Op& op, const KParams& p, cudaStream_t stream, int* launches);
cudaError_t launch_microtree(const MicroOpDev* ops, const int32_t* cta_start, int n_ctas, int smem_ops, int threads, const double* leaves,
                             double* arena, const double* arena0, const long long* leaf_off, double modp,
                             cudaStream_t stream);
cudaError_t launch_accum(DevState* st, const double* root, const long long* leaf_off, int root_leaf, double* results,
                         cudaStream_t stream);
cudaError_t launch_final_sum(double* acc, const double* results, int count, double initial, int use_previous, double modp,
                             cudaStream_t stream);
cudaError_t launch_begin_slice(DevState* st, SliceTables t, cudaStream_t stream);
cudaError_t launch_permute(const double* in, double* out, int rank, const int32_t* src_bit, cudaStream_t stream);
cudaError_t configure_kernels();

}  // namespace tob
