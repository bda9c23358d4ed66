// Host-side plan compiler: flat plan (tob_plan_desc) -> Program (ops, canonical layouts, arena).
// Pure host code, no CUDA calls: runs (and is tested) on machines without a GPU.
//
// Reference semantics restated here (paths relative to /root/reference):
//   * contracted indices of a join = edge ids present in both children
//     (ContractionTreeContext.compute_join_properties, src/contraction_methods/contraction_tree.pyx:248-288);
//   * sliced axes are dropped from the tree (TensorNetwork.remove_sliced_indices_from,
//     src/tensor_network/tensor_network.pyx:444-468) and each leaf is indexed by the slice id
//     (SliceSequence.reordered_tensor, src/tensor_network/tensor.py:95-119);
//   * only the rank-0 root is observable (BaseTensorAPI.contract_sliced, base_api.py:26-27), so the
//     layout of every intermediate is ours to choose.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <set>
#include <sstream>
#include <chrono>

#include "tob_dispatch_table.h"
#include "tob_internal.h"

namespace tob {

static const int kMaxRank = 40;       // 2^40 doubles is far beyond 180 GB; guards the bit math
static const int64_t kAlign = 32;     // arena alignment in doubles (256 B)
// SM count of the device the plans run on: 148 on a full B200; tob_exec.cu replaces it with the queried
// cudaDevAttrMultiProcessorCount (other sm_100 parts, MIG slices) before the first plan is compiled
static int g_num_sms = 148;
int num_sms() { return g_num_sms; }
void set_num_sms(int n) { if (n > 0) g_num_sms = n; }
#define kNumSMs (num_sms())

static int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// Kernel choice for a canonical join  C[2^(m+n)] = A[2^m x 2^k] . B[2^n x 2^k]^T
// ------------------------------------------------------------------------------------------------
Tuning& tuning() {
    static Tuning t = [] {
        Tuning x;
        x.gemm_min_free = TOB_TUNE_GEMM_MIN_FREE;
        x.gemm_min_k = TOB_TUNE_GEMM_MIN_K;
        x.gemm_smallk_min_free = TOB_TUNE_GEMM_SMALLK_MIN_FREE;
        const int by_k[17] = TOB_TUNE_GEMM_MIN_OUT_BY_K;
        for (int i = 0; i < 17; i++) x.gemm_min_out[i] = by_k[i];
        x.t1_max_k = TOB_TUNE_T1_MAX_K;
        x.t1_small_out = TOB_TUNE_T1_SMALL_OUT;
        x.t1_small_max_k = TOB_TUNE_T1_SMALL_MAX_K;
        x.t32_max_k = TOB_TUNE_T32_MAX_K;
        x.t32_min_out = TOB_TUNE_T32_MIN_OUT;
        x.persist_max_k = TOB_TUNE_PERSIST_MAX_K;
        x.sm_gflops = TOB_TUNE_SM_GFLOPS;
        x.alone_frac = TOB_TUNE_ALONE_FRAC;
        x.gemm_fix_us = TOB_TUNE_GEMM_FIX_US;
        x.reduce_gbs = TOB_TUNE_REDUCE_GBS;
        x.reduce_fix_us = TOB_TUNE_REDUCE_FIX_US;
        x.max_ksplit_log2 = TOB_TUNE_MAX_KSPLIT_LOG2;
        x.min_k_per_split_log2 = TOB_TUNE_MIN_K_PER_SPLIT_LOG2;
        x.force_ksplit_log2 = -1;
        x.streamk = TOB_TUNE_STREAMK;
        x.streamk_min_tiles_log2 = TOB_TUNE_STREAMK_MIN_TILES_LOG2;
        x.streamk_max_tiles_log2 = TOB_TUNE_STREAMK_MAX_TILES_LOG2;
        x.streamk_max_steps = TOB_TUNE_STREAMK_MAX_STEPS;
        x.store_tile = TOB_TUNE_STORE_TILE;
        x.ws_min_k = TOB_TUNE_WS_MIN_K;
        x.t256_ctas_log2 = TOB_TUNE_T256_CTAS_LOG2;
        x.permute_low_bits = 0;
        x.permute_ctas_per_sm = 0;
        x.streamk_fix_us = TOB_TUNE_STREAMK_FIX_US;
        x.store_group_log2 = TOB_TUNE_STORE_GROUP_LOG2;
        return x;
    }();
    return t;
}

namespace {
struct TuneField { const char* key; int Tuning::*i; double Tuning::*d; };
const TuneField kTuneFields[] = {
    {"gemm_min_free", &Tuning::gemm_min_free, nullptr}, {"gemm_min_k", &Tuning::gemm_min_k, nullptr},
    {"gemm_smallk_min_free", &Tuning::gemm_smallk_min_free, nullptr},
    {"t1_max_k", &Tuning::t1_max_k, nullptr}, {"t1_small_out", &Tuning::t1_small_out, nullptr},
    {"t1_small_max_k", &Tuning::t1_small_max_k, nullptr},
    {"t32_max_k", &Tuning::t32_max_k, nullptr}, {"t32_min_out", &Tuning::t32_min_out, nullptr},
    {"persist_max_k", &Tuning::persist_max_k, nullptr},
    {"sm_gflops", nullptr, &Tuning::sm_gflops},
    {"alone_frac", nullptr, &Tuning::alone_frac}, {"gemm_fix_us", nullptr, &Tuning::gemm_fix_us},
    {"reduce_gbs", nullptr, &Tuning::reduce_gbs}, {"reduce_fix_us", nullptr, &Tuning::reduce_fix_us},
    {"max_ksplit_log2", &Tuning::max_ksplit_log2, nullptr}, {"min_k_per_split_log2", &Tuning::min_k_per_split_log2, nullptr},
    {"force_ksplit_log2", &Tuning::force_ksplit_log2, nullptr},
    {"streamk", &Tuning::streamk, nullptr}, {"streamk_min_tiles_log2", &Tuning::streamk_min_tiles_log2, nullptr},
    {"streamk_max_tiles_log2", &Tuning::streamk_max_tiles_log2, nullptr}, {"streamk_max_steps", &Tuning::streamk_max_steps, nullptr}, {"store_tile", &Tuning::store_tile, nullptr}, {"ws_min_k", &Tuning::ws_min_k, nullptr}, {"t256_ctas_log2", &Tuning::t256_ctas_log2, nullptr},
    {"permute_low_bits", &Tuning::permute_low_bits, nullptr}, {"permute_ctas_per_sm", &Tuning::permute_ctas_per_sm, nullptr},
    {"streamk_fix_us", nullptr, &Tuning::streamk_fix_us}, {"store_group_log2", &Tuning::store_group_log2, nullptr},
};
// "gemm_min_out" sets every k at once, "gemm_min_out.<k>" one entry
int min_out_index(const char* key) {
    const std::string s = key ? key : "";
    if (s == "gemm_min_out") return 17;
    if (s.rfind("gemm_min_out.", 0) != 0) return -1;
    const int k = atoi(s.c_str() + 13);
    return (k >= 0 && k <= 16) ? k : -1;
}
}  // namespace

bool tuning_set(const char* key, double value) {
    const int mo = min_out_index(key);
    if (mo == 17) { for (int i = 1; i < 17; i++) tuning().gemm_min_out[i] = (int)value; return true; }
    if (mo >= 0) { tuning().gemm_min_out[mo] = (int)value; return true; }
    for (const TuneField& f : kTuneFields)
        if (key && std::string(key) == f.key) {
            if (f.i) tuning().*(f.i) = (int)value; else tuning().*(f.d) = value;
            return true;
        }
    return false;
}
bool tuning_get(const char* key, double* value) {
    const int mo = min_out_index(key);
    if (mo >= 0 && mo <= 16) { *value = tuning().gemm_min_out[mo]; return true; }
    for (const TuneField& f : kTuneFields)
        if (key && std::string(key) == f.key) {
            *value = f.i ? (double)(tuning().*(f.i)) : tuning().*(f.d);
            return true;
        }
    return false;
}

// A GEMM launch of `blocks` CTAs on `slots` CTA slots: full waves run with every SM shared by its resident CTAs;
// a last partial wave that leaves each CTA an SM to itself runs faster per CTA (alone_frac of the SM's DMMA rate
// instead of 1/2), which is what makes a short grid less bad than its wave count says and a split that merely fills
// more slots less good.  Split-K adds the partials' traffic (written once, read once) and the reduce launch.
double gemm_time_model_us(int m, int n, int k, int tm_log2, int tn_log2, int c) {
    const Tuning& T = tuning();
    const int per_sm = (tm_log2 == 7 && tn_log2 == 6) ? 2 : 1;
    const double sms = kNumSMs, slots = sms * per_sm;
    const double blocks = std::ldexp(1.0, (m - tm_log2) + (n - tn_log2) + c);
    const double ksteps = std::max(1.0, std::ldexp(1.0, k - c - 4));  // K steps of 16
    const double step_flops = 2.0 * std::ldexp(1.0, tm_log2 + tn_log2 + 4);
    const double sm_flops_per_us = T.sm_gflops * 1e3;
    const double step_shared = step_flops / (sm_flops_per_us / per_sm);
    const double step_alone = step_flops / (sm_flops_per_us * std::min(1.0, per_sm == 2 ? T.alone_frac : 1.0));
    const double full = std::floor(blocks / slots), rem = blocks - full * slots;
    double t = full * ksteps * step_shared;
    if (rem > 0) t += ksteps * (rem <= sms ? step_alone : step_shared);
    t += T.gemm_fix_us;
    if (c > 0) t += 16.0 * std::ldexp(1.0, m + n + c) / (T.reduce_gbs * 1e3) + T.reduce_fix_us;
    return t;
}

// Stream-K: every CTA slot carries the same number of K steps (ceil(tiles * KT / slots)), whatever the tile count; the
// price is one partial 128x64 tile written and read per CTA plus the owners' waits (streamk_fix_us, fitted).  Eligible:
// 128x64 tiles, K >= 256 (the warp-specialised pipeline), 64..256 tiles: fewer and an owner sums too many partials (32
// tiles: 9 each, measured 6 % slower than split-K); more and the one-tile-per-CTA grid is already balanced by the block
// scheduler while stream-K ranges spread over the whole tile space lose the raster's L2 locality (512 tiles: 13-25 %
// slower, 1024+: 25-40 % — profiles/r02h_kernel_lab_streamk.md).  And only up to streamk_max_steps K steps per CTA: with a long
// K a power-of-two split of the one-tile-per-CTA grid balances just as well — its partials' round trip and reduce pass are
// amortised over the long K — and the exchange buys nothing (m=11,n=10: k=10 +7 %, k=11 +9 %, k=12 0, k=13 -0.7 %;
// profiles/r02i_kernel_lab_streamk_long.md).
double streamk_time_model_us(int m, int n, int k, int* ctas) {
    const Tuning& T = tuning();
    if (m < 7 || n < 6 || k < 8) return -1.0;
    const int tiles_log2 = (m - 7) + (n - 6);
    if (tiles_log2 < T.streamk_min_tiles_log2 || tiles_log2 > T.streamk_max_tiles_log2 || tiles_log2 + (k - 4) > 40) return -1.0;
    const double slots = 2.0 * kNumSMs;
    if (slots > (kSkFlagBytes / 4 - 1)) return -1.0;
    const double G = std::ldexp(1.0, tiles_log2 + k - 4), KT = std::ldexp(1.0, k - 4);
    if (G < slots) return -1.0;
    const double steps = std::ceil(G / slots);
    if (steps / KT + 2.0 > kSkMaxSegs || steps > T.streamk_max_steps) return -1.0;
    if (ctas) *ctas = (int)slots;
    const double step_shared = 2.0 * std::ldexp(1.0, 7 + 6 + 4) / (T.sm_gflops * 1e3 / 2.0);
    return steps * step_shared + T.gemm_fix_us + T.streamk_fix_us;
}

void choose_kernel(Op* op, int32_t kernel_policy, bool allow_splitk) {
    const Tuning& T = tuning();
    const int m = op->m, n = op->n, k = op->k;
    op->flops = 2.0 * std::ldexp(1.0, m + n + k);
    op->bytes = 8.0 * (std::ldexp(1.0, m + k) + std::ldexp(1.0, n + k) + std::ldexp(1.0, m + n));
    op->ksplit_log2 = 0;
    op->streamk = 0;
    // k >= gemm_min_k: the DMMA pipeline proper.  1 <= k < gemm_min_k with a large two-sided output (outer-product-
    // like joins): the same kernel with a zero-filled K step, i.e. a tiled store kernel with full operand reuse,
    // instead of one thread per output re-reading both rows from L2.  The crossover against the generic kernels is
    // a measured table by k (tob_dispatch_table.h): few outputs with a long K belong to the warp / CTA-per-output
    // kernels, whose K split fills the machine where one or two GEMM tiles cannot.
    const int min_free = k >= T.gemm_min_k ? T.gemm_min_free : T.gemm_smallk_min_free;
    const bool gemm_ok = k >= 1 && m >= min_free && n >= min_free && (m + n) >= T.gemm_min_out[std::min(k, 16)];
    if (kernel_policy != 1 && gemm_ok) {
        op->kind = OP_GEMM;
        op->tm_log2 = std::min(m, 7);
        op->tn_log2 = std::min(n, 6);  // 128x64 tiles, two CTAs per SM (64x64 when m == 6)
        int ks = 0;
        if (allow_splitk) {
            // the power-of-two K split with the shortest modelled time (gemm_time_model_us); ties go to the smaller split
            double best = gemm_time_model_us(m, n, k, op->tm_log2, op->tn_log2, 0);
            for (int c = 1; c <= T.max_ksplit_log2 && (k - c) >= T.min_k_per_split_log2; c++) {
                const double t = gemm_time_model_us(m, n, k, op->tm_log2, op->tn_log2, c);
                if (t < best * 0.97) { best = t; ks = c; }
            }
            if (T.force_ksplit_log2 >= 0) ks = std::max(0, std::min(T.force_ksplit_log2, k - 4));
            if (T.streamk > 0 && op->tm_log2 == 7 && op->tn_log2 == 6) {
                int ctas = 0;
                const double t = streamk_time_model_us(m, n, k, &ctas);
                if (t > 0.0 && (T.streamk >= 2 || (T.force_ksplit_log2 < 0 && t < best * 0.98))) {
                    op->streamk = ctas;
                    ks = 0;
                }
            }
        }
        op->ksplit_log2 = ks;
        return;
    }
    op->kind = OP_GENERIC;
    const int outs = m + n;
    if (k <= ((outs <= T.t1_small_out) ? T.t1_small_max_k : T.t1_max_k)) {
        op->threads_per_out = 1;
    } else if (outs >= T.t32_min_out && k <= T.t32_max_k) {
        op->threads_per_out = 32;
    } else {
        op->threads_per_out = 256;
    }
    if (op->threads_per_out == 256 && allow_splitk) {
        // one CTA per (output, k-chunk); want >= 4 CTAs per SM, chunks of >= 2^12 elements
        int ks = 0;
        while (outs + ks < T.t256_ctas_log2 && (k - ks) > 12) ks++;
        op->ksplit_log2 = ks;
    }
}

// ------------------------------------------------------------------------------------------------
// Offline arena allocator (first fit, lowest address), sizes in doubles
// ------------------------------------------------------------------------------------------------
struct Arena {
    // A free block remembers which joins released its pieces, as the range [tmin, tmax] of their post-order
    // positions (empty range = released before the last barrier: safe for everyone).  A join whose subtree
    // covers that range already depends on those joins, so reusing the block adds no ordering edge to the
    // DAG schedule; any other reuse would serialise two independent subtrees (WAR) — for small tensors
    // the arena grows instead, for large ones memory wins and the false dependency is accepted.
    struct Free { int64_t size; int32_t tmin, tmax; };
    std::map<int64_t, Free> free_;  // offset -> block
    int64_t top = 0;                // high-water mark
    static constexpr int64_t kSmall = (int64_t)1 << 17;  // doubles: below this never take an unrelated block
    static bool safe(const Free& f, int32_t lo, int32_t hi) { return f.tmin > f.tmax || (f.tmin >= lo && f.tmax <= hi); }
    int64_t take(std::map<int64_t, Free>::iterator it, int64_t size) {
        const int64_t off = it->first;
        Free rest = it->second;
        rest.size -= size;
        free_.erase(it);
        if (rest.size > 0) free_[off + size] = rest;
        return off;
    }
    // [lo, hi]: post-order range of the allocating join's subtree (joins it transitively depends on)
    int64_t alloc(int64_t size, int32_t lo = 0, int32_t hi = -1, bool dag = false) {
        size = round_up(std::max<int64_t>(size, 1), kAlign);
        if (dag)
            for (auto it = free_.begin(); it != free_.end(); ++it)
                if (it->second.size >= size && safe(it->second, lo, hi)) return take(it, size);
        if (!dag || size > kSmall)
            for (auto it = free_.begin(); it != free_.end(); ++it)
                if (it->second.size >= size) return take(it, size);
        // extend: if the last free block touches the top (and may be used), grow it
        if (!free_.empty()) {
            auto last = std::prev(free_.end());
            if (last->first + last->second.size == top && (!dag || size > kSmall || safe(last->second, lo, hi))) {
                int64_t off = last->first;
                free_.erase(last);
                top = off + size;
                return off;
            }
        }
        int64_t off = top;
        top += size;
        return off;
    }
    void release(int64_t off, int64_t size, int32_t tag = -1) {
        size = round_up(std::max<int64_t>(size, 1), kAlign);
        Free f{size, tag < 0 ? 1 : tag, tag < 0 ? 0 : tag};
        auto merge = [](Free& a, const Free& b) {
            a.size += b.size;
            if (b.tmin <= b.tmax) {
                if (a.tmin > a.tmax) { a.tmin = b.tmin; a.tmax = b.tmax; }
                else { a.tmin = std::min(a.tmin, b.tmin); a.tmax = std::max(a.tmax, b.tmax); }
            }
        };
        auto it = free_.emplace(off, f).first;
        auto nx = std::next(it);
        if (nx != free_.end() && it->first + it->second.size == nx->first) {
            merge(it->second, nx->second);
            free_.erase(nx);
        }
        if (it != free_.begin()) {
            auto pv = std::prev(it);
            if (pv->first + pv->second.size == it->first) {
                merge(pv->second, it->second);
                free_.erase(it);
            }
        }
    }
    void barrier() {  // everything released so far is ordered before whatever comes next
        for (auto& kv : free_) { kv.second.tmin = 1; kv.second.tmax = 0; }
    }
};

// ------------------------------------------------------------------------------------------------
// DAG schedule of one op list.  The contraction tree's joins form a forest of dependencies of depth
// ~20 for trees of 200+ joins (line-graph plans are caterpillars hanging off a few long spines), but one
// stream runs them as a chain and every launch-bound join then costs a full dependent-launch latency.
// Ops are spread over `max_branches` streams: an op follows one of its producers on that producer's
// stream when it can, independent subtrees start on a free stream, and every other ordering the
// sequential program relied on becomes an explicit wait:
//   RAW  the producers of both operands,
//   WAR/WAW  every earlier op that read or wrote arena space the result overwrites (the offline arena
//        reuses space along the post-order),
//   the split-K workspace (one per lane) serialises its users.
// OP_MICRO and OP_ACCUM are barriers (the executor joins all branches before them and forks after).
// ------------------------------------------------------------------------------------------------
int schedule_branches(std::vector<Op>* list_p, int max_branches) {
    std::vector<Op>& list = *list_p;
    const int n = (int)list.size();
    for (Op& op : list) { op.branch = 0; op.signal = 0; op.waits.clear(); }
    if (max_branches <= 1 || n < 3) return 1;
    const int B = std::min(max_branches, 32);
    int32_t max_node = 0;
    for (const Op& op : list) max_node = std::max(max_node, op.node);
    std::vector<int32_t> idx_of_node((size_t)max_node + 1, -1);
    for (int j = 0; j < n; j++)
        if (list[j].kind == OP_GENERIC || list[j].kind == OP_GEMM) idx_of_node[list[j].node] = j;
    auto producer = [&](const OperandRef& r) -> int {
        if (r.space == 0 || r.node < 0 || r.node > max_node) return -1;
        return idx_of_node[r.node];
    };
    std::vector<int32_t> consumer(n, -1);
    for (int j = 0; j < n; j++) {
        if (list[j].kind != OP_GENERIC && list[j].kind != OP_GEMM) continue;
        for (const OperandRef* r : {&list[j].a, &list[j].b}) {
            const int pi = producer(*r);
            if (pi >= 0) consumer[pi] = j;
        }
    }
    struct Rec { int64_t lo, hi; int32_t op; };
    std::vector<Rec> recs;          // arena regions touched by earlier ops (since the last barrier)
    std::vector<int32_t> tail(B, -1);                     // last op of each branch
    std::vector<char> is_free(B, 1);                      // tail already consumed elsewhere (or unused)
    std::vector<std::vector<int32_t>> waited(B, std::vector<int32_t>(B, -1));
    int barrier = -1, last_ws = -1, used = 1;
    std::vector<int32_t> deps;
    for (int j = 0; j < n; j++) {
        Op& op = list[j];
        if (op.kind != OP_GENERIC && op.kind != OP_GEMM) {  // barrier: joins everything, runs on branch 0
            barrier = j;
            recs.clear();
            last_ws = -1;
            for (int b = 0; b < B; b++) { tail[b] = -1; is_free[b] = 1; }
            tail[0] = j;
            continue;
        }
        deps.clear();
        const int pa = producer(op.a), pb = producer(op.b);
        if (pa > barrier) deps.push_back(pa);
        if (pb > barrier) deps.push_back(pb);
        const int64_t lo = op.c_offset, hi = op.c_offset + ((int64_t)1 << (op.m + op.n));
        size_t keep = 0;
        for (size_t r = 0; r < recs.size(); r++) {
            const Rec& R = recs[r];
            const bool overlap = R.lo < hi && lo < R.hi;
            if (overlap) deps.push_back(R.op);
            if (!(overlap && R.lo >= lo && R.hi <= hi)) recs[keep++] = R;  // fully overwritten records are superseded
        }
        recs.resize(keep);
        if (op.ksplit_log2 > 0 || op.streamk > 0) {  // the lane's one workspace (and its stream-K flags)
            if (last_ws > barrier) deps.push_back(last_ws);
            last_ws = j;
        }
        // branch: directly behind a producer whose branch has not moved on, else a free one, else the stalest
        int br = -1;
        for (int pi : {pa, pb})
            if (br < 0 && pi > barrier && tail[list[pi].branch] == pi) br = list[pi].branch;
        // a stream never used since the barrier first: behind a finished chain the new one would inherit that
        // chain's stream order (a false dependency, also in the captured graph); then the free stream whose
        // tail is oldest
        if (br < 0)
            for (int b = 0; b < B && br < 0; b++)
                if (is_free[b] && tail[b] < 0) br = b;
        if (br < 0)
            for (int b = 0; b < B; b++)
                if (is_free[b] && (br < 0 || tail[b] < tail[br])) br = b;
        if (br < 0) {
            br = 0;
            for (int b = 1; b < B; b++)
                if (tail[b] < tail[br]) br = b;
        }
        op.branch = br;
        used = std::max(used, br + 1);
        for (int pi : {pa, pb})
            if (pi > barrier && list[pi].branch != br && tail[list[pi].branch] == pi) is_free[list[pi].branch] = 1;
        tail[br] = j;
        is_free[br] = consumer[j] < 0 ? 1 : 0;
        std::sort(deps.begin(), deps.end());
        deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
        for (int i : deps) {
            const int bi = list[i].branch;
            if (bi == br || i <= waited[br][bi]) continue;  // stream order / an earlier wait already covers it
            op.waits.push_back(i);
            list[i].signal = 1;
            waited[br][bi] = i;
        }
        for (const OperandRef* r : {&op.a, &op.b})
            if (r->space != 0) {
                const int rank = (r == &op.a ? op.m : op.n) + op.k;
                recs.push_back(Rec{r->offset, r->offset + ((int64_t)1 << rank), j});
            }
        recs.push_back(Rec{lo, hi, j});
    }
    return used;
}

static bool contains(const std::vector<int32_t>& sorted, int32_t e) {
    return std::binary_search(sorted.begin(), sorted.end(), e);
}

int compile(const tob_plan_desc* d, const tob_options* opt_in, Program* P, std::string* err) {
    auto fail = [&](const std::string& m) { *err = m; return TOB_E_INVALID; };
    if (!d || d->n_nodes <= 0 || d->n_leaves < 0) return fail("empty plan");
    if (d->n_slice_groups < 0 || d->n_slice_groups > 62) return fail("n_slice_groups out of range");
    tob_options opt;
    if (opt_in) opt = *opt_in; else tob_default_options(&opt);
    P->opt = opt;
    P->n_slice_groups = d->n_slice_groups;
    P->src_leaf_len = d->leaf_data_len;
    const int S = d->n_slice_groups;

    static const bool trace_compile = getenv("TOB_TRACE_COMPILE") != nullptr;
    auto tnow = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tmark = tnow();
    auto lap = [&](const char* what) {
        if (!trace_compile) return;
        const double t = tnow();
        fprintf(stderr, "[tob compile] %-28s %8.1f us\n", what, t - tmark);
        tmark = t;
    };
    // ---- leaves ----
    P->leaves.resize(d->n_leaves);
    for (int l = 0; l < d->n_leaves; l++) {
        LeafInfo& L = P->leaves[l];
        L.rank = d->leaf_rank[l];
        if (L.rank < 0 || L.rank > 30) return fail("leaf rank out of range");
        if (d->leaf_axis_start[l + 1] - d->leaf_axis_start[l] != L.rank) return fail("leaf axis table inconsistent");
        L.src_offset = d->leaf_data_offset[l];
        if (L.src_offset < 0 || L.src_offset + ((int64_t)1 << L.rank) > d->leaf_data_len)
            return fail("leaf data offset out of range");
        L.axis_edge.assign(d->leaf_axis_edge + d->leaf_axis_start[l], d->leaf_axis_edge + d->leaf_axis_start[l + 1]);
        for (int e : L.axis_edge)
            if (e < 0 && -(e + 1) >= S) return fail("leaf axis names a slice group that does not exist");
    }

    lap("leaves");
    // ---- nodes: structure + edge sets (bottom-up) ----
    const int N = d->n_nodes;
    P->nodes.assign(N, NodeInfo());
    std::vector<int> used(N, 0), leaf_used(d->n_leaves, 0);
    for (int i = 0; i < N; i++) {
        NodeInfo& X = P->nodes[i];
        X.leaf = d->node_leaf[i];
        X.left = d->node_left[i];
        X.right = d->node_right[i];
        if (X.leaf >= 0) {
            if (X.leaf >= d->n_leaves) return fail("node names a leaf that does not exist");
            if (leaf_used[X.leaf]++) return fail("leaf table entry used by two nodes (duplicate it instead)");
            const LeafInfo& L = P->leaves[X.leaf];
            for (int e : L.axis_edge) {
                if (e >= 0) X.edges.push_back(e);
                else X.slice_dependent = true;
            }
            std::sort(X.edges.begin(), X.edges.end());
            if (std::adjacent_find(X.edges.begin(), X.edges.end()) != X.edges.end())
                return fail("a leaf carries the same edge on two axes (self loop)");
        } else {
            if (X.left < 0 || X.right < 0 || X.left >= i || X.right >= i || X.left == X.right)
                return fail("nodes are not in post-order");
            if (used[X.left]++ || used[X.right]++) return fail("a node has two parents");
            NodeInfo& A = P->nodes[X.left];
            NodeInfo& B = P->nodes[X.right];
            A.parent = i;
            B.parent = i;
            X.edges.reserve(A.edges.size() + B.edges.size());
            std::set_symmetric_difference(A.edges.begin(), A.edges.end(), B.edges.begin(), B.edges.end(),
                                          std::back_inserter(X.edges));
            X.slice_dependent = A.slice_dependent || B.slice_dependent;
        }
        if ((int)X.edges.size() > kMaxRank) return fail("intermediate tensor rank exceeds the supported maximum");
    }
    for (int i = 0; i < N - 1; i++)
        if (!used[i]) return fail("plan is a forest: a non-root node has no parent");
    if (!P->nodes[N - 1].edges.empty())
        return fail("root tensor is not rank 0 (the network has open indices)");

    lap("nodes + edge sets");
    // ---- canonical layouts (top-down) ----
    int32_t max_edge = -1;
    for (const NodeInfo& X : P->nodes)
        for (int32_t e : X.edges) max_edge = std::max(max_edge, e);
    std::vector<int32_t> stamp((size_t)max_edge + 2, -1);  // stamp[e] == tag  <=>  e in the tagged node's edge set
    for (int i = N - 1; i >= 0; i--) {
        NodeInfo& X = P->nodes[i];
        if (X.leaf >= 0) continue;
        NodeInfo& A = P->nodes[X.left];
        NodeInfo& B = P->nodes[X.right];
        std::vector<int32_t> K;
        K.reserve(std::min(A.edges.size(), B.edges.size()));
        std::set_intersection(A.edges.begin(), A.edges.end(), B.edges.begin(), B.edges.end(), std::back_inserter(K));
        for (int side = 0; side < 2; side++) {
            NodeInfo* C = side ? &B : &A;
            const int32_t tag = side ? X.right : X.left;
            for (int32_t e : C->edges) stamp[e] = tag;
            C->layout.reserve(C->edges.size());
            C->layout = K;  // ascending edge id, identical for both siblings
            C->k_with_sibling = (int)K.size();
            for (int32_t e : X.layout)
                if (stamp[e] == tag) C->layout.push_back(e);
            if (C->layout.size() != C->edges.size()) return fail("internal: layout does not cover the node's edges");
        }
    }

    lap("layouts");
    // ---- leaves: device placement, upload permutation, slice terms ----
    int64_t leaf_top = 0;
    for (int i = 0; i < N; i++) {
        NodeInfo& X = P->nodes[i];
        if (X.leaf < 0) continue;
        LeafInfo& L = P->leaves[X.leaf];
        L.live_rank = (int)X.edges.size();
        L.dev_offset = leaf_top;
        leaf_top += round_up((int64_t)1 << L.rank, kAlign);
        L.src_bit.assign(L.rank, -1);
        for (int p = 0; p < L.live_rank; p++) {
            int axis = -1;
            for (int j = 0; j < L.rank; j++)
                if (L.axis_edge[j] == X.layout[p]) axis = j;
            L.src_bit[p] = L.rank - 1 - axis;
        }
        int p = L.live_rank;
        for (int j = 0; j < L.rank; j++) {
            int e = L.axis_edge[j];
            if (e >= 0) continue;
            int g = -(e + 1);
            L.src_bit[p] = L.rank - 1 - j;
            L.slice_id_bit.push_back(S - 1 - g);
            L.slice_addr_bit.push_back(p);
            p++;
        }
        X.where.space = 0;
        X.where.offset = L.dev_offset;
        X.where.leaf = L.slice_id_bit.empty() ? -1 : X.leaf;
        X.where.node = i;
    }
    P->leaf_doubles = leaf_top;

    lap("leaf placement");
    // ---- ops (post-order), micro subtrees, hoisting, arena ----
    const bool hoist = opt.hoist_invariant && S > 0;
    const int max_branches = opt.dag_branches == 0 ? 16 : opt.dag_branches;
    const bool dag = max_branches > 1;
    Arena arena;
    int64_t ws_max = 0;
    std::vector<int64_t> size_of(N, 0);

    auto build_op = [&](int i) {
        NodeInfo& X = P->nodes[i];
        NodeInfo& A = P->nodes[X.left];
        NodeInfo& B = P->nodes[X.right];
        Op op;
        op.node = i;
        op.k = A.k_with_sibling;
        op.m = (int)A.edges.size() - op.k;
        op.n = (int)B.edges.size() - op.k;
        op.a = A.where;
        op.b = B.where;
        uint64_t mask = 0;
        for (int32_t e : A.edges) stamp[e] = -2 - i;  // unique tag per join
        for (size_t p = 0; p < X.layout.size(); p++)
            if (stamp[X.layout[p]] == -2 - i) mask |= (uint64_t)1 << p;
        op.mask_m = mask;
        if (op.n > op.m) {  // keep the larger free side as M (the GEMM tiles assume m >= n)
            std::swap(op.a, op.b);
            std::swap(op.m, op.n);
            op.mask_m = ~mask & ((op.m + op.n) >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << (op.m + op.n)) - 1));
        }
        return op;
    };

    // ---- micro stages ----
    // A join is "mini" when one CTA finishes it in about a launch latency (operands and result <= 2^14
    // doubles, <= 2^17 multiply-adds).  Line-graph trees are long chains of such joins — a running
    // intermediate absorbing one small tensor after the other — hanging between a few large joins.  Each
    // phase (slice-invariant prologue / per-slice part) is cut into stages: stage L = ONE launch running
    // every mini join of level L, one CTA per CHAIN (each join consumes the previous one's result, handed
    // over in shared memory), followed by the large joins of level L as kernels of their own.  Levels are
    // Strahler numbers: a mini join continues the chain of its highest-level mini child; two mini children
    // of equal level end both chains and start a new one a level up; a large child pushes its consumer one
    // level up; a large join runs after the stage of its highest child.  Chains of one stage are therefore
    // independent, everything a stage reads was produced by an earlier stage (or earlier in its own chain),
    // and the number of stages grows with the tree's branching depth (2-5), not with its size.
    static const int kMiniRank = getenv("TOB_MINI_RANK") ? std::min(14, std::max(4, atoi(getenv("TOB_MINI_RANK")))) : 14;  // experiments
    const int kMiniWork = kMiniRank + 3;
    std::vector<char> mini(N, 0), phase(N, 1);
    std::vector<int32_t> lvl(N, 0), subtree_nodes(N, 1), frag_root(N, -1);  // frag_root: first join of the chain
    for (int i = 0; i < N; i++) {
        const NodeInfo& X = P->nodes[i];
        phase[i] = (!hoist || X.slice_dependent) ? 1 : 0;
        if (X.leaf >= 0) continue;
        subtree_nodes[i] = 1 + subtree_nodes[X.left] + subtree_nodes[X.right];
        const int k = P->nodes[X.left].k_with_sibling;
        const int m = (int)P->nodes[X.left].edges.size() - k, n = (int)P->nodes[X.right].edges.size() - k;
        // k <= 4: a thread walks the contracted range of its output alone, 16 terms at most (the dot products
        // near the root have few outputs and a long K: those go to the warp / CTA-per-output kernels)
        mini[i] = (opt.use_microtree && m + k <= kMiniRank && n + k <= kMiniRank && m + n <= kMiniRank &&
                   m + n + k <= kMiniWork && k <= 4) ? 1 : 0;
        int l = 0, n_at = 0, cont = -1;  // highest child level, mini children at that level, one of them
        for (int c : {X.left, X.right}) {
            if (P->nodes[c].leaf >= 0 || phase[c] != phase[i]) continue;  // leaves and hoisted results are just there
            const int e = lvl[c] + ((mini[i] && !mini[c]) ? 1 : 0);
            if (e > l) { l = e; n_at = 0; cont = -1; }
            if (e == l && mini[c]) { n_at++; cont = c; }
        }
        if (!mini[i]) { lvl[i] = l; continue; }
        if (n_at == 2) { lvl[i] = l + 1; frag_root[i] = i; }
        else { lvl[i] = l; frag_root[i] = (n_at == 1 && lvl[cont] == l) ? frag_root[cont] : i; }
    }
    std::vector<char> persistent(N, 0);
    if (hoist) {
        for (int i = 0; i < N; i++) {
            const NodeInfo& X = P->nodes[i];
            if (X.leaf < 0 && !phase[i] && X.parent >= 0 && phase[X.parent]) persistent[i] = 1;
        }
    }

    auto place = [&](Op& op, int i) {
        NodeInfo& X = P->nodes[i];
        size_of[i] = (int64_t)1 << (op.m + op.n);
        op.c_offset = arena.alloc(size_of[i], i - subtree_nodes[i] + 1, i, dag);
        X.where.space = persistent[i] ? 2 : 1;  // hoisted results are read by every lane from lane 0's arena
        X.where.offset = op.c_offset;
        X.where.leaf = -1;
        X.where.node = i;
        P->total_flops += op.flops;
        P->total_bytes += op.bytes;
    };
    auto emit = [&](int i, std::vector<Op>* list) {
        NodeInfo& X = P->nodes[i];
        Op op = build_op(i);
        op.invariant = phase[i] ? 0 : 1;
        choose_kernel(&op, opt.kernel_policy, true);
        place(op, i);
        if (op.ksplit_log2 > 0) {
            op.ws_offset = 0;
            ws_max = std::max(ws_max, round_up(size_of[i] << op.ksplit_log2, kAlign));
        }
        if (op.streamk > 0) {
            op.ws_offset = 0;
            ws_max = std::max(ws_max, round_up((int64_t)op.streamk * kSkSlotDoubles, kAlign));
        }
        for (int c : {X.left, X.right}) {
            const NodeInfo& Cn = P->nodes[c];
            if (Cn.leaf < 0 && !persistent[c]) arena.release(Cn.where.offset, size_of[c], i);
        }
        list->push_back(op);
    };
    auto run_stage = [&](int dep, int level, std::vector<Op>* list) {
        arena.barrier();  // everything released so far is ordered before this stage
        // 1. the mini joins of this level: one launch, one CTA per (packed) chain
        std::vector<std::vector<Op>> frags;
        std::vector<double> work;
        std::vector<int32_t> frag_of_root(N, -1);
        std::vector<std::pair<int64_t, int64_t>> deferred;  // arena space freed only after the launch
        int max_outs_log2 = 0;
        for (int i = 0; i < N; i++) {
            NodeInfo& X = P->nodes[i];
            if (X.leaf >= 0 || !mini[i] || phase[i] != dep || lvl[i] != level) continue;
            int32_t& f = frag_of_root[frag_root[i]];
            if (f < 0) { f = (int32_t)frags.size(); frags.emplace_back(); work.push_back(0.0); }
            Op op = build_op(i);
            op.invariant = dep ? 0 : 1;
            op.kind = OP_GENERIC;
            op.threads_per_out = 1;
            op.flops = 2.0 * std::ldexp(1.0, op.m + op.n + op.k);
            op.bytes = 8.0 * (std::ldexp(1.0, op.m + op.k) + std::ldexp(1.0, op.n + op.k) + std::ldexp(1.0, op.m + op.n));
            place(op, i);
            for (int c : {X.left, X.right}) {
                const NodeInfo& Cn = P->nodes[c];
                if (Cn.leaf < 0 && !persistent[c]) deferred.emplace_back(Cn.where.offset, size_of[c]);
            }
            max_outs_log2 = std::max(max_outs_log2, op.m + op.n);
            // cost model for packing: per-join latency floor + multiply-adds
            work[f] += 2000.0 + std::ldexp(1.0, op.m + op.n + op.k);
            frags[f].push_back(op);
        }
        if (!frags.empty()) {
            P->micro.emplace_back();
            MicroProgram& mp = P->micro.back();
            static const int force_threads = getenv("TOB_MICRO_THREADS") ? atoi(getenv("TOB_MICRO_THREADS")) : 0;  // experiments
            static const int big_from = getenv("TOB_MICRO_BIG_FROM") ? atoi(getenv("TOB_MICRO_BIG_FROM")) : 10;
            mp.threads = force_threads == 256 || force_threads == 1024 ? force_threads : (max_outs_log2 > big_from ? 1024 : 256);
            const int n_cta = (int)std::min<size_t>(frags.size(), 2 * kNumSMs);
            std::vector<size_t> order(frags.size());
            for (size_t j = 0; j < order.size(); j++) order[j] = j;
            std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return work[a] > work[b]; });
            std::vector<std::vector<size_t>> bins(n_cta);
            std::vector<double> load(n_cta, 0.0);
            for (size_t j : order) {  // longest first onto the least loaded CTA
                int best = 0;
                for (int c = 1; c < n_cta; c++)
                    if (load[c] < load[best]) best = c;
                bins[best].push_back(j);
                load[best] += work[j];
            }
            mp.cta_start.reserve(n_cta + 1);
            size_t n_mini = 0;
            for (const auto& fr : frags) n_mini += fr.size();
            mp.ops.reserve(n_mini);
            mp.cta_start.push_back(0);
            for (int c = 0; c < n_cta; c++) {
                if (bins[c].size() > 1) std::sort(bins[c].begin(), bins[c].end());  // chains of one stage are independent; keep tree order
                for (size_t j : bins[c])
                    for (const Op& op : frags[j]) mp.ops.push_back(op);
                mp.cta_start.push_back((int32_t)mp.ops.size());
            }
            Op launch;
            launch.kind = OP_MICRO;
            launch.micro_which = (int32_t)P->micro.size() - 1;
            launch.invariant = dep ? 0 : 1;
            for (const Op& op : mp.ops) { launch.flops += op.flops; launch.bytes += op.bytes; }
            list->push_back(launch);
            for (auto& d : deferred) arena.release(d.first, d.second);
            arena.barrier();  // the launch orders everything released so far
        }
        // 2. the large joins of this level, post-order
        for (int i = 0; i < N; i++)
            if (P->nodes[i].leaf < 0 && !mini[i] && phase[i] == dep && lvl[i] == level) emit(i, list);
    };
    lap("levels");
    P->invariant_ops.reserve(hoist ? N / 4 + 8 : 0);
    P->slice_ops.reserve(N / 4 + 8);
    for (int dep = hoist ? 0 : 1; dep <= 1; dep++) {
        int max_level = -1;
        for (int i = 0; i < N; i++)
            if (P->nodes[i].leaf < 0 && phase[i] == dep) max_level = std::max(max_level, lvl[i]);
        for (int level = 0; level <= max_level; level++) run_stage(dep, level, dep ? &P->slice_ops : &P->invariant_ops);
    }
    lap("stages: ops + arena");
    P->root = P->nodes[N - 1].where;
    Op acc;
    acc.kind = OP_ACCUM;
    acc.node = N - 1;
    acc.a = P->root;
    P->slice_ops.push_back(acc);
    P->arena_doubles = arena.top;
    P->ws_doubles = ws_max;
    // Two slices in flight (own arena + workspace + stream each) when that costs at most 2 GiB extra: the
    // launch-bound stretches of one slice then overlap the GEMMs of the other.  opt.slice_lanes: 0 auto, 1, 2.
    const int want_lanes = opt.slice_lanes;
    bool cheap = (P->arena_doubles + P->ws_doubles) * 8 <= ((int64_t)2 << 30);
    // under a byte budget (tob_options.mem_limit_bytes, e.g. the b200_mem slicer's --mem_limit) the second lane is the
    // first thing to go: it must not be what pushes a plan over the limit and triggers one more slice
    if (opt.mem_limit_bytes > 0 && 8 * (P->leaf_doubles + 2 * (P->arena_doubles + P->ws_doubles)) + (1 << 17) > opt.mem_limit_bytes)
        cheap = false;
    P->lanes = (S > 0 && (want_lanes == 2 || (want_lanes == 0 && cheap))) ? 2 : 1;
    P->branches = std::max(schedule_branches(&P->invariant_ops, max_branches), schedule_branches(&P->slice_ops, max_branches));
    lap("dag schedule");
    return TOB_OK;
}

// ------------------------------------------------------------------------------------------------
std::string describe(const Program& P) {
    std::ostringstream o;
    o.precision(17);
    auto ref = [&](const OperandRef& r) {
        o << "{\"space\":" << r.space << ",\"offset\":" << r.offset << ",\"leaf\":" << r.leaf << ",\"node\":" << r.node << "}";
    };
    std::function<void(const std::vector<Op>&)> ops = [&](const std::vector<Op>& v) {
        o << "[";
        for (size_t i = 0; i < v.size(); i++) {
            const Op& op = v[i];
            if (i) o << ",";
            if (op.kind == OP_MICRO) {
                const MicroProgram& mp = P.micro[op.micro_which];
                o << "{\"kind\":3,\"which\":" << op.micro_which << ",\"threads\":" << mp.threads << ",\"invariant\":" << op.invariant << ",\"flops\":" << op.flops
                  << ",\"bytes\":" << op.bytes << ",\"cta_start\":[";
                for (size_t j = 0; j < mp.cta_start.size(); j++) o << (j ? "," : "") << mp.cta_start[j];
                o << "],\"micro\":";
                ops(mp.ops);
                o << "}";
                continue;
            }
            o << "{\"kind\":" << op.kind << ",\"node\":" << op.node << ",\"a\":";
            ref(op.a);
            o << ",\"b\":";
            ref(op.b);
            o << ",\"c_offset\":" << op.c_offset << ",\"m\":" << op.m << ",\"n\":" << op.n << ",\"k\":" << op.k
              << ",\"mask_m\":" << op.mask_m << ",\"threads_per_out\":" << op.threads_per_out
              << ",\"ksplit_log2\":" << op.ksplit_log2 << ",\"streamk\":" << op.streamk << ",\"tm_log2\":" << op.tm_log2 << ",\"tn_log2\":" << op.tn_log2
              << ",\"invariant\":" << op.invariant << ",\"flops\":" << op.flops << ",\"bytes\":" << op.bytes
              << ",\"branch\":" << op.branch << ",\"signal\":" << op.signal << ",\"waits\":[";
            for (size_t w = 0; w < op.waits.size(); w++) o << (w ? "," : "") << op.waits[w];
            o << "]}";
        }
        o << "]";
    };
    o << "{\"n_slice_groups\":" << P.n_slice_groups << ",\"lanes\":" << P.lanes << ",\"branches\":" << P.branches << ",\"leaf_doubles\":" << P.leaf_doubles
      << ",\"arena_doubles\":" << P.arena_doubles << ",\"ws_doubles\":" << P.ws_doubles
      << ",\"total_flops\":" << P.total_flops << ",\"total_bytes\":" << P.total_bytes << ",\"leaves\":[";
    for (size_t l = 0; l < P.leaves.size(); l++) {
        const LeafInfo& L = P.leaves[l];
        if (l) o << ",";
        o << "{\"rank\":" << L.rank << ",\"live_rank\":" << L.live_rank << ",\"src_offset\":" << L.src_offset
          << ",\"dev_offset\":" << L.dev_offset << ",\"src_bit\":[";
        for (size_t i = 0; i < L.src_bit.size(); i++) o << (i ? "," : "") << L.src_bit[i];
        o << "],\"slice_id_bit\":[";
        for (size_t i = 0; i < L.slice_id_bit.size(); i++) o << (i ? "," : "") << L.slice_id_bit[i];
        o << "],\"slice_addr_bit\":[";
        for (size_t i = 0; i < L.slice_addr_bit.size(); i++) o << (i ? "," : "") << L.slice_addr_bit[i];
        o << "]}";
    }
    o << "],\"invariant_ops\":";
    ops(P.invariant_ops);
    o << ",\"slice_ops\":";
    ops(P.slice_ops);
    o << ",\"root\":";
    ref(P.root);
    o << "}";
    return o.str();
}

}  // namespace tob
