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

#include "tob_internal.h"

namespace tob {

static const int kMaxRank = 40;       // 2^40 doubles is far beyond 180 GB; guards the bit math
static const int64_t kAlign = 32;     // arena alignment in doubles (256 B)
static const int kNumSMs = 148;

static int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// Kernel choice for a canonical join  C[2^(m+n)] = A[2^m x 2^k] . B[2^n x 2^k]^T
// ------------------------------------------------------------------------------------------------
int gemm_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TOB_GEMM_VARIANT");
        v = e ? atoi(e) : 3;
    }
    return v;
}

void choose_kernel(Op* op, int32_t kernel_policy, bool allow_splitk) {
    const int m = op->m, n = op->n, k = op->k;
    op->flops = 2.0 * std::ldexp(1.0, m + n + k);
    op->bytes = 8.0 * (std::ldexp(1.0, m + k) + std::ldexp(1.0, n + k) + std::ldexp(1.0, m + n));
    op->ksplit_log2 = 0;
    // k >= 4: the DMMA pipeline proper.  1 <= k <= 3 with a large two-sided output (outer-product-like
    // joins): the same kernel with a zero-filled K step, i.e. a tiled store kernel with full operand
    // reuse, instead of one thread per output re-reading both rows from L2.
    const bool gemm_ok = (k >= 4 && m >= 6 && n >= 6 && (m + n + k) >= 20) || (k >= 1 && k <= 3 && m >= 7 && n >= 7 && (m + n) >= 18);
    if (kernel_policy != 1 && gemm_ok) {
        op->kind = OP_GEMM;
        op->tm_log2 = std::min(m, 7);
        op->tn_log2 = std::min(n, 7);
        if (gemm_variant() == 3 && op->tn_log2 == 7) op->tn_log2 = 6;
        int64_t tiles = (int64_t)1 << ((m - op->tm_log2) + (n - op->tn_log2));
        int ks = 0;
        // concurrent CTA slots: the 128x64 kernels run two CTAs per SM
        const int slots = kNumSMs * ((op->tm_log2 == 7 && op->tn_log2 == 6) ? 2 : 1);
        if (allow_splitk && tiles < 8 * slots) {
            // short grid: pick the power-of-two K split with the best wave efficiency (blocks / SMs
            // rounded up), keeping >= 128 K elements per split; ties go to the smaller split
            // ... and keep the split-K partials (written once, read once: 16 * 2^(m+n+c) bytes at ~5 TB/s)
            // below ~15 % of the join's own tensor-pipe time: 2^c <= 0.0027 * 2^k
            double best = -1.0;
            for (int c = 0; (k - c) >= 7 && c <= 6 && std::ldexp(1.0, c) <= std::max(1.0, 0.0027 * std::ldexp(1.0, k)); c++) {
                const double blocks = (double)(tiles << c);
                const double waves = blocks / slots;
                const double eff = waves / std::ceil(waves) - 0.004 * c;
                if (eff > best + 1e-9) { best = eff; ks = c; }
            }
        }
        op->ksplit_log2 = ks;
        return;
    }
    op->kind = OP_GENERIC;
    const int outs = m + n;
    if (k <= 6) {
        op->threads_per_out = 1;
    } else if (outs >= 12 && k <= 11) {
        op->threads_per_out = 32;
    } else {
        op->threads_per_out = 256;
    }
    if (op->threads_per_out == 256 && allow_splitk) {
        // one CTA per (output, k-chunk); want >= 4 CTAs per SM, chunks of >= 2^12 elements
        int ks = 0;
        while (outs + ks < 10 && (k - ks) > 12) ks++;
        op->ksplit_log2 = ks;
    }
}

// ------------------------------------------------------------------------------------------------
// Offline arena allocator (first fit, lowest address), sizes in doubles
// ------------------------------------------------------------------------------------------------
struct Arena {
    std::map<int64_t, int64_t> free_;  // offset -> size
    int64_t top = 0;                   // high-water mark
    int64_t alloc(int64_t size) {
        size = round_up(std::max<int64_t>(size, 1), kAlign);
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second >= size) {
                int64_t off = it->first;
                int64_t rest = it->second - size;
                free_.erase(it);
                if (rest > 0) free_[off + size] = rest;
                return off;
            }
        }
        // extend: if the last free block touches the top, grow it
        if (!free_.empty()) {
            auto last = std::prev(free_.end());
            if (last->first + last->second == top) {
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
    void release(int64_t off, int64_t size) {
        size = round_up(std::max<int64_t>(size, 1), kAlign);
        auto it = free_.emplace(off, size).first;
        auto nx = std::next(it);
        if (nx != free_.end() && it->first + it->second == nx->first) {
            it->second += nx->second;
            free_.erase(nx);
        }
        if (it != free_.begin()) {
            auto pv = std::prev(it);
            if (pv->first + pv->second == it->first) {
                pv->second += it->second;
                free_.erase(it);
            }
        }
    }
};

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

    // ---- ops (post-order), micro subtrees, hoisting, arena ----
    const bool hoist = opt.hoist_invariant && S > 0;
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

    // micro-closed subtrees
    std::vector<char> closed(N, 0), is_root(N, 0), eff_dep(N, 0);
    std::vector<int32_t> subtree_nodes(N, 1);
    for (int i = 0; i < N; i++) {
        const NodeInfo& X = P->nodes[i];
        eff_dep[i] = X.slice_dependent ? 1 : 0;
        if (X.leaf >= 0) { closed[i] = 1; continue; }
        subtree_nodes[i] = 1 + subtree_nodes[X.left] + subtree_nodes[X.right];
        const int k = P->nodes[X.left].k_with_sibling;
        const int m = (int)P->nodes[X.left].edges.size() - k, n = (int)P->nodes[X.right].edges.size() - k;
        const bool tiny = (m + k <= 12 && n + k <= 12 && m + n <= 12 && m + n + k <= 15);
        closed[i] = (opt.use_microtree && tiny && closed[X.left] && closed[X.right]) ? 1 : 0;
    }
    for (int i = 0; i < N; i++) {
        const NodeInfo& X = P->nodes[i];
        if (X.leaf >= 0 || !closed[i]) continue;
        is_root[i] = (X.parent < 0 || !closed[X.parent]) ? 1 : 0;
    }
    // every join of a micro subtree runs in the phase of the subtree's root
    for (int i = N - 1; i >= 0; i--) {
        const NodeInfo& X = P->nodes[i];
        if (X.leaf >= 0 || !closed[i]) continue;
        if (!is_root[i]) eff_dep[i] = eff_dep[X.parent];
    }
    std::vector<char> persistent(N, 0);
    if (hoist) {
        for (int i = 0; i < N; i++) {
            const NodeInfo& X = P->nodes[i];
            if (X.leaf < 0 && !eff_dep[i] && X.parent >= 0 && eff_dep[X.parent]) persistent[i] = 1;
        }
    }

    auto place = [&](Op& op, int i) {
        NodeInfo& X = P->nodes[i];
        size_of[i] = (int64_t)1 << (op.m + op.n);
        op.c_offset = arena.alloc(size_of[i]);
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
        op.invariant = eff_dep[i] ? 0 : 1;
        choose_kernel(&op, opt.kernel_policy, true);
        place(op, i);
        if (op.ksplit_log2 > 0) {
            op.ws_offset = 0;
            ws_max = std::max(ws_max, round_up(size_of[i] << op.ksplit_log2, kAlign));
        }
        for (int c : {X.left, X.right}) {
            const NodeInfo& Cn = P->nodes[c];
            if (Cn.leaf < 0 && !persistent[c]) arena.release(Cn.where.offset, size_of[c]);
        }
        list->push_back(op);
    };
    auto run_phase = [&](int dep, std::vector<Op>* list, int which) {
        // 1. all micro subtrees of this phase: one launch, one CTA per (packed) subtree
        MicroProgram& mp = P->micro[which];
        std::vector<std::vector<Op>> subtrees;
        std::vector<double> work;
        std::vector<std::pair<int64_t, int64_t>> deferred;  // arena space freed only after the launch
        for (int r = 0; r < N; r++) {
            if (!is_root[r] || eff_dep[r] != dep) continue;
            subtrees.emplace_back();
            work.push_back(0.0);
            for (int i = r - subtree_nodes[r] + 1; i <= r; i++) {
                NodeInfo& X = P->nodes[i];
                if (X.leaf >= 0) continue;
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
                // cost model for packing: per-join latency floor + multiply-adds
                work.back() += 2000.0 + std::ldexp(1.0, op.m + op.n + op.k);
                subtrees.back().push_back(op);
            }
        }
        if (!subtrees.empty()) {
            const int n_cta = (int)std::min<size_t>(subtrees.size(), 2 * kNumSMs);
            std::vector<size_t> order(subtrees.size());
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
            mp.cta_start.push_back(0);
            for (int c = 0; c < n_cta; c++) {
                for (size_t j : bins[c])
                    for (const Op& op : subtrees[j]) mp.ops.push_back(op);
                mp.cta_start.push_back((int32_t)mp.ops.size());
            }
            Op launch;
            launch.kind = OP_MICRO;
            launch.micro_which = which;
            launch.invariant = dep ? 0 : 1;
            for (const Op& op : mp.ops) { launch.flops += op.flops; launch.bytes += op.bytes; }
            list->push_back(launch);
            for (auto& d : deferred) arena.release(d.first, d.second);
        }
        // 2. the other joins of this phase, post-order
        for (int i = 0; i < N; i++)
            if (P->nodes[i].leaf < 0 && !closed[i] && eff_dep[i] == dep) emit(i, list);
    };
    if (hoist) {
        run_phase(0, &P->invariant_ops, 0);
        run_phase(1, &P->slice_ops, 1);
    } else {
        for (int i = 0; i < N; i++) eff_dep[i] = 1;
        run_phase(1, &P->slice_ops, 1);
    }
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
    const bool cheap = (P->arena_doubles + P->ws_doubles) * 8 <= ((int64_t)2 << 30);
    P->lanes = (S > 0 && (want_lanes == 2 || (want_lanes == 0 && cheap))) ? 2 : 1;
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
                o << "{\"kind\":3,\"which\":" << op.micro_which << ",\"invariant\":" << op.invariant << ",\"flops\":" << op.flops
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
              << ",\"ksplit_log2\":" << op.ksplit_log2 << ",\"tm_log2\":" << op.tm_log2 << ",\"tn_log2\":" << op.tn_log2
              << ",\"invariant\":" << op.invariant << ",\"flops\":" << op.flops << ",\"bytes\":" << op.bytes << "}";
        }
        o << "]";
    };
    o << "{\"n_slice_groups\":" << P.n_slice_groups << ",\"lanes\":" << P.lanes << ",\"leaf_doubles\":" << P.leaf_doubles
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
