"""Random closed tensor networks with random contraction trees and random slicings (test infrastructure).

Every index has extent 2 and joins exactly two tensors, like the edges of the reference's `cnf_count`
networks (src/tensor_network/tensor_network.pyx:11-49).  `make` returns a `FlatPlan` — what
`flatten_plan` hands to `tob_plan_create` — and the expected value of the whole contraction computed with
numpy alone (sum over every index), which is also the sum over all slices whatever the slicing
(BaseTensorAPI.contract_sliced, base_api.py:21-28)."""
import numpy as np

from tensororder_b200.flatten import FlatPlan


def _tree(rng, leaves, shape):
    """shape: 'random' (uniformly random merges), 'caterpillar' (a spine absorbing one leaf at a time, what the
    line-graph planners emit), 'balanced'."""
    nodes = [("leaf", t) for t in leaves]
    if shape == "caterpillar":
        order = list(rng.permutation(len(nodes)))
        cur = nodes[order[0]]
        for j in order[1:]:
            cur = ("join", cur, nodes[j]) if rng.random() < 0.5 else ("join", nodes[j], cur)
        return cur
    pool = list(nodes)
    while len(pool) > 1:
        if shape == "balanced":
            rng.shuffle(pool)
            nxt = [("join", pool[i], pool[i + 1]) for i in range(0, len(pool) - 1, 2)]
            if len(pool) % 2:
                nxt.append(pool[-1])
            pool = nxt
        else:
            i, j = sorted(rng.choice(len(pool), size=2, replace=False))
            b = pool.pop(j)
            a = pool.pop(i)
            pool.append(("join", a, b))
    return pool[0]


def make(seed, n_tensors=10, n_edges=16, n_slice_groups=0, shape="random", integer=True, max_rank=6):
    rng = np.random.default_rng(seed)
    # edges: pairs of distinct tensors (multi-edges allowed, self loops not), bounded tensor rank
    rank = [0] * n_tensors
    pairs = []
    # a spanning path first so the network is connected (scalar factors otherwise: legal, but less of a test)
    perm = list(rng.permutation(n_tensors))
    for a, b in zip(perm, perm[1:]):
        pairs.append((int(a), int(b)))
        rank[a] += 1
        rank[b] += 1
    tries = 0
    while len(pairs) < n_edges and tries < 1000:
        tries += 1
        a, b = (int(x) for x in rng.choice(n_tensors, size=2, replace=False))
        if rank[a] >= max_rank or rank[b] >= max_rank:
            continue
        pairs.append((a, b))
        rank[a] += 1
        rank[b] += 1
    n_edges = len(pairs)
    axes = [[] for _ in range(n_tensors)]  # tensor -> edge ids, in (shuffled) axis order
    for e, (a, b) in enumerate(pairs):
        axes[a].append(e)
        axes[b].append(e)
    for t in range(n_tensors):
        rng.shuffle(axes[t])
    data = []
    for t in range(n_tensors):
        shp = (2,) * len(axes[t])
        if integer:
            data.append(rng.integers(0, 3, size=shp).astype(np.float64))
        else:
            data.append(rng.uniform(0.25, 1.5, size=shp))
    # expected: plain pairwise numpy contraction in index order of the tensors (exact for small integers)
    import string
    alphabet = string.ascii_letters
    assert n_edges <= len(alphabet)
    spec = ",".join("".join(alphabet[e] for e in axes[t]) for t in range(n_tensors)) + "->"
    expected = float(np.einsum(spec, *data, optimize="greedy"))
    # slicing: one edge per group (the reference's groups hold the edges of one CNF variable's copy tensor,
    # which are equal anyway; independent edges fixed jointly would drop the cross terms)
    assert n_slice_groups <= n_edges
    edge_ids = list(rng.permutation(n_edges))
    group_of = {int(edge_ids[g]): g for g in range(n_slice_groups)}
    tree = _tree(rng, list(range(n_tensors)), shape)
    node_left, node_right, node_leaf = [], [], []
    leaf_rank, leaf_off, axis_start, axis_edge = [], [], [0], []
    chunks, total = [], 0

    def walk(node):
        nonlocal total
        if node[0] == "leaf":
            t = node[1]
            node_left.append(-1)
            node_right.append(-1)
            node_leaf.append(len(leaf_rank))
            leaf_rank.append(len(axes[t]))
            leaf_off.append(total)
            chunks.append(data[t].reshape(-1))
            total += data[t].size
            axis_edge.extend(-(group_of[e] + 1) if e in group_of else e for e in axes[t])
            axis_start.append(len(axis_edge))
        else:
            l = walk(node[1])
            r = walk(node[2])
            node_left.append(l)
            node_right.append(r)
            node_leaf.append(-1)
        return len(node_left) - 1

    walk(tree)
    flat = FlatPlan(
        node_left=np.asarray(node_left, dtype=np.int32), node_right=np.asarray(node_right, dtype=np.int32),
        node_leaf=np.asarray(node_leaf, dtype=np.int32), leaf_rank=np.asarray(leaf_rank, dtype=np.int32),
        leaf_data_offset=np.asarray(leaf_off, dtype=np.int64), leaf_axis_start=np.asarray(axis_start, dtype=np.int32),
        leaf_axis_edge=np.asarray(axis_edge, dtype=np.int32),
        leaf_data=np.ascontiguousarray(np.concatenate(chunks), dtype=np.float64),
        n_slice_groups=n_slice_groups, leaf_tensor_index=list(range(n_tensors)))
    return flat, expected
