"""Flattens a sliced execution plan (reference objects or a stored plan) into the C arrays of
`tob_plan_desc` (include/tob200.h).

Reads, by duck typing, exactly what the reference's own backends read:
  * `plan.tree.iterate_postorder()`, node `.is_leaf` / `.tensor_index`
    (src/contraction_methods/contraction_tree.pyx:13-27,174-183);
  * `plan.network.index_list(t)`, `plan.network[t].build(factory)`, `.shape`
    (src/tensor_network/tensor_network.pyx:23-39, src/tensor_network/tensor.py:33-34);
  * `plan.groups_to_slice` (src/tensor_network/sliced_execution_plan.py:17-18).
Slicing follows the JAX backend's scheme (jax_apis.py:253-277): the sliced axes leave the tree
(`remove_sliced_indices_from`, tensor_network.pyx:444-468) and each leaf is indexed by the slice
assignment (`get_tensor_slices`, tensor_network.pyx:404-442); empty groups are skipped exactly as
`slice_groups` skips them (tensor_network.pyx:371-373)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class FlatPlan:
    node_left: np.ndarray
    node_right: np.ndarray
    node_leaf: np.ndarray
    leaf_rank: np.ndarray
    leaf_data_offset: np.ndarray
    leaf_axis_start: np.ndarray
    leaf_axis_edge: np.ndarray
    leaf_data: np.ndarray  # float64, every leaf C-ordered
    n_slice_groups: int
    leaf_tensor_index: List[int]

    @property
    def n_nodes(self) -> int:
        return int(self.node_left.shape[0])

    @property
    def n_leaves(self) -> int:
        return int(self.leaf_rank.shape[0])


def _host_factory(shape, default_value=None):
    if default_value is None:
        return np.empty(shape, dtype=np.float64)
    return np.full(shape, default_value, dtype=np.float64)


class _LeafArena:
    """`tensor_factory` that hands out views of one growing float64 buffer, so the reference's
    `Tensor.build(factory)` writes each leaf straight into the buffer passed to `tob_plan_upload`
    (no per-leaf allocation, no concatenate)."""

    def __init__(self, capacity=4096):
        self.buf = np.empty(capacity, dtype=np.float64)
        self.used = 0
        self._last = None

    def factory(self, shape, default_value=None):
        n = 1
        for s in shape:
            n *= int(s)
        if self.used + n > self.buf.shape[0]:
            grown = np.empty(max(2 * self.buf.shape[0], self.used + n), dtype=np.float64)
            grown[: self.used] = self.buf[: self.used]
            self.buf = grown
        view = self.buf[self.used: self.used + n].reshape(shape)
        if default_value is not None:
            view[...] = default_value
        self._last = view
        return view

    def commit(self, built, n):
        """Keep `built` (normally the view just handed out) as the next n doubles of the buffer."""
        offset = self.used
        if built is self._last:  # the usual case: build() filled and returned the view it was handed
            self.used = offset + n
            return offset
        target = self.buf[offset: offset + n]
        if not (isinstance(built, np.ndarray) and built.dtype == np.float64 and np.shares_memory(built, target)
                and built.flags["C_CONTIGUOUS"]):
            if offset + n > self.buf.shape[0]:
                self.factory((n,))  # grow
                target = self.buf[offset: offset + n]
            target[:] = np.asarray(built, dtype=np.float64).reshape(-1)
        self.used = offset + n
        return offset


try:  # the compiled loop (tensororder_b200/_flatten_fast.pyx, built by tensororder_b200.build); same logic as below
    from . import _flatten_fast
except ImportError:  # not built: the Python loop below does the same work
    _flatten_fast = None

USE_COMPILED = True  # tests flip this to run both implementations


def rebuild_leaf_data(plan, flat: FlatPlan):
    """Leaf values of `plan` re-read through `Tensor.build()` in the order `flatten_plan` stored them (plan
    cache hits: the tree walk and the compile are skipped, the caller's tensors are still read on every call).
    Returns None when the leaves no longer fit the cached structure (the caller then re-flattens)."""
    network = plan.network
    if _flatten_fast is not None and USE_COMPILED:
        try:
            return _flatten_fast.rebuild_leaves(network, flat.leaf_tensor_index, flat.leaf_rank, flat.leaf_data_offset,
                                                int(flat.leaf_data.shape[0]))
        except (IndexError, KeyError, AttributeError, TypeError):
            return None
    arena = _LeafArena(max(int(flat.leaf_data.shape[0]), 16))
    seen = set()
    try:
        for j, t in enumerate(flat.leaf_tensor_index):
            if t in seen:
                continue
            seen.add(t)
            size = 1 << int(flat.leaf_rank[j])
            built = network[t].build(arena.factory)
            if getattr(built, "size", size) != size or arena.used != int(flat.leaf_data_offset[j]):
                return None
            arena.commit(built, size)
    except (IndexError, KeyError, AttributeError, TypeError):
        return None
    if arena.used != int(flat.leaf_data.shape[0]):
        return None
    return arena.buf[: arena.used]




def flatten_plan(plan, tensor_factory=None) -> FlatPlan:
    network = plan.network
    # edge id -> slice group index (non-empty groups only, in order)
    group_of = {}
    n_groups = 0
    for group in plan.groups_to_slice:
        if len(group) == 0:
            continue
        for e in group:
            group_of[int(e)] = n_groups
        n_groups += 1
    if _flatten_fast is not None and USE_COMPILED and tensor_factory is None:
        (nl, nr, nf, lr, lo, ast, ae, data, lti) = _flatten_fast.flatten_tree(plan, group_of)
        return FlatPlan(node_left=nl, node_right=nr, node_leaf=nf, leaf_rank=lr, leaf_data_offset=lo,
                        leaf_axis_start=ast, leaf_axis_edge=ae, leaf_data=data, n_slice_groups=n_groups,
                        leaf_tensor_index=lti)

    node_left: List[int] = []
    node_right: List[int] = []
    node_leaf: List[int] = []
    leaf_rank: List[int] = []
    leaf_off: List[int] = []
    axis_start: List[int] = [0]
    axis_edge: List[int] = []
    leaf_tensor_index: List[int] = []
    arena = _LeafArena()
    chunks: List[np.ndarray] = []  # only with a caller-supplied factory
    built_at = {}  # tensor index -> offset (a tensor named by two leaves is stored once)
    total = 0
    stack: List[int] = []
    # hot loop (one iteration per tree node): bind everything to locals
    index_list = network.index_list
    factory, commit = arena.factory, arena.commit
    nl_append, nr_append, nf_append = node_left.append, node_right.append, node_leaf.append
    push, pop = stack.append, stack.pop
    pos = 0
    n_leaves = 0
    for node in plan.tree.iterate_postorder():
        if node.is_leaf:
            t = node.tensor_index
            edges = index_list(t)
            rank = len(edges)
            if rank and min(edges) < 0:
                raise ValueError("tensor %d has a dangling index; the network cannot contract to a scalar" % t)
            off = built_at.get(t)
            if off is None:
                size = 1 << rank
                if tensor_factory is None:
                    built = network[t].build(factory)
                    if getattr(built, "size", size) != size:
                        raise ValueError("tensor %d: only indices of extent 2 are supported" % t)
                    off = commit(built, size)
                else:
                    data = np.ascontiguousarray(network[t].build(tensor_factory), dtype=np.float64)
                    if data.size != size:
                        raise ValueError("tensor %d: only indices of extent 2 are supported" % t)
                    off = total
                    chunks.append(data.reshape(-1))
                    total += data.size
                built_at[t] = off
            nl_append(-1)
            nr_append(-1)
            nf_append(n_leaves)
            n_leaves += 1
            leaf_rank.append(rank)
            leaf_off.append(off)
            if group_of:
                axis_edge.extend([-(group_of[e] + 1) if e in group_of else e for e in edges])
            else:
                axis_edge.extend(edges)
            axis_start.append(len(axis_edge))
            leaf_tensor_index.append(t)
        else:
            right = pop()
            nl_append(pop())
            nr_append(right)
            nf_append(-1)
        push(pos)
        pos += 1
    if len(stack) != 1:
        raise ValueError("contraction tree is not a single rooted tree")
    if tensor_factory is None:
        leaf_data = arena.buf[: arena.used]
    else:
        leaf_data = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.float64)
    return FlatPlan(
        node_left=np.asarray(node_left, dtype=np.int32),
        node_right=np.asarray(node_right, dtype=np.int32),
        node_leaf=np.asarray(node_leaf, dtype=np.int32),
        leaf_rank=np.asarray(leaf_rank, dtype=np.int32),
        leaf_data_offset=np.asarray(leaf_off, dtype=np.int64),
        leaf_axis_start=np.asarray(axis_start, dtype=np.int32),
        leaf_axis_edge=np.asarray(axis_edge, dtype=np.int32),
        leaf_data=np.ascontiguousarray(leaf_data, dtype=np.float64),
        n_slice_groups=n_groups,
        leaf_tensor_index=leaf_tensor_index,
    )
