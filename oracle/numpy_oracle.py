"""CPU oracle for the TensorOrder contraction-executor hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, what the reference's numpy backend computes for a sliced
execution plan.  It is imported only by `tests/`, `__graft_entry__.smoke()` and
`bench.py` (cpu_baseline / `--impl reference` legs) as the *checker* and the *timed CPU
baseline*; nothing under `tensororder_b200/` may import it and the product path never
falls back to it.

Parity status: PINNED.  The reference has no tests of its own for this path (SURVEY.md §4,
"parity unpinned" upstream), so the pins were created by running the reference itself
(numpy backend, built in a scratch copy) in the build container; the outputs are committed
as `tests/golden/*.json.gz` together with the generating script
`tests/golden/make_golden.py`, and `tests/test_oracle.py` checks this oracle against every
one of them (bit-exact), plus the README instance count 2802717837.

Where the arithmetic really lives: third-party numpy (`numpy.tensordot` = transpose-copy x2
+ OpenBLAS dgemm), pinned `numpy==1.18.1` in the reference's requirements.txt:36, numpy 2.3.5
here; called from `src/tensor_network/tensor_apis/numpy_apis.py:42-43`.  The oracle calls the
same `numpy.tensordot` with the same operands, axes and operand order, so it reproduces the
reference bit for bit on the same machine.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import itertools
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# Leaf tensors
# --------------------------------------------------------------------------------------
def build_or_tensor(literals_positive: Sequence[bool], output_index: Optional[int] = None) -> np.ndarray:
    """`OrTensor.build`, src/tensor_network/tensor_network_constructions.py:69-99."""
    rank = len(literals_positive)
    result = np.full((2,) * rank, 1, dtype=np.float64)
    if output_index is None:
        result[tuple(0 if p else 1 for p in literals_positive)] = 0
    else:
        out_false = 0 if literals_positive[output_index] else 1
        result[tuple(out_false if i == output_index else slice(0, 2) for i in range(rank))] = 0
        all_false = [0 if p else 1 for p in literals_positive]
        result[tuple(all_false)] = 1
        all_false[output_index] = 1 - all_false[output_index]
        result[tuple(all_false)] = 0
    return result


def build_variable_tensor(rank: int, positive_weight: float, negative_weight: float) -> np.ndarray:
    """`VariableTensor.build`, src/tensor_network/tensor_network_constructions.py:144-152."""
    result = np.full((2,) * rank, 0, dtype=np.float64)
    if rank == 0:
        result[()] = negative_weight + positive_weight
    else:
        result[(0,) * rank] = negative_weight
        result[(1,) * rank] = positive_weight
    return result


def build_leaf(tensor_doc: Dict[str, Any]) -> np.ndarray:
    """`BuiltTensor.build` (src/tensor_network/tensor.py:43-49): a fresh copy of the stored data."""
    shape = tuple(tensor_doc["shape"])
    return np.array(tensor_doc["data"], dtype=np.float64).reshape(shape).copy()


# --------------------------------------------------------------------------------------
# Contraction tree: join properties
# --------------------------------------------------------------------------------------
def compute_join_properties(left_free: Sequence[int], right_free: Sequence[int]) -> Tuple[List[int], List[int], List[int]]:
    """`ContractionTreeContext.compute_join_properties`,
    src/contraction_methods/contraction_tree.pyx:248-288.

    Returns (left_edge_map, right_edge_map, free_edges): paired axis positions (ascending in
    the left operand), and the surviving edges = left-free in order then right-free in order,
    which is exactly the axis order `numpy.tensordot` produces."""
    left_edge_map: List[int] = []
    right_edge_map: List[int] = []
    free: List[int] = []
    for i, le in enumerate(left_free):
        found = False
        for j, re in enumerate(right_free):
            if le == re:
                left_edge_map.append(i)
                right_edge_map.append(j)
                found = True
                break
        if not found:
            free.append(le)
    for re in right_free:
        if re not in left_free:
            free.append(re)
    return left_edge_map, right_edge_map, free


def tree_properties(doc: Dict[str, Any]) -> List[Dict[str, Any]]:
    """Per post-order node: free_edges / edge maps.  Leaves take the tensor's index list
    (`ContractionTreeContext.leaf`, contraction_tree.pyx:224-235)."""
    props: List[Dict[str, Any]] = []
    for node in doc["postorder"]:
        if len(node) == 1:
            props.append({"is_leaf": True, "tensor_index": node[0], "free_edges": list(doc["index_lists"][node[0]]),
                          "left_edge_map": None, "right_edge_map": None})
        else:
            lm, rm, free = compute_join_properties(props[node[0]]["free_edges"], props[node[1]]["free_edges"])
            props.append({"is_leaf": False, "left": node[0], "right": node[1], "free_edges": free,
                          "left_edge_map": lm, "right_edge_map": rm})
    return props


# --------------------------------------------------------------------------------------
# Slicing (numpy-backend scheme: sliced axes stay as size-1 axes)
# --------------------------------------------------------------------------------------
def slice_lookups(doc: Dict[str, Any]) -> Tuple[List[List[Tuple[int, int]]], List[List[int]]]:
    """First half of `TensorNetwork.slice_groups`, src/tensor_network/tensor_network.pyx:363-394:
    for every non-empty group the (tensor, axis) pairs it touches and the value range."""
    tensor_infos: List[List[Tuple[int, int]]] = []
    index_values: List[List[int]] = []
    for group in doc["groups_to_slice"]:
        if len(group) == 0:
            continue
        infos: List[Tuple[int, int]] = []
        # the reference iterates a Python set of ints; order inside a group does not change the result
        for e in group:
            t1, t2 = doc["edges"][e]
            infos.append((t1, doc["index_lists"][t1].index(e)))
            infos.append((t2, doc["index_lists"][t2].index(e)))
        tensor_infos.append(infos)
        t, axis = infos[-1]
        index_values.append(list(range(doc["tensors"][t]["shape"][axis])))
    return tensor_infos, index_values


def iter_slice_assignments(doc: Dict[str, Any]) -> Iterable[Tuple[int, ...]]:
    """`itertools.product` order of tensor_network.pyx:396 (first group most significant)."""
    _, index_values = slice_lookups(doc)
    return itertools.product(*index_values)


def sliced_leaves(doc: Dict[str, Any], assignment: Sequence[int]) -> List[np.ndarray]:
    """Leaves of one slice network: `Tensor.get_slice` / `SlicedTensor.build`
    (src/tensor_network/tensor.py:20-25,66-73, applied at tensor_network.pyx:396-401).
    Every leaf is rebuilt, then indexed so that sliced axes keep extent 1."""
    tensor_infos, _ = slice_lookups(doc)
    lookups: Dict[int, List[Any]] = {}
    for infos, value in zip(tensor_infos, assignment):
        for t, axis in infos:
            lk = lookups.setdefault(t, [slice(0, s) for s in doc["tensors"][t]["shape"]])
            lk[axis] = slice(value, value + 1)
    leaves = []
    for t, tensor_doc in enumerate(doc["tensors"]):
        built = build_leaf(tensor_doc)
        if t in lookups:
            built = built[tuple(lookups[t])]
        leaves.append(built)
    return leaves


# --------------------------------------------------------------------------------------
# The executor
# --------------------------------------------------------------------------------------
def identify(doc: Dict[str, Any], leaves: List[np.ndarray], props: Optional[List[Dict[str, Any]]] = None,
             record: Optional[List[Any]] = None) -> np.ndarray:
    """`TensorNetwork.identify`, src/tensor_network/tensor_network.pyx:142-156, with
    `NumpyAPI.tensordot` = `numpy.tensordot` (numpy_apis.py:42-43)."""
    if props is None:
        props = tree_properties(doc)
    stack: List[np.ndarray] = []
    for node in props:
        if node["is_leaf"]:
            stack.append(leaves[node["tensor_index"]])
        else:
            right = stack.pop()
            left = stack.pop()
            result = np.tensordot(left, right, (node["left_edge_map"], node["right_edge_map"]))
            if record is not None:
                record.append((left.shape, right.shape, node["left_edge_map"], node["right_edge_map"]))
            stack.append(result)
    return stack[0]


def contract_sliced(doc: Dict[str, Any], num_slice_limit: Optional[int] = None,
                    per_slice: Optional[List[float]] = None):
    """`BaseTensorAPI.contract_sliced`, src/tensor_network/tensor_apis/base_api.py:17-28:
    sequential float64 sum, in slice order, of the rank-0 result of every slice network."""
    props = tree_properties(doc)
    result = 0
    assignments: Iterable[Tuple[int, ...]] = iter_slice_assignments(doc)
    if num_slice_limit is not None:
        assignments = itertools.islice(assignments, num_slice_limit)
    for assignment in assignments:
        leaves = sliced_leaves(doc, assignment)
        tensor_result = identify(doc, leaves, props)
        value = tensor_result[tuple()]
        if per_slice is not None:
            per_slice.append(float(value))
        result += value
    return result


def contract_einsum(doc: Dict[str, Any]):
    """`TensorNetwork.contract_einsum`, src/tensor_network/tensor_network.pyx:126-140:
    tree-independent check for networks with at most 26 distinct indices."""
    index_lists = doc["index_lists"]
    dims = sorted(set().union(*map(set, index_lists))) if index_lists else []
    if len(dims) > 26:
        raise ValueError("Unable to use einsum: more than 26 dimensions")
    labels = {d: chr(ord("a") + i) for i, d in enumerate(dims)}
    operands = ",".join("".join(labels[i] for i in il) for il in index_lists)
    return np.einsum(operands, *[build_leaf(t) for t in doc["tensors"]])


def tensordot(a: np.ndarray, b: np.ndarray, axes) -> np.ndarray:
    """`NumpyAPI.tensordot`, numpy_apis.py:42-43."""
    return np.tensordot(a, b, axes)


# --------------------------------------------------------------------------------------
# Work model: algorithmic bytes / flops per join node
# --------------------------------------------------------------------------------------
def node_work(doc: Dict[str, Any]) -> List[Dict[str, int]]:
    """(rA, rB, k, rC), flops = 2*2^(fL+fR+k), bytes = 8*(2^rA + 2^rB + 2^rC) per join,
    with sliced edges removed (they have extent 1)."""
    sliced = set()
    for g in doc["groups_to_slice"]:
        sliced |= set(g)
    props = tree_properties(doc)
    out = []
    for node in props:
        if node["is_leaf"]:
            continue
        la = [e for e in props[node["left"]]["free_edges"] if e not in sliced]
        rb = [e for e in props[node["right"]]["free_edges"] if e not in sliced]
        k = len(set(la) & set(rb))
        rc = len(la) + len(rb) - 2 * k
        out.append({"rA": len(la), "rB": len(rb), "k": k, "rC": rc,
                    "flops": 2 * 2 ** (rc + k), "bytes": 8 * (2 ** len(la) + 2 ** len(rb) + 2 ** rc)})
    return out
