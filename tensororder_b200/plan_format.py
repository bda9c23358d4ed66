"""Portable plan wire format ("toplan") for TensorOrder sliced execution plans.

The reference stores a plan as a pickle of Cython classes (`planning.py:108-111`,
reloaded at `execution.py:95`), which can only be read with the reference built
and importable.  This module defines a self-contained JSON document holding the
same information the executor consumes:

* the tensor network: per tensor its shape, dense float64 data (what
  `Tensor.build(factory)` returns, `tensor_network_constructions.py:69-99,144-152`)
  and its index list (edge id per axis, `tensor_network.pyx:38-39`);
* the contraction tree in post-order (`contraction_tree.pyx:174-183`);
* the slice groups (`sliced_execution_plan.py:58-82`).

`export_plan` reads a live reference plan by duck typing (nothing is imported from
the reference), `PortablePlan.as_execution_plan()` gives back objects with the same
attribute surface a tensor-library backend touches (`plan.tree`, `plan.network`,
`plan.groups_to_slice`; `base_api.py:17-28`), so the same backend code path handles
both a reference plan and a stored one.
"""
from __future__ import annotations

import gzip
import json
import os
from dataclasses import dataclass, field
from typing import Any, Dict, Iterator, List, Optional, Sequence

import numpy as np

FORMAT = "tensororder-b200-plan/1"


# --------------------------------------------------------------------------------------
# Duck-typed views (mirror the attribute surface of the reference objects)
# --------------------------------------------------------------------------------------
class PlanTensor:
    """Stands in for `tensor_network.tensor.Tensor` (`tensor.py:4-35`)."""

    def __init__(self, shape: Sequence[int], data: np.ndarray, diagonal: bool, label: str):
        self.shape = tuple(int(s) for s in shape)
        self._data = np.asarray(data, dtype=np.float64).reshape(self.shape)
        self._diagonal = bool(diagonal)
        self.label = label

    @property
    def rank(self) -> int:
        return len(self.shape)

    @property
    def diagonal(self) -> bool:
        return self._diagonal

    def build(self, tensor_factory):
        # Same contract as BuiltTensor.build (`tensor.py:43-49`)
        result = tensor_factory(self.shape)
        if len(self.shape) == 0:
            result[tuple()] = self._data[tuple()]
        else:
            result[:] = self._data[:]
        return result


class PlanNetwork:
    """Stands in for `TensorNetwork` (`tensor_network.pyx:11-49`), read-only."""

    def __init__(self, tensors: List[PlanTensor], index_lists: List[List[int]], edges: List[List[int]]):
        self._tensors = tensors
        self._index_lists = index_lists
        self._edges = edges

    @property
    def tensors(self) -> Iterator[PlanTensor]:
        return iter(self._tensors)

    def tensor(self, tensor_id: int) -> PlanTensor:
        return self._tensors[tensor_id]

    def __getitem__(self, tensor_id: int) -> PlanTensor:
        return self._tensors[tensor_id]

    def __len__(self) -> int:
        return len(self._tensors)

    def index_list(self, tensor_id: int) -> List[int]:
        return self._index_lists[tensor_id]

    @property
    def edges(self):
        return iter(self.edge(i) for i in range(len(self._edges)))

    def edge(self, edge_index: int) -> Dict[str, int]:
        t1, t2 = self._edges[edge_index]
        return {"id": edge_index, "tensor1_id": t1, "tensor2_id": t2}

    def num_edges(self) -> int:
        return len(self._edges)


class PlanTreeNode:
    """One node of a stored tree; the subset of `ContractionTree` (`contraction_tree.pyx:4-47`)
    an executor needs: `is_leaf`, `tensor_index`, `left`, `right`, `iterate_postorder`."""

    __slots__ = ("_tree", "_pos")

    def __init__(self, tree: "PlanTree", pos: int):
        self._tree = tree
        self._pos = pos

    @property
    def is_leaf(self) -> bool:
        return self._tree.postorder[self._pos][0] == "leaf"

    @property
    def tensor_index(self) -> int:
        return self._tree.postorder[self._pos][1]

    @property
    def left(self) -> "PlanTreeNode":
        return PlanTreeNode(self._tree, self._tree.postorder[self._pos][1])

    @property
    def right(self) -> "PlanTreeNode":
        return PlanTreeNode(self._tree, self._tree.postorder[self._pos][2])

    def iterate_postorder(self) -> Iterator["PlanTreeNode"]:
        return self._tree.iterate_postorder()


class PlanTree:
    """Post-order list of ("leaf", tensor_index) / ("join", left_pos, right_pos)."""

    def __init__(self, postorder: List[tuple]):
        self.postorder = postorder

    def iterate_postorder(self) -> Iterator[PlanTreeNode]:
        for pos in range(len(self.postorder)):
            yield PlanTreeNode(self, pos)

    @property
    def is_leaf(self) -> bool:
        return self.postorder[-1][0] == "leaf"


class StoredExecutionPlan:
    """Attribute surface of `SlicedExecutionPlan` used at the backend boundary
    (`sliced_execution_plan.py:9-27`, consumed at `base_api.py:17-28`)."""

    def __init__(self, tree: PlanTree, network: PlanNetwork, groups_to_slice: List[set]):
        self.tree = tree
        self.network = network
        self.groups_to_slice = groups_to_slice
        self.edges_to_slice = set().union(*groups_to_slice) if groups_to_slice else set()


# --------------------------------------------------------------------------------------
# The document
# --------------------------------------------------------------------------------------
@dataclass
class PortablePlan:
    name: str
    tensors: List[Dict[str, Any]]  # {"shape": [...], "data": [...], "diagonal": bool, "kind": str}
    index_lists: List[List[int]]
    edges: List[List[int]]
    postorder: List[List[int]]  # [tensor_index] for a leaf, [left_pos, right_pos] for a join
    groups_to_slice: List[List[int]]
    expected: Dict[str, Any] = field(default_factory=dict)
    meta: Dict[str, Any] = field(default_factory=dict)
    tree_check: Optional[Dict[str, Any]] = None
    # other slicings of the same tree+network: [{"name", "groups_to_slice", "expected"}]
    variants: List[Dict[str, Any]] = field(default_factory=list)

    # ---- (de)serialisation ----
    def to_json(self) -> Dict[str, Any]:
        doc = {
            "format": FORMAT,
            "name": self.name,
            "meta": self.meta,
            "tensors": self.tensors,
            "index_lists": self.index_lists,
            "edges": self.edges,
            "postorder": self.postorder,
            "groups_to_slice": self.groups_to_slice,
            "expected": self.expected,
        }
        if self.tree_check is not None:
            doc["tree_check"] = self.tree_check
        if self.variants:
            doc["variants"] = self.variants
        return doc

    @staticmethod
    def from_json(doc: Dict[str, Any]) -> "PortablePlan":
        if doc.get("format") != FORMAT:
            raise ValueError("not a %s document (format=%r)" % (FORMAT, doc.get("format")))
        return PortablePlan(
            name=doc["name"],
            tensors=doc["tensors"],
            index_lists=doc["index_lists"],
            edges=doc["edges"],
            postorder=doc["postorder"],
            groups_to_slice=doc["groups_to_slice"],
            expected=doc.get("expected", {}),
            meta=doc.get("meta", {}),
            tree_check=doc.get("tree_check"),
            variants=doc.get("variants", []),
        )

    def save(self, path: str) -> None:
        text = json.dumps(self.to_json(), separators=(",", ":"))
        opener = gzip.open if path.endswith(".gz") else open
        tmp = path + ".tmp"
        with opener(tmp, "wt") as f:
            f.write(text)
        os.replace(tmp, path)

    @staticmethod
    def load(path: str) -> "PortablePlan":
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rt") as f:
            return PortablePlan.from_json(json.load(f))

    # ---- views ----
    def as_execution_plan(self) -> StoredExecutionPlan:
        tensors = [
            PlanTensor(t["shape"], np.array(t["data"], dtype=np.float64), t.get("diagonal", False), t.get("kind", ""))
            for t in self.tensors
        ]
        network = PlanNetwork(tensors, [list(x) for x in self.index_lists], [list(e) for e in self.edges])
        postorder = [("leaf", n[0]) if len(n) == 1 else ("join", n[0], n[1]) for n in self.postorder]
        return StoredExecutionPlan(PlanTree(postorder), network, [set(g) for g in self.groups_to_slice])

    def variant(self, which) -> "PortablePlan":
        """The stored slicing variant `which` (index or name) as a plan of its own."""
        if isinstance(which, str):
            matches = [v for v in self.variants if v["name"] == which]
            if not matches:
                raise KeyError(which)
            v = matches[0]
        else:
            v = self.variants[which]
        return self.with_slices(v["groups_to_slice"], v.get("expected"), name=self.name + "/" + v["name"])

    def with_slices(self, groups_to_slice: List[Sequence[int]], expected: Optional[Dict[str, Any]] = None,
                    name: Optional[str] = None) -> "PortablePlan":
        return PortablePlan(
            name=name or self.name,
            tensors=self.tensors,
            index_lists=self.index_lists,
            edges=self.edges,
            postorder=self.postorder,
            groups_to_slice=[sorted(int(e) for e in g) for g in groups_to_slice],
            expected=dict(expected or {}),
            meta=dict(self.meta),
            tree_check=self.tree_check,
        )


def _host_factory(shape, default_value=None):
    if default_value is None:
        return np.empty(shape, dtype=np.float64)
    return np.full(shape, default_value, dtype=np.float64)


def export_plan(plan, name: str, meta: Optional[Dict[str, Any]] = None, with_tree_check: bool = False) -> PortablePlan:
    """Serialise a live plan (reference `SlicedExecutionPlan` or anything with the same
    attributes).  Reads only: `plan.tree.iterate_postorder()` + node `.is_leaf/.tensor_index`
    (`contraction_tree.pyx:174-183`), `plan.network` `.tensors/.index_list/.edges`
    (`tensor_network.pyx:23-45`), `plan.groups_to_slice`."""
    network = plan.network
    tensors = []
    for t in network.tensors:
        built = np.asarray(t.build(_host_factory), dtype=np.float64)
        tensors.append(
            {
                "shape": [int(s) for s in t.shape],
                "data": [float(x) for x in built.reshape(-1)],
                "diagonal": bool(t.diagonal),
                "kind": type(t).__name__,
            }
        )
    index_lists = [[int(e) for e in network.index_list(i)] for i in range(len(network))]
    edges = [[int(e["tensor1_id"]), int(e["tensor2_id"])] for e in network.edges]

    postorder: List[List[int]] = []
    stack: List[int] = []
    check = {"free_edges": [], "left_edge_map": [], "right_edge_map": []} if with_tree_check else None
    for node in plan.tree.iterate_postorder():
        pos = len(postorder)
        if node.is_leaf:
            postorder.append([int(node.tensor_index)])
        else:
            right = stack.pop()
            left = stack.pop()
            postorder.append([left, right])
        stack.append(pos)
        if check is not None:
            check["free_edges"].append([int(e) for e in node.free_edges])
            if node.is_leaf:
                check["left_edge_map"].append(None)
                check["right_edge_map"].append(None)
            else:
                check["left_edge_map"].append([int(i) for i in node.left_edge_map])
                check["right_edge_map"].append([int(i) for i in node.right_edge_map])
    if len(stack) != 1:
        raise ValueError("contraction tree is not a single rooted tree")

    return PortablePlan(
        name=name,
        tensors=tensors,
        index_lists=index_lists,
        edges=edges,
        postorder=postorder,
        groups_to_slice=[sorted(int(e) for e in g) for g in plan.groups_to_slice],
        expected={},
        meta=dict(meta or {}),
        tree_check=check,
    )
