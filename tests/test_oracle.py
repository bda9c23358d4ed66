"""The oracle (oracle/numpy_oracle.py) against the golden vectors produced by the reference
itself (tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import numpy_oracle as O

ALL = golden_names()
# keep the CPU suite to a few minutes: contract with the oracle only where the reference took < 1.5 s
def _cheap(pp_expected):
    return pp_expected.get("numpy_seconds_buildbox_8c", 1e9) < 1.5


def test_readme_instance_count(golden):
    pp = golden("vc50_lineflow")
    assert pp.expected["count"] == 2802717837.0  # README.md:21 command, SURVEY.md §4
    assert O.contract_sliced(pp.to_json()) == 2802717837.0


@pytest.mark.parametrize("name", ALL)
def test_tree_properties_match_reference(name, golden):
    pp = golden(name)
    if pp.tree_check is None:
        pytest.skip("no tree_check stored")
    props = O.tree_properties(pp.to_json())
    for i, node in enumerate(props):
        assert node["free_edges"] == pp.tree_check["free_edges"][i]
        assert node["left_edge_map"] == pp.tree_check["left_edge_map"][i]
        assert node["right_edge_map"] == pp.tree_check["right_edge_map"][i]


@pytest.mark.parametrize("name", ALL)
def test_leaf_builders_match_reference(name, golden):
    pp = golden(name)
    for t in pp.tensors:
        params = t.get("params")
        if params is None:
            continue
        if t["kind"] == "OrTensor":
            built = O.build_or_tensor(params["literals_positive"], params["output_index"])
        else:
            built = O.build_variable_tensor(params["rank"], params["positive_weight"], params["negative_weight"])
        assert list(built.shape) == t["shape"]
        assert built.reshape(-1).tolist() == t["data"]


@pytest.mark.parametrize("name", ALL)
def test_unsliced_count_bit_exact(name, golden):
    pp = golden(name)
    if "count" not in pp.expected or not _cheap(pp.expected):
        pytest.skip("reference count not stored / too slow for the CPU suite")
    got = O.contract_sliced(pp.to_json())
    want = pp.expected["count"]
    if pp.expected["maxrank"] <= 16:
        # small GEMMs: OpenBLAS summation order is the same on any machine -> bit exact
        assert float(got).hex() == pp.expected["count_hex"]
    else:
        assert math.isclose(got, want, rel_tol=1e-12)


def _variants():
    out = []
    for name in ALL:
        pp = load_golden(name)
        for i, v in enumerate(pp.variants):
            if "count" in v.get("expected", {}) and _cheap(v["expected"]):
                out.append((name, i))
    return out


@pytest.mark.parametrize("name,vi", _variants())
def test_sliced_counts_and_partial_sums(name, vi, golden):
    pp = golden(name).variant(vi)
    doc = pp.to_json()
    per = []
    got = O.contract_sliced(doc, per_slice=per)
    exp = pp.expected
    assert len(per) == exp["num_slices"]
    assert math.isclose(got, exp["count"], rel_tol=1e-12)
    if "per_slice" in exp:
        assert np.allclose(per, exp["per_slice"], rtol=1e-12, atol=0)
    if "slice_cutoff" in exp:
        got_cut = O.contract_sliced(doc, num_slice_limit=exp["slice_cutoff"])
        assert math.isclose(got_cut, exp["count_cutoff"], rel_tol=1e-12)
        assert math.isclose(got_cut, sum(per[: exp["slice_cutoff"]]), rel_tol=1e-12)


@pytest.mark.parametrize("name", [n for n in ALL if n.startswith("toy")])
def test_einsum_agrees(name, golden):
    pp = golden(name)
    if "einsum" not in pp.expected:
        pytest.skip("more than 26 indices")
    assert math.isclose(float(O.contract_einsum(pp.to_json())), pp.expected["einsum"], rel_tol=1e-12)
    if name == "toy_unit_neg_lineflow":
        # Reference quirk kept as an input fact: the line-graph planner matches the rank-0 tensor of the
        # never-used variable 5 against a bag (line_graph_method.py:17-21, the empty set is a subset of
        # every bag) AND include_rank_zero_tensors adds it again (contraction_tree.pyx:302-310), so the
        # reference's own tree holds that leaf twice and its count is 2x the einsum value.
        assert sum(1 for n in pp.postorder if n == [4]) == 2
        assert pp.expected["count"] == 2 * pp.expected["einsum"]
    else:
        assert math.isclose(pp.expected["einsum"], pp.expected["count"], rel_tol=1e-12)


def test_unweighted_counts_invariant_under_slicing(golden):
    pp = golden("vc50_lineflow")
    for v in pp.variants:
        assert v["expected"]["count"] == 2802717837.0
