"""Edge cases of the plan boundary: degenerate trees, empty slice groups, zero slices, scalar operands,
more slices than the device result buffer holds.  CPU versions run the compiled program through the numpy
interpreter; the GPU versions run the real kernels."""
import math

import numpy as np
import pytest

from conftest import load_golden
from program_sim import run_program
from tensororder_b200.api import CompiledPlan
from tensororder_b200.flatten import flatten_plan
from tensororder_b200.plan_format import PortablePlan


def _plan(tensors, index_lists, edges, postorder, groups=()):
    return PortablePlan(name="edge", tensors=[{"shape": list(np.shape(t)), "data": [float(x) for x in np.ravel(t)],
                                               "diagonal": False, "kind": "BuiltTensor"} for t in tensors],
                        index_lists=index_lists, edges=edges, postorder=postorder,
                        groups_to_slice=[sorted(g) for g in groups])


def _variable_groups(pp, n_groups):
    """Valid slicings made by hand: one group = every edge of one variable tensor (all equivalent through the
    diagonal tensor, src/tensor_network/tensor_network.pyx:470-489)."""
    groups = []
    for t, doc in enumerate(pp.tensors):
        if doc.get("kind") == "VariableTensor" and len(pp.index_lists[t]) > 0:
            groups.append(list(pp.index_lists[t]))
            if len(groups) == n_groups:
                break
    return groups


CASES = {
    # a network that is one rank-0 tensor: the tree is a single leaf
    "single_scalar": (_plan([np.float64(3.5)], [[]], [], [[0]]), 3.5),
    # two scalars joined (m = n = k = 0)
    "two_scalars": (_plan([np.float64(3.0), np.float64(0.5)], [[], []], [], [[0], [1], [0, 1]]), 1.5),
    # scalar times a rank-2 tensor, then closed by a rank-2 tensor
    "scalar_operand": (_plan([np.float64(2.0), np.arange(4.0).reshape(2, 2), np.ones((2, 2))], [[], [0, 1], [0, 1]],
                             [[1, 2], [1, 2]], [[0], [1], [0, 1], [2], [2, 3]]), 12.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_degenerate_trees_compile_and_interpret(name):
    pp, want = CASES[name]
    flat = flatten_plan(pp.as_execution_plan())
    cp = CompiledPlan(flat)
    assert math.isclose(run_program(cp.describe(), flat), want, rel_tol=1e-15)
    cp.close()


def test_empty_slice_groups_are_skipped_like_the_reference():
    pp = load_golden("vc50_lineflow").variant("min3")
    with_empty = pp.with_slices([[]] + [list(g) for g in pp.groups_to_slice[:1]] + [[]] + [list(g) for g in pp.groups_to_slice[1:]])
    flat = flatten_plan(with_empty.as_execution_plan())
    assert flat.n_slice_groups == 3  # tensor_network.pyx:371-373 `if len(group) == 0: continue`
    cp = CompiledPlan(flat)
    assert cp.num_slices == 8
    assert run_program(cp.describe(), flat) == 2802717837.0
    cp.close()


def test_hand_made_variable_slicing_is_count_invariant():
    pp = load_golden("vc50_lineflow")
    sliced = pp.with_slices(_variable_groups(pp, 5))
    flat = flatten_plan(sliced.as_execution_plan())
    cp = CompiledPlan(flat)
    assert cp.num_slices == 32
    assert run_program(cp.describe(), flat) == 2802717837.0
    cp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_degenerate_trees_on_the_gpu(name):
    from tensororder_b200.api import B200API

    pp, want = CASES[name]
    api = B200API()
    api.add_argument("entry_type", "float64")
    assert float(api.contract_sliced(pp.as_execution_plan())) == want


@pytest.mark.gpu
def test_zero_slices_and_more_slices_than_the_result_buffer():
    from tensororder_b200.api import B200API

    pp = load_golden("vc50_lineflow")
    api = B200API()
    api.add_argument("entry_type", "float64")
    assert float(api.contract_sliced(pp.variant("min3").as_execution_plan(), num_slice_limit=0)) == 0.0
    many = pp.with_slices(_variable_groups(pp, 13))  # 8192 slices > 4096 buffered per batch
    got = api.contract_sliced(many.as_execution_plan())
    assert float(got) == 2802717837.0  # integer partial sums: exact in any grouping
    assert api.last_stats["slices"] == 8192
