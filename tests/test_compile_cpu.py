"""Host-side logic without a GPU: the C-ABI library loads and exports every declared symbol, the plan
flattener + C++ plan compiler produce programs whose canonical-form semantics reproduce the reference
counts (interpreted in numpy by tests/program_sim.py), error paths behave like the reference's."""
import math
import os
import re

import numpy as np
import pytest

from conftest import REPO, golden_names, load_golden
from program_sim import run_program
from tensororder_b200 import cabi
from tensororder_b200.api import B200API, CompiledPlan
from tensororder_b200.flatten import flatten_plan

ALL = golden_names()
SMALL = [n for n in ALL if load_golden(n).expected.get("maxrank", 99) <= 19]


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "tob200.h")).read()
    declared = set(re.findall(r"\b(tob_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(cabi.SYMBOLS), (declared ^ set(cabi.SYMBOLS))
    for name in declared:
        assert hasattr(cabi.lib, name), name
    assert b"sm_100a" in cabi.lib.tob_version()


def test_add_argument_contract():
    api = B200API()
    api.add_argument("entry_type", "float64")
    assert api.get_entry_size() == 8
    with pytest.raises(ValueError):
        api.add_argument("entry_type", "complex64")  # not in the reference's table either (numpy_apis.py:15-27)
    with pytest.raises(ValueError, match="Invalid argument"):
        api.add_argument("TPU", "1.2.3.4")  # base_api.py:9-12
    t = api.create_tensor((2, 2), 1)
    assert t.dtype == np.float64 and t.sum() == 4


@pytest.mark.parametrize("name", SMALL)
def test_compiled_program_reproduces_reference_count(name):
    pp = load_golden(name)
    flat = flatten_plan(pp.as_execution_plan())
    cp = CompiledPlan(flat)
    desc = cp.describe()
    got = run_program(desc, flat)
    assert math.isclose(got, pp.expected["count"], rel_tol=1e-12), (got, pp.expected["count"])
    assert cp.peak_bytes >= 8 * (desc["leaf_doubles"] + desc["arena_doubles"])
    cp.close()


def _small_variants():
    out = []
    for name in SMALL:
        pp = load_golden(name)
        for i, v in enumerate(pp.variants):
            if "count" in v.get("expected", {}) and v["expected"]["num_slices"] <= 32:
                out.append((name, i))
    return out


@pytest.mark.parametrize("name,vi", _small_variants())
@pytest.mark.parametrize("hoist", [True, False])
def test_compiled_sliced_program(name, vi, hoist):
    pp = load_golden(name).variant(vi)
    flat = flatten_plan(pp.as_execution_plan())
    cp = CompiledPlan(flat, hoist_invariant=hoist)
    desc = cp.describe()
    assert cp.num_slices == pp.expected["num_slices"]
    got = run_program(desc, flat)
    assert math.isclose(got, pp.expected["count"], rel_tol=1e-12)
    if "per_slice" in pp.expected:
        # the multi-GPU partition: rank r of 2 takes slices r, r+2, ...
        parts = [run_program(desc, flat, first=r, stride=2) for r in range(2)]
        assert math.isclose(sum(parts), pp.expected["count"], rel_tol=1e-12)
        assert math.isclose(parts[1], sum(pp.expected["per_slice"][1::2]), rel_tol=1e-12)
    if "slice_cutoff" in pp.expected:
        got_cut = run_program(desc, flat, first=0, count=pp.expected["slice_cutoff"])
        assert math.isclose(got_cut, pp.expected["count_cutoff"], rel_tol=1e-12)
    if hoist:
        assert all(op["invariant"] == 1 for op in desc["invariant_ops"])
        assert all(op["invariant"] == 0 for op in desc["slice_ops"] if op["kind"] != 2)
    cp.close()


def test_forced_generic_policy_and_gemm_selection():
    pp = load_golden("vc150_lineflow")
    flat = flatten_plan(pp.as_execution_plan())
    auto = CompiledPlan(flat).describe()
    kinds = [op["kind"] for op in auto["slice_ops"]]
    assert 1 in kinds, "the dominant joins of n=150 must go to the DMMA GEMM kernel"
    import ctypes

    from tensororder_b200 import cabi

    def tuned(key):
        v = ctypes.c_double()
        assert cabi.lib.tob_tuning_get(key.encode(), ctypes.byref(v)) == 0
        return int(v.value)

    for op in auto["slice_ops"]:
        if op["kind"] == 1:  # what the measured dispatch table (csrc/tob_dispatch_table.h) allows on the GEMM kernel
            min_free = tuned("gemm_min_free") if op["k"] >= tuned("gemm_min_k") else tuned("gemm_smallk_min_free")
            assert op["m"] >= op["n"] >= min_free and op["k"] >= 1
            assert op["m"] + op["n"] >= tuned("gemm_min_out.%d" % min(op["k"], 16))
            assert op["k"] - op["ksplit_log2"] >= tuned("min_k_per_split_log2") or op["ksplit_log2"] == 0
    forced = CompiledPlan(flat, kernel_policy=1).describe()
    assert all(op["kind"] in (0, 2, 3) for op in forced["slice_ops"])


def _streamk_partition(tiles_log2, k, ctas):
    """The range/segment arithmetic of k_gemm_dmma_sk (csrc/tob_kernels.cu), restated: returns for every CTA rank its
    segments in PROCESSING order, and for every tile whose finishing CTA did not start it the (owner, first contributor)."""
    kt = 1 << (k - 4)
    total = kt << tiles_log2
    q, r = divmod(total, ctas)
    segs, owners = {}, {}
    for rank in range(ctas):
        g0 = rank * q + min(rank, r)
        g1 = g0 + q + (1 if rank < r else 0)
        t_first, t_last = g0 >> (k - 4), (g1 - 1) >> (k - 4)
        nseg, end1 = t_last - t_first + 1, g1 & (kt - 1)
        rotate = end1 != 0 and nseg > 1
        mine = []
        for x in range(nseg):
            s_ = (nseg - 1 if x == 0 else x - 1) if rotate else x
            kt0 = (g0 & (kt - 1)) if s_ == 0 else 0
            kt1 = end1 if (s_ == nseg - 1 and end1 != 0) else kt
            mine.append((t_first + s_, kt0, kt1))
            if kt1 == kt and kt0 > 0:
                s0 = (t_first + s_) << (k - 4)
                owners[t_first + s_] = (rank, s0 // (q + 1) if s0 < r * (q + 1) else r + (s0 - r * (q + 1)) // q)
        segs[rank] = mine
    return kt, segs, owners


@pytest.mark.parametrize("tiles_log2,k", [(8, 10), (8, 9), (6, 12), (3, 10), (12, 9), (7, 8), (10, 8), (0, 13)])
def test_streamk_partition_covers_every_k_step_once(tiles_log2, k):
    """Stream-K: every (tile, K step) belongs to exactly one CTA; a CTA publishes at most one partial tile and does so in its
    FIRST segment (so the owner's wait never chains); an owner's contributors are exactly the lower ranks holding the
    tile's earlier K steps, in order — the deterministic summation order of the kernel."""
    ctas = 296
    kt, segs, owners = _streamk_partition(tiles_log2, k, ctas)
    if (kt << tiles_log2) < ctas:
        pytest.skip("fewer K steps than CTAs: choose_kernel never picks stream-K here")
    seen, partial = set(), {}
    for rank, mine in segs.items():
        for pos, (tile, kt0, kt1) in enumerate(mine):
            for step in range(kt0, kt1):
                assert (tile, step) not in seen
                seen.add((tile, step))
            if kt1 < kt:
                assert rank not in partial and pos == 0
                partial[rank] = (tile, kt0, kt1)
    assert len(seen) == kt << tiles_log2
    used = set()
    for tile, (owner, first) in owners.items():
        steps = []
        for j in range(first, owner):
            t, a, b = partial[j]
            assert t == tile and j < owner
            steps += list(range(a, b))
            used.add(j)
        own_start = [kt0 for (t, kt0, kt1) in segs[owner] if t == tile][0]
        assert steps == list(range(own_start))
    assert used == set(partial)


def test_streamk_is_chosen_for_badly_quantised_tile_counts_only():
    """Config 3's dominant joins (m=11, n=10, k=10: 256 tiles of 128x64 on 296 CTA slots) run on the stream-K kernel with one
    64 KB partial-tile slot per CTA in the workspace; the 2048-tile joins of the sliced family members do not."""
    import ctypes

    from tensororder_b200 import cabi

    flat = flatten_plan(load_golden("vc150_mcc_factorflow").as_execution_plan())
    desc = CompiledPlan(flat).describe()
    sk = [op for op in desc["slice_ops"] + desc["invariant_ops"] if op["kind"] == 1 and op["streamk"] > 0]
    assert sk and all(op["ksplit_log2"] == 0 and 6 <= (op["m"] - 7) + (op["n"] - 6) <= 8 and op["k"] >= 8 for op in sk)
    assert (11, 10, 10) in {(op["m"], op["n"], op["k"]) for op in sk}
    assert desc["ws_doubles"] >= max(op["streamk"] for op in sk) * 128 * 64
    big = CompiledPlan(flatten_plan(load_golden("vc210_lineflow").variant("min3").as_execution_plan())).describe()
    assert all(op["streamk"] == 0 for op in big["slice_ops"] if op["kind"] == 1 and op["m"] + op["n"] >= 24)
    assert cabi.lib.tob_tuning_set(b"streamk", 0.0) == 0
    try:
        off = CompiledPlan(flat).describe()
        assert all(op.get("streamk", 0) == 0 for op in off["slice_ops"] + off["invariant_ops"] if op["kind"] in (0, 1))
    finally:
        cabi.lib.tob_tuning_set(b"streamk", 1.0)


@pytest.mark.parametrize("name,variant", [("vc50_lineflow", None), ("vc100_lineflow", "min4"), ("vc150_lineflow", None)])
def test_micro_subtrees(name, variant):
    """Mini joins collapse into one launch per stage (level of a phase); results are unchanged with the
    feature off."""
    pp = load_golden(name)
    if variant:
        pp = pp.variant(variant)
    flat = flatten_plan(pp.as_execution_plan())
    on = CompiledPlan(flat).describe()
    off = CompiledPlan(flat, use_microtree=False).describe()
    n_joins = sum(1 for n in pp.postorder if len(n) == 2)
    assert sum(1 for op in off["slice_ops"] + off["invariant_ops"] if op["kind"] in (0, 1)) == n_joins
    micro = [op for op in on["invariant_ops"] + on["slice_ops"] if op["kind"] == 3]
    assert 1 <= len(micro) <= 8
    assert [op["which"] for op in micro] == list(range(len(micro)))  # stages in execution order
    in_micro = sum(len(op["micro"]) for op in micro)
    rest = sum(1 for op in on["slice_ops"] + on["invariant_ops"] if op["kind"] in (0, 1))
    assert in_micro + rest == n_joins and in_micro > n_joins // 3
    for op in micro:
        assert op["cta_start"] == sorted(op["cta_start"]) and len(op["cta_start"]) - 1 <= 2 * 148
    a, b = run_program(on, flat), run_program(off, flat)
    assert math.isclose(a, b, rel_tol=1e-12) and math.isclose(a, pp.expected["count"], rel_tol=1e-12)


def _dag_cases():
    return [("vc50_lineflow", None), ("vc100_lineflow", None), ("vc100_lineflow", "min4"), ("vc150_lineflow", None),
            ("vc150_lineflow", "min4"), ("vc150_mcc_factorflow", "min3"), ("rand3cnf_24_lineflow", None),
            ("rand3cnf_24_lineflow", "min3"), ("rand4cnf_18_mcc_lineflow", "min2"), ("vc100_mcc_factorflow", None)]


@pytest.mark.parametrize("name,variant", _dag_cases())
def test_dag_schedule_orders_every_hazard(name, variant):
    """Joins of independent subtrees run concurrently on several streams (Op::branch / waits).  The
    simulator executes the ops in random orders that respect only the declared schedule, against ONE
    shared arena: a missing RAW / WAR / WAW edge changes the result (or reads NaN-poisoned space)."""
    pp = load_golden(name)
    if variant:
        pp = pp.variant(variant)
    flat = flatten_plan(pp.as_execution_plan())
    want = pp.expected.get("count", load_golden(name).expected["count"])
    for branches in (0, 3, 32):
        for microtree in (True, False):
            cp = CompiledPlan(flat, dag_branches=branches, use_microtree=microtree)
            desc = cp.describe()
            ops = desc["invariant_ops"] + desc["slice_ops"]
            used = 1 + max(op.get("branch", 0) for op in ops)
            assert used == desc["branches"] <= (16 if branches == 0 else branches)
            seq = run_program(desc, flat)
            assert math.isclose(seq, want, rel_tol=1e-12)
            for seed in range(4):
                got = run_program(desc, flat, dag_seed=seed)
                assert got == seq or math.isclose(got, seq, rel_tol=1e-13), (branches, microtree, seed, got, seq)
            cp.close()
    one = CompiledPlan(flat, dag_branches=1).describe()
    assert one["branches"] == 1 and all(not op.get("waits") for op in one["invariant_ops"] + one["slice_ops"])
    # the point of the schedule: the longest chain is far shorter than the op count
    if name == "vc150_lineflow" and variant is None:
        desc = CompiledPlan(flat, use_microtree=False).describe()
        ops = desc["slice_ops"]
        depth = [0] * len(ops)
        last = {}
        for j, op in enumerate(ops):
            if op["kind"] in (2, 3):
                depth[j] = 1 + max(depth[:j], default=0)
                last = {}
                continue
            b = op["branch"]
            d = max([depth[w] for w in op["waits"]] + [depth[last[b]] if b in last else 0] +
                    [depth[i] for i in range(j) if ops[i]["kind"] in (2, 3)])
            depth[j] = d + 1
            last[b] = j
        assert max(depth) < len(ops) // 2, (max(depth), len(ops))


@pytest.mark.parametrize("shape", ["random", "caterpillar", "balanced"])
@pytest.mark.parametrize("seed", range(12))
def test_random_networks_and_trees(seed, shape):
    """Random closed networks, random trees of three shapes, random slicings, every combination of the
    compiler's features: the compiled program (simulated in list order and in random DAG orders) equals the
    plain numpy contraction of the network — exactly, the tensors hold small integers."""
    from random_plans import make

    rng = np.random.default_rng(1000 + seed)
    n_tensors = int(rng.integers(3, 40))
    n_edges = int(n_tensors * rng.uniform(1.0, 1.8))
    groups = int(rng.integers(0, 4))
    flat, want = make(seed, n_tensors=n_tensors, n_edges=n_edges, n_slice_groups=min(groups, n_edges), shape=shape)
    for hoist in (True, False):
        for microtree in (True, False):
            for branches in (0, 1, 4):
                cp = CompiledPlan(flat, hoist_invariant=hoist, use_microtree=microtree, dag_branches=branches)
                desc = cp.describe()
                assert run_program(desc, flat) == want, (hoist, microtree, branches)
                if branches != 1:
                    assert run_program(desc, flat, dag_seed=seed) == want, (hoist, microtree, branches)
                if cp.num_slices >= 2:
                    assert run_program(desc, flat, first=0, stride=2) + run_program(desc, flat, first=1, stride=2) == want
                cp.close()


@pytest.mark.parametrize("name", ALL)
def test_compiled_flatten_equals_python_flatten(name):
    """`flatten_plan` has its hot loop twice: tensororder_b200/_flatten_fast.pyx (Cython, built in-tree) and the
    Python loop it was written from.  Same arrays, bit for bit, on every fixture and variant."""
    from tensororder_b200 import flatten

    assert flatten._flatten_fast is not None, "python -m tensororder_b200.build builds _flatten_fast"
    pp = load_golden(name)
    for v in [None] + list(range(len(pp.variants))):
        plan = (pp if v is None else pp.variant(v)).as_execution_plan()
        try:
            flatten.USE_COMPILED = True
            a = flatten.flatten_plan(plan)
            flatten.USE_COMPILED = False
            b = flatten.flatten_plan(plan)
        finally:
            flatten.USE_COMPILED = True
        for fld in ("node_left", "node_right", "node_leaf", "leaf_rank", "leaf_data_offset", "leaf_axis_start",
                    "leaf_axis_edge", "leaf_data"):
            x, y = getattr(a, fld), getattr(b, fld)
            assert x.dtype == y.dtype and np.array_equal(x, y), fld
        assert a.leaf_tensor_index == b.leaf_tensor_index and a.n_slice_groups == b.n_slice_groups


def test_compiled_flatten_error_paths_and_foreign_arrays():
    from tensororder_b200 import flatten
    from tensororder_b200.plan_format import PlanNetwork, PlanTensor, PlanTree, StoredExecutionPlan

    class OwnArray(PlanTensor):  # build() that ignores the factory's buffer (allowed by the reference interface)
        def build(self, tensor_factory):
            return np.array(self._data, dtype=np.float32)

    for cls in (PlanTensor, OwnArray):
        tensors = [cls((2, 2), [[1, 2], [3, 4]], False, "a"), cls((2, 2), [[1, 0], [0, 1]], False, "b")]
        net = PlanNetwork(tensors, [[0, 1], [0, 1]], [[0, 1], [0, 1]])
        plan = StoredExecutionPlan(PlanTree([("leaf", 0), ("leaf", 1), ("join", 0, 1)]), net, [])
        for compiled in (True, False):
            flatten.USE_COMPILED = compiled
            try:
                f = flatten.flatten_plan(plan)
            finally:
                flatten.USE_COMPILED = True
            assert f.leaf_data.tolist() == [1, 2, 3, 4, 1, 0, 0, 1] and f.node_left.tolist() == [-1, -1, 0]
    bad = StoredExecutionPlan(PlanTree([("leaf", 0), ("leaf", 1)]), net, [])  # two roots
    dangling = StoredExecutionPlan(PlanTree([("leaf", 0), ("leaf", 1), ("join", 0, 1)]),
                                   PlanNetwork(tensors, [[0, -1], [0, 1]], [[0, 1], [0, 1]]), [])
    for compiled in (True, False):
        flatten.USE_COMPILED = compiled
        try:
            with pytest.raises(ValueError, match="single rooted tree"):
                flatten.flatten_plan(bad)
            with pytest.raises(ValueError, match="dangling"):
                flatten.flatten_plan(dangling)
        finally:
            flatten.USE_COMPILED = True


def test_plan_errors():
    pp = load_golden("toy_path_lineflow")
    plan = pp.as_execution_plan()
    flat = flatten_plan(plan)
    # open index: drop the last join so the root keeps free edges
    bad = flatten_plan(plan)
    bad.node_left = bad.node_left[:-1].copy()
    bad.node_right = bad.node_right[:-1].copy()
    bad.node_leaf = bad.node_leaf[:-1].copy()
    with pytest.raises(ValueError):
        CompiledPlan(bad)
    cp = CompiledPlan(flat)
    if cabi.lib.tob_device_count() == 0:
        with pytest.raises(RuntimeError, match="CUDA device"):
            cp.upload()  # no CPU fallback: fails loudly without a GPU
    cp.close()
