"""Parity of the CUDA path (through the C ABI) against the reference's golden counts and the oracle.
Run on the B200 box:  python -m pytest tests -m gpu"""
import math

import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

ALL = golden_names()
REL = 1e-9  # north_star tolerance for float64 counts that are not exactly representable / weighted


def _api(**kw):
    from tensororder_b200.api import B200API

    api = B200API()
    api.add_argument("entry_type", "float64")
    for k, v in kw.items():
        api.add_argument(k, v)
    return api


def _check(got, exp):
    want = exp["count"]
    if want < 2 ** 53 and float(want).is_integer():
        assert float(got) == want, (got, want)  # unweighted, exactly representable: bit exact
    else:
        assert math.isclose(float(got), want, rel_tol=REL), (got, want)


def test_readme_instance():
    pp = load_golden("vc50_lineflow")
    got = _api().contract_sliced(pp.as_execution_plan())
    assert str(got) == "2802717837.0"  # what `Count:` prints, tensororder.py:299


def _unsliced_cases():
    """(name, kernel policy): every fixture with a stored count on the dispatch table; the generic-kernels-only policy
    (a debug path, far too slow for the big GEMM nodes) up to max-rank 24."""
    names = [n for n in ALL if "count" in load_golden(n).expected]
    return [(n, 0) for n in names] + [(n, 1) for n in names if load_golden(n).expected["maxrank"] <= 24]


@pytest.mark.parametrize("name,policy", _unsliced_cases())
def test_unsliced_counts(name, policy):
    pp = load_golden(name)
    api = _api(kernel_policy=policy)
    got = api.contract_sliced(pp.as_execution_plan())
    _check(got, pp.expected)
    assert api.last_stats["launches"] > 0


def _variants():
    out = []
    for name in ALL:
        pp = load_golden(name)
        for i, v in enumerate(pp.variants):
            if "count" in v.get("expected", {}):
                out.append((name, i))
    return out


@pytest.mark.parametrize("name,vi", _variants())
@pytest.mark.parametrize("graph", [True, False])
def test_sliced_counts(name, vi, graph):
    pp = load_golden(name).variant(vi)
    api = _api(use_graph=graph)
    got = api.contract_sliced(pp.as_execution_plan())
    _check(got, pp.expected)
    if "slice_cutoff" in pp.expected:
        cut = api.contract_sliced(pp.as_execution_plan(), num_slice_limit=pp.expected["slice_cutoff"])
        assert math.isclose(float(cut), pp.expected["count_cutoff"], rel_tol=REL)


@pytest.mark.parametrize("name,vi", [x for x in _variants() if "per_slice" in load_golden(x[0]).variants[x[1]]["expected"]][:8])
def test_per_slice_results_and_partition(name, vi):
    from tensororder_b200.api import CompiledPlan
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden(name).variant(vi)
    cp = CompiledPlan(flatten_plan(pp.as_execution_plan()), hoist_invariant=True)
    cp.upload()
    per = pp.expected["per_slice"]
    for s in range(min(len(per), 16)):
        assert math.isclose(cp.run(first=s, count=1), per[s], rel_tol=REL, abs_tol=0.0)
    for world in (2, 4):
        if len(per) >= world:
            parts = [cp.run(first=r, stride=world) for r in range(world)]
            assert math.isclose(sum(parts), pp.expected["count"], rel_tol=REL)
    cp.close()


def test_hoisting_does_not_change_result():
    pp = load_golden("vc100_lineflow").variant("min4")
    a = _api(hoist_invariant=True).contract_sliced(pp.as_execution_plan())
    b = _api(hoist_invariant=False).contract_sliced(pp.as_execution_plan())
    assert math.isclose(float(a), float(b), rel_tol=1e-12)
    _check(a, pp.expected)


@pytest.mark.parametrize("name,variant", [("vc100_lineflow", None), ("vc150_lineflow", None), ("vc150_lineflow", "min4"),
                                          ("vc190_lineflow", None), ("vc200_lineflow", "min3"),
                                          ("vc150_mcc_factorflow", "min3"), ("rand3cnf_24_lineflow", "min3"),
                                          ("rand4cnf_18_mcc_lineflow", None)])
def test_dag_schedule_is_bit_identical_to_post_order(name, variant):
    """Independent subtrees run concurrently on up to 16 streams per lane (fork/join by events, captured as a
    DAG when the slice replays as a graph).  Every join computes the same sums in the same order whatever
    runs beside it, so the count equals the one-stream post-order run bit for bit — in stream mode, as
    graphs, on the first run and on replays of a resident plan."""
    from tensororder_b200.api import CompiledPlan
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden(name)
    if variant:
        pp = pp.variant(variant)
    flat = flatten_plan(pp.as_execution_plan())
    results = {}
    for branches in (1, 0, 3, 32):
        for graph in (0, 1, 2):
            cp = CompiledPlan(flat, dag_branches=branches, use_graph=graph)
            cp.upload()
            got = [cp.run() for _ in range(3)]
            assert got[0] == got[1] == got[2], (branches, graph, got)
            results[(branches, graph)] = got[0]
            if branches != 1:
                assert cp.describe()["branches"] > 1
            cp.close()
    assert len(set(results.values())) == 1, results
    want = pp.expected.get("count", load_golden(name).expected.get("count"))
    if want is not None:
        assert math.isclose(results[(0, 2)], want, rel_tol=REL)


@pytest.mark.parametrize("shape", ["random", "caterpillar", "balanced"])
@pytest.mark.parametrize("seed", range(12))
def test_random_networks_and_trees(seed, shape):
    """Random closed networks / trees / slicings (tests/random_plans.py; the CPU suite runs the same cases
    through the program simulator): the CUDA path equals the plain numpy contraction exactly (small-integer
    tensors), with every executor feature on and off, as stream launches and as graphs."""
    from random_plans import make
    from tensororder_b200.api import CompiledPlan

    rng = np.random.default_rng(1000 + seed)
    n_tensors = int(rng.integers(3, 40))
    n_edges = int(n_tensors * rng.uniform(1.0, 1.8))
    groups = int(rng.integers(0, 4))
    flat, want = make(seed, n_tensors=n_tensors, n_edges=n_edges, n_slice_groups=min(groups, n_edges), shape=shape)
    for kw in ({}, {"use_graph": 1}, {"use_graph": 0, "dag_branches": 1}, {"use_microtree": False},
               {"hoist_invariant": False, "use_graph": 1}, {"kernel_policy": 1, "dag_branches": 5}):
        cp = CompiledPlan(flat, **kw)
        cp.upload()
        assert cp.run() == want, kw
        assert cp.run() == want, kw  # replay (graph instantiated on the second run in auto mode)
        if cp.num_slices >= 2:
            assert cp.run(first=0, stride=2) + cp.run(first=1, stride=2) == want, kw
        cp.close()
    flat_w, want_w = make(seed, n_tensors=n_tensors, n_edges=n_edges, n_slice_groups=min(groups, n_edges), shape=shape,
                          integer=False)
    cp = CompiledPlan(flat_w)
    cp.upload()
    assert math.isclose(cp.run(), want_w, rel_tol=REL)
    cp.close()


@pytest.mark.parametrize("name", [n for n in ALL if "count" not in load_golden(n).expected])
def test_large_instances_slicing_invariance(name):
    """No reference count is stored for the largest family members (numpy needs minutes and tens of
    GB); size-independent property instead: the count is invariant under the reference slicer's
    slicings of the same tree."""
    pp = load_golden(name)
    base = float(_api().contract_sliced(pp.as_execution_plan()))  # n=250 unsliced: 30.6 GB arena, ~2.6 s
    assert base > 0 and math.isfinite(base)
    for i, v in enumerate(pp.variants):
        got = float(_api().contract_sliced(pp.variant(i).as_execution_plan()))
        assert math.isclose(got, base, rel_tol=REL), (v["name"], got, base)
        # where the reference itself replayed a sliced plan of this instance (tests/golden/ref_replay.py: n=240, 250
        # with 8 slices), the UNSLICED device count is held against that reference count too
        ref_count = v.get("expected", {}).get("count")
        if ref_count is not None:
            assert math.isclose(base, ref_count, rel_tol=REL), (v["name"], base, ref_count)


@pytest.mark.parametrize("name", ["vc50_lineflow", "vc120_lineflow", "vc150_mcc_factorflow", "toy_unit_neg_lineflow"])
def test_microtree_kernel_matches_per_join_launches(name):
    pp = load_golden(name)
    on = _api(use_microtree=True)
    off = _api(use_microtree=False)
    a = on.contract_sliced(pp.as_execution_plan())
    b = off.contract_sliced(pp.as_execution_plan())
    assert math.isclose(float(a), float(b), rel_tol=1e-12)
    _check(a, pp.expected)
    assert on.last_stats["launches"] < off.last_stats["launches"]


def test_interruptible_run_is_bit_identical_and_honours_sigalrm():
    """The slice loop returns to the interpreter between chunks (the reference's TimeoutTimer is a SIGALRM
    handler raising TimeoutError, src/util/util.py:32-39) and carries the device accumulator across
    chunks, so the result equals the single-call sequential sum bit for bit."""
    import signal
    import time

    from tensororder_b200.api import CompiledPlan
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden("vc150_lineflow").variant("min4")
    cp = CompiledPlan(flatten_plan(pp.as_execution_plan()))
    cp.upload()
    one = cp.run()
    chunked = cp.run_interruptible(target_s=0.0005)
    assert float(one).hex() == float(chunked).hex()
    cp.close()

    big = load_golden("vc250_lineflow").variant("min6")  # 64 slices, ~3 s of device work
    api = _api()

    def handler(signum, frame):
        raise TimeoutError()

    old = signal.signal(signal.SIGALRM, handler)
    signal.setitimer(signal.ITIMER_REAL, 0.3)
    t0 = time.time()
    try:
        with pytest.raises(TimeoutError):
            api.contract_sliced(big.as_execution_plan())
    finally:
        signal.setitimer(signal.ITIMER_REAL, 0)
        signal.signal(signal.SIGALRM, old)
    assert time.time() - t0 < 1.5  # interrupted after a chunk, not after all 64 slices


@pytest.mark.parametrize("name,variant", [("vc100_lineflow", "min4"), ("vc150_lineflow", "min4"), ("vc50_mcc_lineflow", "min3"),
                                          ("vc200_lineflow", "min3")])
def test_two_slice_lanes_give_the_sequential_sum(name, variant):
    """Slices run two at a time on separate streams/arenas; per-slice results are summed afterwards in
    slice order, so the count is bit-identical to the one-lane run (and to the reference's loop order)."""
    from tensororder_b200.api import CompiledPlan
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden(name).variant(variant)
    flat = flatten_plan(pp.as_execution_plan())
    one = CompiledPlan(flat, slice_lanes=1)
    two = CompiledPlan(flat, slice_lanes=2)
    assert one.describe()["lanes"] == 1 and two.describe()["lanes"] == 2
    one.upload()
    two.upload()
    a, b = one.run(), two.run()
    assert float(a).hex() == float(b).hex()
    _check(b, pp.expected if "count" in pp.expected else load_golden(name).expected)
    assert float(two.run()).hex() == float(a).hex()  # second run: graph replay on both lanes
    assert float(two.run(first=1, count=5, stride=1)).hex() == float(one.run(first=1, count=5, stride=1)).hex()
    one.close()
    two.close()


def test_contract_single_network_entry():
    pp = load_golden("vc50_factorflow")
    plan = pp.as_execution_plan()
    res = _api().contract(plan.network, plan.tree)
    assert res[tuple()] == 2802717837.0  # base_api.py:26-27 indexes the result with ()


def test_out_of_memory_maps_to_reference_error():
    from tensororder_b200.api import CompiledPlan, OutOfMemoryError
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden("vc150_lineflow")
    cp = CompiledPlan(flatten_plan(pp.as_execution_plan()), mem_limit_bytes=1 << 20)
    with pytest.raises(OutOfMemoryError):
        cp.upload()
    cp.close()


@pytest.mark.parametrize("name,variant", [("vc150_lineflow", None), ("vc200_lineflow", "min3"), ("vc220_lineflow", None)])
def test_count_is_multilinear_in_a_variable_weight(name, variant):
    """Size-independent property at BASELINE sizes (no reference count needed): the weighted count is linear in
    the weight pair of every variable, Z(w-, w+) = w- * Z(x=0) + w+ * Z(x=1).  A variable's leaf is the copy
    tensor diag(w-, w+) (VariableTensor.build, tensor_network_constructions.py:144-152): rewrite it in the
    flat plan's leaf buffer and contract three times."""
    import dataclasses

    from tensororder_b200.api import CompiledPlan
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden(name)
    if variant:
        pp = pp.variant(variant)
    flat = flatten_plan(pp.as_execution_plan())
    rng = np.random.default_rng(7)
    picked = 0
    for leaf in rng.permutation(flat.n_leaves):
        r = int(flat.leaf_rank[leaf])
        off = int(flat.leaf_data_offset[leaf])
        d = flat.leaf_data[off: off + (1 << r)]
        if r < 2 or np.count_nonzero(d) != 2 or d[0] != 1.0 or d[-1] != 1.0:
            continue  # not a copy tensor
        results = {}
        for w in ((1.0, 0.0), (0.0, 1.0), (0.375, 1.75)):
            data = flat.leaf_data.copy()
            data[off], data[off + (1 << r) - 1] = w
            cp = CompiledPlan(dataclasses.replace(flat, leaf_data=data))
            cp.upload()
            results[w] = cp.run()
            cp.close()
        z0, z1, z = results[(1.0, 0.0)], results[(0.0, 1.0)], results[(0.375, 1.75)]
        assert z0 >= 0 and z1 > 0
        assert math.isclose(z, 0.375 * z0 + 1.75 * z1, rel_tol=REL), (leaf, z, z0, z1)
        assert math.isclose(z0 + z1, load_golden(name).expected["count"], rel_tol=REL)
        picked += 1
        if picked == 2:
            break
    assert picked == 2


def test_host_threads_contract_concurrently():
    """Eight host threads, three rounds, instances of every size class through the public call at once (the bench's e2e
    arm): plans on their own streams, the shared block / stream pools, the plan cache, stream-K flag blocks of concurrent
    plans.  Every count must match and repeat bit for bit; a stalled pool shows up as the timeout."""
    from concurrent.futures import ThreadPoolExecutor

    jobs = [("vc200_lineflow", "min3"), ("vc150_mcc_factorflow", None), ("vc190_lineflow", None), ("vc180_lineflow", None),
            ("vc160_lineflow", None), ("vc150_lineflow", "min4"), ("vc120_lineflow", None), ("vc100_lineflow", "min4"),
            ("vc50_lineflow", None), ("vc50_mcc_lineflow", "min3"), ("vc170_lineflow", None), ("vc140_lineflow", None)]
    plans, want = [], []
    for name, variant in jobs:
        pp = load_golden(name)
        if variant:
            pp = pp.variant(variant)
        plans.append(pp.as_execution_plan())
        want.append(pp.expected.get("count", load_golden(name).expected.get("count")))

    def call(i):
        return float(_api().contract_sliced(plans[i]))

    first = None
    with ThreadPoolExecutor(max_workers=8) as pool:
        for _ in range(3):
            got = list(pool.map(call, range(len(jobs)), timeout=120))
            for g, w, job in zip(got, want, jobs):
                assert math.isclose(g, w, rel_tol=REL), (job, g, w)
            if first is None:
                first = got
            assert [x.hex() for x in got] == [x.hex() for x in first]
