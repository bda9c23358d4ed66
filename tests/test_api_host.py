"""Host logic of `B200API` with the device stage replaced by the numpy interpreter of the compiled program
(tests/program_sim.py): plan cache, entry types, exact mode with mixed signs."""
import numpy as np
import pytest

from conftest import load_golden
from program_sim import install_fake_device
from tensororder_b200 import api as api_mod
from tensororder_b200.api import PLAN_CACHE, B200API


@pytest.fixture
def fake(monkeypatch):
    PLAN_CACHE.clear()
    PLAN_CACHE.hits = PLAN_CACHE.misses = 0
    yield install_fake_device(api_mod.CompiledPlan, monkeypatch.setattr)
    PLAN_CACHE.clear()


def _api(entry_type="float64", **kw):
    api = B200API()
    api.add_argument("entry_type", entry_type)
    for k, v in kw.items():
        api.add_argument(k, v)
    return api


def test_plan_cache_hits_on_the_same_plan_and_rereads_the_leaves(fake):
    pp = load_golden("vc50_mcc_lineflow")
    plan = pp.as_execution_plan()
    want = pp.expected["count"]
    a = _api()
    assert float(a.contract_sliced(plan)) == pytest.approx(want, rel=1e-12)
    assert a.last_stats["plan_cache_hit"] is False
    b = _api()  # a different API object: the cache is process-wide, like the CLI's one-shot API objects
    assert float(b.contract_sliced(plan)) == pytest.approx(want, rel=1e-12)
    assert b.last_stats["plan_cache_hit"] is True
    # the caller's tensors are re-read on a hit: change one weight, the count follows
    t = next(i for i, x in enumerate(pp.tensors) if len(x["shape"]) >= 1 and any(v not in (0.0, 1.0) for v in x["data"]))
    plan.network[t]._data *= 2.0
    assert float(_api().contract_sliced(plan)) == pytest.approx(2.0 * want, rel=1e-12)
    assert PLAN_CACHE.hits == 2 and PLAN_CACHE.misses == 1


def test_plan_cache_misses_when_the_slicing_or_the_options_change(fake):
    pp = load_golden("vc50_lineflow")
    plan = pp.as_execution_plan()
    assert float(_api().contract_sliced(plan)) == 2802717837.0
    sliced = pp.variant("min3").as_execution_plan()
    plan.groups_to_slice = sliced.groups_to_slice  # what slicer.slice_once does to the SAME plan object (execution.py:140-142)
    api = _api()
    assert float(api.contract_sliced(plan)) == 2802717837.0
    assert api.last_stats["plan_cache_hit"] is False and api.last_stats["slices"] == 8
    api = _api(hoist_invariant=False)
    assert float(api.contract_sliced(plan)) == 2802717837.0
    assert api.last_stats["plan_cache_hit"] is False
    api = _api(plan_cache=False)
    assert float(api.contract_sliced(plan)) == 2802717837.0
    assert api.last_stats["plan_cache_hit"] is False
    assert float(_api().contract_sliced(plan)) == 2802717837.0 and PLAN_CACHE.hits == 1


def test_plan_cache_is_bounded(fake, monkeypatch):
    monkeypatch.setattr(api_mod.PlanCache, "MAX_ENTRIES", 3)
    pp = load_golden("toy_path_lineflow")
    plans = [pp.as_execution_plan() for _ in range(6)]
    for p in plans:
        _api().contract_sliced(p)
    assert len(PLAN_CACHE.entries) == 3


@pytest.mark.parametrize("entry_type,np_type", [("int", np.int64), ("uint", np.uint64)])
def test_integer_entry_types_wrap_like_numpy(fake, entry_type, np_type):
    """numpy's int64 / uint64 tensordot wraps modulo 2^64 (numpy_apis.py:19-20): n=100 has 8.2e18 < 2^63 covers,
    n=150 (2.3e28) wraps several times."""
    for name in ("vc50_lineflow", "vc100_lineflow", "vc150_lineflow"):
        pp = load_golden(name)
        if "count_exact" not in pp.expected or pp.expected["maxrank"] > 15:
            continue
        exact = int(pp.expected["count_exact"])
        got = _api(entry_type).contract_sliced(pp.as_execution_plan())
        assert isinstance(got, np_type)
        want = exact % (1 << 64)
        if entry_type == "int" and want >= (1 << 63):
            want -= 1 << 64
        assert int(got) == want


def test_integer_entry_types_truncate_weights_like_numpy(fake):
    pp = load_golden("vc50_mcc_lineflow")  # weights in (0.5, 1.5): int64 leaves hold 0 or 1
    api = _api("int")
    got = api.contract_sliced(pp.as_execution_plan())
    # the reference semantics: every leaf built with dtype=int64, i.e. truncated toward zero, then exact arithmetic
    from oracle import numpy_oracle

    doc = pp.to_json()
    for t in doc["tensors"]:
        t["data"] = [float(np.trunc(v)) for v in t["data"]]
        t.pop("params", None)
    assert int(got) == int(numpy_oracle.contract_sliced(doc))
    with pytest.raises(ValueError):
        bad = load_golden("vc50_mcc_lineflow")
        bad.tensors[0]["data"] = [-1.0 * v - 1.0 for v in bad.tensors[0]["data"]]
        _api("uint").contract_sliced(bad.as_execution_plan())


@pytest.mark.parametrize("entry_type,np_type,tol", [("float32", np.float32, 5e-6), ("float16", np.float16, 5e-3)])
def test_reduced_precision_entry_types(fake, entry_type, np_type, tol):
    pp = load_golden("vc50_mcc_lineflow")
    got = _api(entry_type).contract_sliced(pp.as_execution_plan())
    assert isinstance(got, np_type)
    if np.isfinite(got):
        assert float(got) == pytest.approx(pp.expected["count"], rel=tol * 50)  # leaf rounding: ~125 leaves
    got = _api("float32").contract_sliced(load_golden("vc50_lineflow").as_execution_plan())
    # correctly rounded; the reference's float32 GEMMs land one ulp (256) below: `Count: 2.8027177e+09`
    assert got == np.float32(2802717837.0) and abs(float(got) - 2.8027177e9) <= 5e-6 * 2.8e9


def test_unknown_entry_type_and_argument_raise_like_the_reference():
    api = B200API()
    with pytest.raises(ValueError, match="Unknown b200 type"):
        api.add_argument("entry_type", "complex128")
    with pytest.raises(ValueError, match="Invalid argument"):
        api.add_argument("TPU", "x")
    for et in ("float64", "float32", "float16", "uint", "int", "bigint"):
        api.add_argument("entry_type", et)
        assert api.create_tensor((2, 2), 1).dtype == api._HOST_DTYPES[et]
        assert api.get_entry_size() == 8


def test_exact_mode_with_mixed_signs_uses_the_a_priori_bound(fake):
    """Integer weights of both signs: the float64 pass may cancel to anything, so the number of primes comes
    from prod_t ||tensor_t||_1 and the CRT value is mapped into the symmetric range."""
    from oracle import numpy_oracle

    pp = load_golden("vc50_lineflow")
    rng = np.random.default_rng(7)
    for t in pp.tensors:
        if len(t["shape"]) == 3:  # variable tensors: diag(w-, w+) with random signed integer weights
            d = [0.0] * 8
            d[0], d[7] = float(rng.integers(-3, 4)), float(rng.integers(-3, 4))
            t["data"] = d
            t.pop("params", None)
    doc = pp.to_json()
    for t in doc["tensors"]:
        t.pop("params", None)
    # exact reference: the oracle on Python ints (object arrays), as the reference's bigint mode does
    want = numpy_oracle.contract_sliced(doc, dtype=object) if "dtype" in numpy_oracle.contract_sliced.__code__.co_varnames else None
    api = _api("bigint")
    got = api.contract_sliced(pp.as_execution_plan())
    assert isinstance(got, int)
    assert api.last_stats["float_estimate"] is None  # no float64 sizing pass with mixed signs
    if want is not None:
        assert got == int(want)
    else:  # float64 agrees to rounding unless it cancels; the sign and magnitude must match
        approx = float(numpy_oracle.contract_sliced(doc))
        assert got == pytest.approx(approx, rel=1e-9, abs=1e3)
    sliced = load_golden("vc50_lineflow").variant("min3")
    sliced.tensors = pp.tensors
    assert _api("bigint").contract_sliced(sliced.as_execution_plan()) == got  # slicing-invariant


def test_plan_cache_with_two_host_threads(fake):
    """Independent plans (and even the same plan object) contracted from two host threads through the public call:
    a cached plan that is running is neither shared nor evicted."""
    from concurrent.futures import ThreadPoolExecutor

    a = load_golden("vc50_lineflow").as_execution_plan()
    b = load_golden("vc50_mcc_lineflow").as_execution_plan()
    want_b = load_golden("vc50_mcc_lineflow").expected["count"]

    def many(plan, n):
        return [float(_api().contract_sliced(plan)) for _ in range(n)]

    with ThreadPoolExecutor(max_workers=3) as pool:
        fa, fb, fa2 = pool.submit(many, a, 6), pool.submit(many, b, 6), pool.submit(many, a, 6)
        assert all(v == 2802717837.0 for v in fa.result() + fa2.result())
        assert all(v == pytest.approx(want_b, rel=1e-12) for v in fb.result())
    assert all(e["compiled"].busy == 0 for e in PLAN_CACHE.entries.values())


def test_collective_order_serves_tickets_in_order_and_rejects_stale_ones():
    """api.CollectiveOrder: threads pass the turnstile in ticket order whatever their arrival order; a ticket that was
    already served is an error rather than a silent re-ordering."""
    import threading
    import time

    from tensororder_b200.api import CollectiveOrder

    order, served = CollectiveOrder(), []

    def worker(ticket, delay):
        time.sleep(delay)
        order.enter(ticket)
        served.append(ticket)
        order.leave(ticket)

    threads = [threading.Thread(target=worker, args=(t, 0.02 * (4 - t))) for t in range(5)]  # arrive 4, 3, 2, 1, 0
    for t in threads:
        t.start()
    for t in threads:
        t.join(10)
        assert not t.is_alive()
    assert served == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        order.enter(2)
    order.enter(5)
    order.leave(5)
    api = B200API()
    with pytest.raises(ValueError):
        api.add_argument("collective_ticket", (object(), 0))
