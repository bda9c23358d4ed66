"""Exact mode (`entry_type=bigint`): residues modulo ~23-bit primes on the same kernels + CRT on the host.
Golden exact counts come from the reference's own `--entry_type=bigint` numpy run (Python ints in object
arrays; tests/golden/ref_replay.py --bigint)."""
import math

import numpy as np
import pytest

from conftest import golden_names, load_golden
from program_sim import run_program
from tensororder_b200 import api as api_mod
from tensororder_b200.api import EXACT_PRIMES, B200API, crt

EXACT = [n for n in golden_names() if "count_exact" in load_golden(n).expected]


def test_primes_and_crt():
    assert len(set(EXACT_PRIMES)) == len(EXACT_PRIMES) and max(EXACT_PRIMES) < 2 ** 23 and min(EXACT_PRIMES) > 2 ** 22
    for p in EXACT_PRIMES[:5] + EXACT_PRIMES[-5:]:
        assert all(p % d for d in range(2, int(p ** 0.5) + 1))
    # 128 products of residues stay exact in float64: the invariant every kernel relies on
    assert 128 * (max(EXACT_PRIMES) - 1) ** 2 < 2 ** 53
    x = 44255948944721021411806 * 10 ** 20 + 12345678901234567890
    ms = EXACT_PRIMES[:8]
    assert crt([x % p for p in ms], ms) == x


def _fake_device(monkeypatch):
    """Replace the device stage by the numpy interpreter of the compiled program (CPU test of the host logic)."""
    from program_sim import install_fake_device

    api_mod.PLAN_CACHE.clear()
    return install_fake_device(api_mod.CompiledPlan, monkeypatch.setattr)


@pytest.mark.parametrize("name", [n for n in EXACT if load_golden(n).expected["maxrank"] <= 15])
def test_exact_count_host_logic(name, monkeypatch):
    _fake_device(monkeypatch)
    pp = load_golden(name)
    api = B200API()
    api.add_argument("entry_type", "bigint")
    got = api.contract_sliced(pp.as_execution_plan())
    assert isinstance(got, int) and got == int(pp.expected["count_exact"])
    assert api.last_stats["exact_passes"] >= 1


def test_exact_mode_rejects_non_integer_weights(monkeypatch):
    _fake_device(monkeypatch)
    api = B200API()
    api.add_argument("entry_type", "bigint")
    with pytest.raises(ValueError, match="integer"):
        api.contract_sliced(load_golden("vc50_mcc_lineflow").as_execution_plan())


@pytest.mark.gpu
@pytest.mark.parametrize("name", EXACT)
def test_exact_counts_on_the_gpu(name):
    pp = load_golden(name)
    api = B200API()
    api.add_argument("entry_type", "bigint")
    got = api.contract_sliced(pp.as_execution_plan())
    assert isinstance(got, int) and got == int(pp.expected["count_exact"]), (got, pp.expected["count_exact"])
    for v in pp.variants:
        if "count_exact" in v.get("expected", {}):
            assert api.contract_sliced(pp.variant(v["name"]).as_execution_plan()) == int(v["expected"]["count_exact"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vc200_lineflow", "vc230_lineflow"])
def test_exact_counts_beyond_the_reference(name):
    """No exact reference exists at this size (Python-int numpy would need hours): size-independent
    properties instead — the exact integer agrees with the float64 count to float precision, is stable when
    a prime is added (built into the reconstruction), and is identical for a sliced plan of the same tree."""
    pp = load_golden(name)
    api = B200API()
    api.add_argument("entry_type", "bigint")
    exact = api.contract_sliced(pp.as_execution_plan())
    assert isinstance(exact, int) and exact > 2 ** 53
    assert math.isclose(float(exact), pp.expected["count"], rel_tol=1e-12)
    assert api.last_stats["exact_passes"] >= math.ceil(math.log2(exact) / 23)
    if name == "vc200_lineflow":
        assert api.contract_sliced(pp.variant("min3").as_execution_plan()) == exact


@pytest.mark.gpu
@pytest.mark.parametrize("name,variant", [("vc100_lineflow", None), ("vc100_lineflow", "min4")])
def test_modulus_change_on_a_resident_plan(name, variant):
    """A resident plan whose slices (and slice-invariant prologue) replay as CUDA graphs: the graphs carry the
    modulus in their kernel parameters, so changing it must re-capture them.  0/1 leaves are residues for every
    prime, so the same upload serves all passes."""
    from tensororder_b200.api import CompiledPlan, EXACT_PRIMES
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden(name)
    want = int(pp.expected["count_exact"])
    if variant:
        pp = pp.variant(variant)
    flat = flatten_plan(pp.as_execution_plan())
    for graph in (1, 2, 0):
        cp = CompiledPlan(flat, use_graph=graph)
        cp.upload()
        for p in (EXACT_PRIMES[0], EXACT_PRIMES[1], 0, EXACT_PRIMES[2]):
            cp.set_modulus(p)
            for _ in range(3):  # auto mode switches to graph replay on the second run
                got = cp.run()
                if p:
                    assert got == want % p, (graph, p, got, want % p)
                else:
                    assert math.isclose(got, float(want), rel_tol=1e-12)
        cp.close()
