"""Diagnostic: one plan under tuning overrides with executor features toggled; prints the count per configuration.
Usage: python tools/diag_run.py NAME[:VARIANT] [key=value ...]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tensororder_b200 import cabi  # noqa: E402
from tensororder_b200.api import CompiledPlan  # noqa: E402
from tensororder_b200.flatten import flatten_plan  # noqa: E402
from tensororder_b200.plan_format import PortablePlan  # noqa: E402

spec = sys.argv[1]
quick = "--quick" in sys.argv
for a in [x for x in sys.argv[2:] if x != "--quick"]:
    key, value = a.split("=")
    assert cabi.lib.tob_tuning_set(key.encode(), float(value)) == 0, cabi.last_error()
name, _, variant = spec.partition(":")
pp = PortablePlan.load(os.path.join(REPO, "tests", "golden", name + ".json.gz"))
want = pp.expected.get("count")
if variant:
    pp = pp.variant(variant)
    want = pp.expected.get("count", want)
flat = flatten_plan(pp.as_execution_plan())
print("reference count", repr(want))
for kw in ({}, {"dag_branches": 1}) if quick else ({}, {"dag_branches": 1}, {"slice_lanes": 1}, {"dag_branches": 1, "slice_lanes": 1}, {"hoist_invariant": False},
           {"use_microtree": False}, {"use_graph": 1}, {"dag_branches": 1, "slice_lanes": 1, "hoist_invariant": False}):
    cp = CompiledPlan(flat, **kw)
    cp.upload()
    vals = [cp.run() for _ in range(3)]
    per_slice = [cp.run(first=s, count=1) for s in range(cp.num_slices)] if cp.num_slices <= 16 and not quick else []
    cp.close()
    print(kw, [repr(v) for v in vals], "rel err %.2e" % (abs(vals[0] - want) / abs(want)) if want else "")
    if per_slice:
        print("     per slice:", [repr(v) for v in per_slice], "sum", repr(sum(per_slice)))
