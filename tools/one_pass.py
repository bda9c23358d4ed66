"""One sequential pass over the bench workload (the 18-instance cubic_vc family, resident plans) — the command `ncu`
wraps for the per-launch list.  Usage: python tools/one_pass.py [passes]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import load_workload  # noqa: E402
from tensororder_b200.api import CompiledPlan  # noqa: E402
from tensororder_b200.flatten import flatten_plan  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
items = load_workload(50, 220)
plans = []
for it in items:
    cp = CompiledPlan(flatten_plan(it["pp"].as_execution_plan()), use_graph=0)  # plain launches: every kernel is visible to ncu
    cp.upload()
    plans.append(cp)
total = 0.0
for _ in range(passes):
    for it, cp in zip(items, plans):
        got = cp.run()
        total += cp.last_ms
        assert abs(got - it["expected"]) <= 1e-9 * abs(it["expected"]), (it["name"], got)
print("one pass over %d instances: %.2f ms device time per pass, %d launches in the last instance" % (len(items), total / passes, plans[-1].last_launches))
