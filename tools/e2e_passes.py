"""Per-instance upload/run times over several passes of the whole workload (pool behaviour)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from tensororder_b200.api import B200API
items = bench.load_workload(50, 220)
plans = [it["pp"].as_execution_plan() for it in items]
for p in range(4):
    t0 = time.perf_counter(); ups = []
    for it, plan in zip(items, plans):
        api = B200API(); api.add_argument("entry_type", "float64")
        api.contract_sliced(plan)
        ups.append(api.last_stats["upload_s"] * 1e3)
    print("pass %d: %.1f ms; upload ms per instance: %s" % (p, (time.perf_counter() - t0) * 1e3, " ".join("%.2f" % u for u in ups)))
