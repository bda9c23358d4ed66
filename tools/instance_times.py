"""Device time of every instance of the bench workload, plans resident (the `value` arm of bench.py,
one instance at a time): total ms, DMMA-GEMM ms, the rest, launches.  For sliced plans also one slice
alone (first=0,count=1) so hoisted work and per-slice work separate.
Usage: python tools/instance_times.py [min_n max_n] [key=value ...]   (options go to CompiledPlan)"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tensororder_b200.api import CompiledPlan
from tensororder_b200.flatten import flatten_plan
from tensororder_b200.plan_format import PortablePlan

pos = [a for a in sys.argv[1:] if "=" not in a]
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[1:] if "=" in a)}
lo, hi = (int(pos[0]), int(pos[1])) if len(pos) >= 2 else (50, 220)
rows = []
tot = totg = 0.0
for n in range(lo, hi + 1, 10):
    pp = PortablePlan.load(os.path.join(REPO, "tests", "golden", "vc%d_lineflow.json.gz" % n))
    if 200 <= n <= 220:
        pp = pp.variant("min3")
    cp = CompiledPlan(flatten_plan(pp.as_execution_plan()), **kw)
    cp.upload()
    cp.set_gemm_timing(True)
    for _ in range(3):
        cp.run()
    best = None
    for _ in range(5):
        cp.run()
        if best is None or cp.last_ms < best[0]:
            best = (cp.last_ms, cp.last_gemm[0], cp.last_launches, cp.last_gemm[2], cp.last_issue_ms)
    one = None
    if cp.num_slices > 1:
        cp.run(first=0, count=1)
        cp.run(first=0, count=1)
        one = cp.last_ms
    d = cp.describe()
    rows.append({"n": n, "ms": best[0], "gemm_ms": best[1], "other_ms": best[0] - best[1], "launches": best[2],
                 "gemm_launches": best[3], "one_slice_ms": one, "ops": len(d["slice_ops"]), "invariant_ops": len(d["invariant_ops"])})
    tot += best[0]
    totg += best[1]
    print("n=%3d  %8.3f ms  gemm %8.3f  other %7.3f  launches %5d (gemm %3d)  ops/slice %3d  hoisted ops %3d  one slice %s  host issue %.3f ms" % (
        n, best[0], best[1], best[0] - best[1], best[2], best[3], len(d["slice_ops"]), len(d["invariant_ops"]),
        "%.3f" % one if one else "-", best[4]), flush=True)
    cp.close()
print("total %.3f ms  gemm %.3f  other %.3f" % (tot, totg, tot - totg))
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
tag = "_".join("%s%d" % (k, v) for k, v in sorted(kw.items())) or "default"
json.dump(rows, open(os.path.join(REPO, "gpurun_out", "instance_times_%s.json" % tag), "w"), indent=1)
