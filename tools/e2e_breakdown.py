"""Where the end-to-end time of B200API.contract_sliced goes, per instance (host stages vs device)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tensororder_b200.api import B200API
from tensororder_b200.plan_format import PortablePlan

names = sys.argv[1:] or ["vc50_lineflow", "vc100_lineflow", "vc150_lineflow", "vc190_lineflow", "vc200_lineflow:min3", "vc220_lineflow:min3"]
for spec in names:
    name, _, var = spec.partition(":")
    pp = PortablePlan.load(os.path.join(REPO, "tests", "golden", name + ".json.gz"))
    if var:
        pp = pp.variant(var)
    plan = pp.as_execution_plan()
    best = None
    for rep in range(4):
        api = B200API()
        api.add_argument("entry_type", "float64")
        t0 = time.perf_counter()
        api.contract_sliced(plan)
        dt = time.perf_counter() - t0
        st = dict(api.last_stats, total_s=dt, close_s=dt - api.last_stats["flatten_compile_s"] - api.last_stats["upload_s"] - api.last_stats["run_s"])
        if rep and (best is None or dt < best["total_s"]):
            best = st
    print("%-22s total %7.3f ms | flatten+compile %6.3f  upload %6.3f  run %7.3f (device %7.3f)  close %6.3f | launches %d" % (
        spec, best["total_s"] * 1e3, best["flatten_compile_s"] * 1e3, best["upload_s"] * 1e3, best["run_s"] * 1e3,
        best["device_ms"], best["close_s"] * 1e3, best["launches"]))
