"""Per-op device times of one stored plan (CUDA events around every op) with the per-node roofline
fractions (SURVEY.md §8d work model).  Usage: python tools/profile_ops.py vc250_lineflow [variant] [top]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tensororder_b200.api import CompiledPlan
from tensororder_b200.flatten import flatten_plan
from tensororder_b200.plan_format import PortablePlan

HBM = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6650.0
FP64 = 35.49
name = sys.argv[1]
variant = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
pp = PortablePlan.load(os.path.join(REPO, "tests", "golden", name + ".json.gz"))
if variant:
    pp = pp.variant(variant)
cp = CompiledPlan(flatten_plan(pp.as_execution_plan()))
cp.upload()
d = cp.describe()
ops = d["invariant_ops"] + d["slice_ops"]
cp.profile(0)
ms, res = cp.profile(0)
tot = sum(ms)
print("%s%s: %d ops, one slice %.3f ms, count(slice 0) %r, peak %.2f GB" % (name, "/" + variant if variant else "", len(ops), tot, res, cp.peak_bytes / 1e9))
rows = sorted(zip(ms, ops), key=lambda x: -x[0])[:top]
out = []
for t, op in rows:
    if op["kind"] in (2, 3):
        if op["kind"] == 3:
            print("  microtree launch: %d joins in %d CTAs  %9.4f ms  %5.1f%% of slice" % (
                len(op["micro"]), len(op["cta_start"]) - 1, t, 100 * t / tot))
        continue
    tf = op["flops"] / (t * 1e-3) / 1e12
    gb = op["bytes"] / (t * 1e-3) / 1e9
    bound = "tensor" if op["flops"] / (FP64 * 1e12) > op["bytes"] / (HBM * 1e9) else "hbm"
    frac = tf / FP64 if bound == "tensor" else gb / HBM
    kind = {0: "generic/%d" % op["threads_per_out"], 1: "gemm%dx%d" % (1 << op["tm_log2"], 1 << op["tn_log2"])}[op["kind"]]
    print("  m=%2d n=%2d k=%2d %-12s ks=%d  %9.4f ms  %5.1f%% of slice  %7.2f TF/s %8.1f GB/s  bound=%-6s frac=%.3f" % (
        op["m"], op["n"], op["k"], kind, op["ksplit_log2"], t, 100 * t / tot, tf, gb, bound, frac))
    out.append({"m": op["m"], "n": op["n"], "k": op["k"], "kernel": kind, "ksplit_log2": op["ksplit_log2"], "ms": t,
                "tflops": tf, "gbs": gb, "bound": bound, "frac": frac})
by_kind = {}
for t, op in zip(ms, ops):
    key = {0: "generic", 1: "gemm", 2: "accum", 3: "microtree"}[op["kind"]]
    by_kind.setdefault(key, [0, 0.0])
    by_kind[key][0] += 1
    by_kind[key][1] += t
print("  by kernel:", {k: (v[0], round(v[1], 3)) for k, v in by_kind.items()})
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump({"plan": name, "variant": variant, "slice_ms": tot, "top": out, "by_kind": by_kind},
          open(os.path.join(REPO, "gpurun_out", "ops_%s%s.json" % (name, "_" + variant if variant else "")), "w"), indent=1)
