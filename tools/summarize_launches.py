"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/..._summary.txt"""
import collections, csv, re, sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = None
for r in rd:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    if len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void\s+", "", r["Kernel Name"])
    name = re.sub(r"\(.*$", "", name).replace("tob::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name][0] += 1
    tot[name][1] += ms
total = sum(v[1] for v in tot.values())
print("%-62s %9s %12s %7s" % ("kernel / graph", "launches", "total ms", "share"))
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %9d %12.3f %6.1f%%" % (name[:62], n, ms, 100 * ms / total if total else 0))
print("%-62s %9d %12.3f" % ("TOTAL", sum(v[0] for v in tot.values()), total))
