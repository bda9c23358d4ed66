"""Summarises `ncu --set full` captures (gpurun_out/*.ncu-rep, read here without a GPU) into profiles/*.json and
refreshes profiles/gemm_traffic.json (the per-launch DRAM traffic bench.py reports as roofline.traffic) with the
build id of the library the capture ran.  Usage: python tools/ncu_summary.py <round tag> <rep> [<rep> ...]"""
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "smsp__cycles_active.avg"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header = rows[0]
    return header, rows[2:] if len(rows) > 2 and not rows[1][0].isdigit() else rows[1:], rows[1] if len(rows) > 1 else None


def to_float(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return x


def main():
    tag = sys.argv[1]
    h = hashlib.sha256()  # identity of the kernel sources (same recipe as bench.py lib_build_id)
    for name in ("tob_kernels.cu", "tob_kernels.cuh", "tob_dispatch_table.h"):
        h.update(open(os.path.join(REPO, "tensororder_b200", "csrc", name), "rb").read())
    build = h.hexdigest()[:16]
    for rep in sys.argv[2:]:
        header, rows, units = raw_rows(rep)
        name = os.path.basename(rep)[:-len(".ncu-rep")]
        launches = []
        for r in rows:
            d = dict(zip(header, r))
            rec = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
            for k_ in header:
                if any(k_.startswith(w) for w in KEEP):
                    rec[k_] = to_float(d[k_])
            if units:
                rec["_units"] = {k_: u for k_, u in zip(header, units) if k_ in rec and u}
            launches.append(rec)
        m = re.match(r"gemm_(\d+)_(\d+)_(\d+)", name)
        doc = {"capture": name, "lib_build_id": build, "launches": launches}
        if m and launches:
            mm, nn, kk = (int(x) for x in m.groups())
            alg = 8.0 * (2.0 ** (mm + kk) + 2.0 ** (nn + kk) + 2.0 ** (mm + nn))
            L = launches[0]
            rd, wr = L.get("dram__bytes_read.sum"), L.get("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            ur = scale.get(L.get("_units", {}).get("dram__bytes_read.sum", "byte"), 1.0)
            uw = scale.get(L.get("_units", {}).get("dram__bytes_write.sum", "byte"), 1.0)
            doc.update({"join": {"m": mm, "n": nn, "k": kk}, "algorithmic_bytes": alg, "algorithmic_flops": 2.0 * 2.0 ** (mm + nn + kk),
                        "dram_bytes_per_launch": (rd * ur + wr * uw) if isinstance(rd, float) and isinstance(wr, float) else None})
            if name == "gemm_14_13_10":
                json.dump({k_: doc[k_] for k_ in ("join", "algorithmic_bytes", "dram_bytes_per_launch", "lib_build_id")} |
                          {"dram_bytes_read": rd * ur, "dram_bytes_write": wr * uw, "source": "profiles/%s_%s_ncu_summary.json" % (tag, name)},
                          open(os.path.join(REPO, "profiles", "gemm_traffic.json"), "w"), indent=1)
        json.dump(doc, open(os.path.join(REPO, "profiles", "%s_%s_ncu_summary.json" % (tag, name)), "w"), indent=1)
        print(name, "->", {k_: v for k_, v in (launches[0] if launches else {}).items() if k_ in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")})


if __name__ == "__main__":
    main()
