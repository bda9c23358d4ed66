#!/usr/bin/env python3
"""Benchmark of the contraction-executor hot path (BASELINE.json metric: contraction seconds per
instance, planning excluded; % of FP64/HBM roofline for the dominant kernel).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* = one pass over the workload: every instance of the cubic vertex-cover family
(n = 50..220 step 10; n=50 is the instance the reference ships, the others are regenerated with the
same recipe) contracted along its stored line-Flow plan.  The upper end is what
the reference's CPU path can also time within minutes, so both arms run the IDENTICAL workload; the
larger members (n = 230..250) are timed separately (`large_instances`) and `--workload sliced250` is
BASELINE config 4.  Instances with
n >= 200 use the reference slicer's `minimum_slice=3` plan (8 slices) at EVERY N so the same work is
compared at 1/2/4/8 GPUs: rank r contracts slices r, r+N, ... ; unsliced instances are independent
objects and are spread over the ranks (longest first); one NCCL all-reduce combines the count vector.

  value : inputs resident in HBM (plans compiled, leaves uploaded) when the timed region starts;
          timed with CUDA events on the stream the kernels run on, one event pair per step, an L2
          flush between steps, MAX over ranks per step.  Per-GEMM event timing is ON in this arm (it feeds
          `roofline`): GEMMs of the two slice lanes are then chained, which costs ~3 % on the sliced
          mid-size instances compared with the default (untimed) path the e2e arm runs.
  e2e   : the same workload through the reference-facing call `B200API.contract_sliced(plan)` with
          HOST leaf buffers: flatten + plan compile + arena allocation + pinned H2D of the leaves +
          kernels + D2H of the count inside the timed region, every step.
  --impl reference : the reference's CPU path (oracle port: the same numpy.tensordot calls the
          reference's numpy backend makes) on all host cores, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
METRIC = "contraction s/instance (plan excluded)"
UNIT = "s/instance"
FP64_PEAK_TFLOPS = 35.49  # cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/r01_fp64_calibration.txt);
# MEASURED_PEAKS.json has no FP64 entry (bf16 + HBM only).  Raw DMMA pipe peak measured 37.1.


def workload_config(args):
    if args.workload == "sliced250":
        return {
            "workload": "cubic_vc n=250 (random 3-regular vertex cover, seed 0), stored line-Flow plan sliced by the "
                        "reference slicer with minimum_slice=%d (%d slices), unweighted float64 (BASELINE config 4)"
                        % (args.slice_bits, 2 ** args.slice_bits),
            "planner": "line-Flow",
            "l2": "256 MiB write between steps; dominant operands exceed the 126 MB L2",
            "parallelism": "slices r::N, one all-reduce of the count",
        }
    return {
        "workload": "cubic_vc family n=%d..%d step 10 (random 3-regular vertex cover, seed 0; n=50 is the shipped "
                    "benchmarks/cubic_vertex_cover/cubic_vc_50_0.cnf), stored line-Flow plans, unweighted float64; "
                    "n>=200 as the reference slicer's minimum_slice=3 plans (8 slices)" % (args.min_n, args.max_n),
        "planner": "line-Flow",
        "l2": "256 MiB write between steps; dominant operands exceed the 126 MB L2",
        "parallelism": "slices r::N of sliced instances, unsliced instances spread over ranks, one all-reduce of the counts",
    }


def load_workload(min_n, max_n, workload="family", slice_bits=3):
    from tensororder_b200.plan_format import PortablePlan

    items = []
    if workload == "sliced250":
        pp = PortablePlan.load(os.path.join(GOLDEN, "vc250_lineflow.json.gz"))
        var = pp.variant("min%d" % slice_bits)
        return [{"n": 250, "name": var.name, "pp": var, "expected": var.expected.get("count", pp.expected.get("count")),
                 "cost": pp.expected.get("estimated_flops", 0.0)}]
    for n in range(min_n, max_n + 1, 10):
        pp = PortablePlan.load(os.path.join(GOLDEN, "vc%d_lineflow.json.gz" % n))
        unsliced_expected = pp.expected
        if n >= 200:
            pp = pp.variant("min3")
        items.append({"n": n, "name": pp.name, "pp": pp, "expected": unsliced_expected.get("count"),
                      "cost": unsliced_expected.get("estimated_flops", 0.0)})
    return items


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the samples taken under load (upper half: idle gaps between steps pull the clock down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": (max(smax) if smax else None),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        import threadpoolctl

        infos = [i for i in threadpoolctl.threadpool_info() if i.get("user_api") == "blas"]
        if infos:
            return int(infos[0]["num_threads"])
    except Exception:
        pass
    return os.cpu_count() or 1


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin OpenBLAS to one thread; the
    reference's numpy path gets every core the box has (OpenBLAS caps at its build maximum)."""
    import numpy  # noqa: F401

    try:
        import threadpoolctl

        threadpoolctl.threadpool_limits(limits=os.cpu_count() or 1, user_api="blas")
    except Exception:
        pass


def run_cpu_sample(items):
    """The reference's numpy path (oracle port) over `items`; returns seconds and counts."""
    from oracle import numpy_oracle

    use_all_cores()

    t0 = time.perf_counter()
    counts = [float(numpy_oracle.contract_sliced(it["pp"].to_json())) for it in items]
    return time.perf_counter() - t0, counts


def reference_arm(args, rank):
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    use_all_cores()

    items = [it for it in load_workload(args.min_n, args.max_n) if it["n"] <= args.cpu_max_n]
    n_workload = len(items)
    # keep the whole K+W run within a few minutes: drop the largest instances if one pass is too slow
    t_first, _ = run_cpu_sample(items)
    budget = 240.0
    while len(items) > 1 and t_first * (args.steps + args.warmup) > budget:
        items = items[:-1]
        t_first, _ = run_cpu_sample(items)
    for _ in range(max(args.warmup - 1, 0)):
        run_cpu_sample(items)
    total = 0.0
    counts = None
    for _ in range(args.steps):
        dt, counts = run_cpu_sample(items)
        total += dt
    ok = all(it["expected"] is None or abs(c - it["expected"]) <= 1e-9 * abs(it["expected"]) for it, c in zip(items, counts))
    per_step = total / args.steps
    value = per_step / len(items)
    sample = "instances n=%d..%d of the workload (%d of %d), one pass per step" % (
        items[0]["n"], items[-1]["n"], len(items), n_workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "counts_ok": bool(ok),
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def b200_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    from tensororder_b200.api import B200API, CompiledPlan
    from tensororder_b200.flatten import flatten_plan

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation (NCCL_DEBUG=VERSION/INFO): keep
        # stdout for the one JSON line by pointing fd 1 at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    items = load_workload(args.min_n, args.max_n, args.workload, args.slice_bits)
    n_inst = len(items)

    # ---- partition: sliced instances are shared by all ranks, unsliced ones go to one rank (LPT) ----
    load = [0.0] * world
    for it in sorted(items, key=lambda x: -x["cost"]):
        nsl = 2 ** len([g for g in it["pp"].groups_to_slice if len(g)])
        if nsl >= world and world > 1 and nsl > 1:
            it["owner"] = None
            for r in range(world):
                load[r] += it["cost"] / world
        else:
            r = min(range(world), key=lambda q: load[q])
            it["owner"] = r
            load[r] += it["cost"]
    mine = [it for it in items if it["owner"] in (None, rank)]

    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def reduce_counts(vec):
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        return vec

    # =============================== value: inputs resident in HBM ===============================
    for it in mine:
        cp = CompiledPlan(flatten_plan(it["pp"].as_execution_plan()), device=local_rank)
        cp.upload()
        cp.set_stream(stream.cuda_stream)
        cp.set_gemm_timing(True)  # CUDA-event pair around every DMMA GEMM of the timed steps (roofline)
        it["cp"] = cp
    counts = torch.zeros(n_inst, dtype=torch.float64, device=dev)
    step_ms, launches = [], 0
    gemm_ms = gemm_flops = 0.0
    gemm_launches = 0
    sampler = ClockSampler(local_rank)
    for step in range(args.warmup + args.steps):
        timed = step >= args.warmup
        if timed and step == args.warmup and rank == 0:
            sampler.start()
        with torch.cuda.stream(stream):
            flush.zero_()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        host = np.zeros(n_inst)
        if timed:
            torch.cuda.nvtx.range_push("timed_step")  # ncu --nvtx --nvtx-include "timed_step/" selects these launches
        with torch.cuda.stream(stream):
            e0.record(stream)
            for j, it in enumerate(items):
                if it["owner"] is None:
                    host[j] = it["cp"].run(first=rank, stride=world)
                elif it["owner"] == rank:
                    host[j] = it["cp"].run()
                else:
                    continue
                if timed:
                    launches += it["cp"].last_launches
                    g = it["cp"].last_gemm
                    gemm_ms += g[0]
                    gemm_flops += g[1]
                    gemm_launches += g[2]
            counts.copy_(torch.from_numpy(host))
            reduce_counts(counts)
            e1.record(stream)
        stream.synchronize()
        if timed:
            torch.cuda.nvtx.range_pop()
        barrier()
        if timed:
            step_ms.append(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor(step_ms, dtype=torch.float64, device=dev)
    lt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms_per_step = float(t.sum().item()) / args.steps
    value_counts = counts.cpu().numpy().copy()
    for it in mine:
        it["cp"].close()
        del it["cp"]

    # =============================== e2e: host buffers through the API ===========================
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    plans = {it["name"]: it["pp"].as_execution_plan() for it in mine}
    h2d = d2h = 0
    e2e_total = 0.0
    e2e_counts = None
    e2e_parts = {"flatten_compile_s": 0.0, "upload_s": 0.0, "run_s": 0.0, "device_ms": 0.0}
    e2e_step_s = []
    for step in range(1 + e2e_steps):  # one warm-up pass
        barrier()
        t0 = time.perf_counter()
        host = np.zeros(n_inst)
        h2d = d2h = 0
        for j, it in enumerate(items):
            if it["owner"] not in (None, rank):
                continue
            api = B200API()
            api.add_argument("entry_type", "float64")
            api.add_argument("device", local_rank)
            api.add_argument("distributed", it["owner"] is None)
            got = float(api.contract_sliced(plans[it["name"]]))
            # sliced instances come back already all-reduced (identical on every rank): count them once
            host[j] = got / world if it["owner"] is None else got
            h2d += api.last_stats["h2d_bytes"]
            d2h += api.last_stats["d2h_bytes"]
            if step >= 1:
                for key in e2e_parts:
                    e2e_parts[key] += api.last_stats[key]
        vec = torch.from_numpy(host).to(dev)
        reduce_counts(vec)
        e2e_counts = vec.cpu().numpy()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if step >= 1:
            e2e_total += float(dt.item())
            e2e_step_s.append(float(dt.item()))
    bytes_t = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(bytes_t, op=dist.ReduceOp.SUM)
    e2e_value = e2e_total / e2e_steps / n_inst

    # =============================== checks, CPU baseline, report ================================
    def counts_ok(vec):
        ok = True
        for it, c in zip(items, vec):
            if it["expected"] is not None:
                ok &= abs(c - it["expected"]) <= 1e-9 * abs(it["expected"])
            else:
                ok &= bool(np.isfinite(c) and c > 0)
        return bool(ok)

    ok = counts_ok(value_counts) and counts_ok(e2e_counts) and bool(
        np.allclose(value_counts, e2e_counts, rtol=1e-12, atol=0))

    if rank == 0:
        line = {
            "metric": METRIC, "value": ms_per_step / 1e3 / n_inst, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(bytes_t[0].item()),
                    "d2h_bytes_per_step": int(bytes_t[1].item()), "steps": e2e_steps,
                    "step_seconds": e2e_step_s,
                    "rank0_per_step": {k: v / e2e_steps for k, v in e2e_parts.items()}},
            "gpu_launches": int(lt.item()), "clocks": clocks, "counts_ok": ok, "instances": n_inst,
        }
        if gemm_launches > 0:
            achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
            traffic = None
            tpath = os.path.join(REPO, "profiles", "gemm_traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            line["roofline"] = {
                "bound": "tensor", "kernel": "k_gemm_dmma (DMMA.8x8x4 FP64)", "achieved": achieved,
                "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "traffic": traffic,
                "launches": int(gemm_launches), "avg_launch_ms": gemm_ms / gemm_launches,
                "flops_per_launch": gemm_flops / gemm_launches,
                "peak_source": "measured cuBLAS DGEMM 8192^3 on this pool (profiles/r01_fp64_calibration.txt); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "share_of_step": gemm_ms / (ms_per_step * args.steps),
            }
        if world == 1 and args.workload == "family" and not args.no_large:
            # beyond what the CPU arm can time: the largest family members, one pass each (device time)
            from tensororder_b200.plan_format import PortablePlan

            large = []
            for n in (230, 240, 250):
                pp = PortablePlan.load(os.path.join(GOLDEN, "vc%d_lineflow.json.gz" % n))
                cp = CompiledPlan(flatten_plan(pp.as_execution_plan()), device=local_rank)
                cp.upload()
                cp.set_gemm_timing(True)
                cp.run()
                c = cp.run()
                g = cp.last_gemm
                large.append({"n": n, "seconds": cp.last_ms / 1e3, "count": c, "peak_gb": cp.peak_bytes / 1e9,
                              "gemm_tflops": (g[1] / (g[0] * 1e-3) / 1e12) if g[0] > 0 else None,
                              "gemm_share": g[0] / cp.last_ms if cp.last_ms > 0 else None,
                              "reference_count": pp.expected.get("count")})
                cp.close()
            line["large_instances"] = large
        if world == 1 and args.workload == "family" and not args.no_cpu_baseline:
            sample_items = [it for it in items if it["n"] <= args.cpu_max_n]
            # our own arm on exactly the same sample (apples to apples next to the CPU number); before the CPU
            # pass, whose BLAS threads keep spinning for a while afterwards
            api_s = 0.0
            for it in sample_items:
                api = B200API()
                api.add_argument("entry_type", "float64")
                t0 = time.perf_counter()
                api.contract_sliced(plans[it["name"]])
                api_s += time.perf_counter() - t0
            cpu_s, cpu_counts = run_cpu_sample(sample_items)
            cpu_ok = all(it["expected"] is None or abs(c - it["expected"]) <= 1e-9 * abs(it["expected"])
                         for it, c in zip(sample_items, cpu_counts))
            line["cpu_baseline"] = {
                "value": cpu_s / len(sample_items), "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                "sample": "instances n=%d..%d of the workload (%d of %d), one pass" % (
                    sample_items[0]["n"], sample_items[-1]["n"], len(sample_items), n_inst),
                "counts_ok": bool(cpu_ok), "b200_e2e_same_sample": api_s / len(sample_items),
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--min-n", type=int, default=50)
    ap.add_argument("--max-n", type=int, default=220)
    ap.add_argument("--cpu-max-n", type=int, default=220, help="largest instance in the bounded CPU sample")
    ap.add_argument("--workload", default="family", choices=["family", "sliced250"])
    ap.add_argument("--slice-bits", type=int, default=3, choices=[3, 6])
    ap.add_argument("--no-large", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533")] + sys.argv
            sys.exit(subprocess.call(cmd))
    b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
