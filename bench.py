#!/usr/bin/env python3
"""Benchmark of the contraction-executor hot path (BASELINE.json metric: contraction seconds per
instance, planning excluded; % of FP64/HBM roofline for the dominant kernel).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* = one pass over the workload: every instance of the cubic vertex-cover family
(n = 50..220 step 10; n=50 is the instance the reference ships, the others are regenerated with the
same recipe) contracted along its stored line-Flow plan.  The upper end is what
the reference's CPU path can also time within minutes, so BOTH ARMS RUN THE IDENTICAL 18 INSTANCES; the
larger members (n = 230..250) are timed separately (`large_instances`).  Instances with n >= 200 use the
reference slicer's `minimum_slice=3` plan (8 slices) at EVERY N so the same work is compared at 1/2/4/8
GPUs: rank r contracts slices r, r+N, ... ; unsliced instances are independent objects and are spread over
the ranks (longest first, by the compiled programs' cost); one NCCL all-reduce combines the count vector.

  value : inputs resident in HBM (plans compiled, leaves uploaded) when the timed region starts.  All of
          a rank's instances are issued asynchronously (tob_plan_run_async: every plan on its own streams,
          ordered behind the timing stream), so the launch-bound stretches of one contraction overlap the
          GEMMs of another; one CUDA-event pair per step on the timing stream, an L2 flush between steps,
          MAX over ranks per step.  Per-GEMM event timing is OFF in this arm (it is the default code path).
  roofline : a separate sequential pass after the timed steps with a CUDA-event pair around every DMMA GEMM
          (algorithmic flops / summed durations); the FP64 peak is MEASURED IN THIS RUN (cuBLAS DGEMM 8192^3
          through torch.matmul) next to the round-1 calibration and the raw DMMA pipe peak.
  e2e   : the same workload through the reference-facing call `B200API.contract_sliced(plan)` with HOST
          leaf buffers: leaf build (`Tensor.build`) + pinned H2D of the leaves + kernels + D2H of the count
          inside the timed region, every step; the plan compile is paid once per plan (plan cache keyed
          by plan identity) in the first of two untimed warm-up passes.  A pool of host threads
          (8 at N = 1) makes the calls, so small contractions overlap the large ones' GEMMs; at N > 1 every sliced
          instance has a thread and a collective ticket of its own (api.CollectiveOrder: the contractions overlap,
          their all-reduces are issued in the same order on every rank).
  extra : BASELINE configs 3, 4 and 5 measured in the same run: `weighted150` (n=150 mcc weights,
          factor-Flow), `sliced250` (n=250, 8 slices, at this N), `rank_sweep` (N=1; single timed launches and
          the same launches back to back).
  rank_split : per rank, where the device time goes (replicated prologues / own slices / owned unsliced
          instances / the count all-reduce).
  --impl reference : the REAL reference (oracle/_ref: vardigroup/TensorOrder's own
          `NumpyAPI.contract_sliced` on reference objects rebuilt from the same stored plans) on all host
          cores, over the same 18 instances.
"""
import argparse
import ctypes
import hashlib
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
FP64_CUBLAS_CALIBRATION = 35.49  # cuBLAS DGEMM 8192^3 on this pool's B200s, round 1 (profiles/r01_fp64_calibration.txt)
FP64_DMMA_PIPE = 37.1            # raw DMMA.8x8x4 issue peak measured by tools/fp64_peak.cu (same file)
E2E_THREADS = int(os.environ.get("TOB_BENCH_E2E_THREADS", "8"))  # host threads of the e2e arm at N = 1 (sliced instances one each, the small ones share the rest)


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return {}


def workload_config(args):
    return {
        "workload": "cubic_vc family n=%d..%d step 10 (random 3-regular vertex cover, seed 0; n=50 is the shipped "
                    "benchmarks/cubic_vertex_cover/cubic_vc_50_0.cnf), stored line-Flow plans, unweighted float64; "
                    "n>=200 as the reference slicer's minimum_slice=3 plans (8 slices); %d instances, identical in "
                    "both arms" % (args.min_n, args.max_n, len(range(args.min_n, args.max_n + 1, 10))),
        "planner": "line-Flow",
        "l2": "256 MiB write between steps; dominant operands exceed the 126 MB L2",
        "parallelism": "slices r::N of sliced instances, unsliced instances spread over ranks, one all-reduce of the counts",
    }


def load_workload(min_n, max_n):
    from tensororder_b200.plan_format import PortablePlan

    items = []
    for n in range(min_n, max_n + 1, 10):
        pp = PortablePlan.load(os.path.join(GOLDEN, "vc%d_lineflow.json.gz" % n))
        unsliced_expected = pp.expected
        if n >= 200:
            pp = pp.variant("min3")
        items.append({"n": n, "name": pp.name, "pp": pp, "expected": unsliced_expected.get("count", pp.expected.get("count")),
                      "cost": unsliced_expected.get("estimated_flops", 0.0)})
    return items


def rel_ok(got, want, tol=1e-9):
    return want is not None and abs(got - want) <= tol * abs(want)


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
        self.nvml = None
        self.samples = []   # (sm MHz, max sm MHz, throttle-reason bitmask, watts) through NVML, every 20 ms
        self.stop_flag = False

    def start(self):
        try:  # NVML in-process: ~50 samples per second instead of nvidia-smi's 5
            import pynvml

            pynvml.nvmlInit()
            cuda_visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(cuda_visible.split(",")[self.gpu]) if cuda_visible and cuda_visible.split(",")[self.gpu].isdigit() else self.gpu
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(index))
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml
        smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), smax,
                                     nv.nvmlDeviceGetCurrentClocksThrottleReasons(h), nv.nvmlDeviceGetPowerUsage(h) / 1e3))
            except Exception:
                pass
            time.sleep(0.02)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            nv = self.nvml[0]
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
            sm = sorted(x[0] for x in self.samples)
            load = sm[len(sm) // 2:] if sm else []  # upper half: idle gaps between steps pull the clock down
            reasons = sorted({name for x in self.samples for name, bit in names if x[2] & bit})
            return {"sm_mhz": (float(load[len(load) // 2]) if load else None), "sm_min_mhz": (float(sm[0]) if sm else None),
                    "sm_max_mhz": (float(max(x[1] for x in self.samples)) if self.samples else None), "reasons": reasons,
                    "samples": len(sm), "power_w_max": (max(x[3] for x in self.samples) if self.samples else None), "source": "nvml, 20 ms"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi, 200 ms"}


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


class CpuArm:
    """The reference's CPU implementation of the path over the workload's instances.  kind "reference": the
    real thing (oracle/_ref — `NumpyAPI.contract_sliced`, src/tensor_network/tensor_apis/base_api.py:17-28 +
    numpy_apis.py:42-55, driving `TensorNetwork.slice_groups` and `identify` on reference objects rebuilt from
    the stored plans); kind "port": the numpy restatement (oracle/numpy_oracle.py), only when oracle/_ref was
    never built."""

    def __init__(self, items):
        from oracle import reference

        use_all_cores()
        self.items = items
        self.kind = "reference" if reference.available() else "port"
        if self.kind == "reference":
            R = reference.import_reference()
            self.api = R["tensor_network"].ALL_APIS["numpy"]()
            self.api.add_argument("entry_type", "float64")
            self.plans = [reference.to_reference_plan(R, it["pp"]) for it in items]
        else:
            self.docs = [it["pp"].to_json() for it in items]

    def one_pass(self):
        t0 = time.perf_counter()
        if self.kind == "reference":
            counts = [float(self.api.contract_sliced(p)) for p in self.plans]
        else:
            from oracle import numpy_oracle

            counts = [float(numpy_oracle.contract_sliced(d)) for d in self.docs]
        return time.perf_counter() - t0, counts

    def counts_ok(self, counts):
        return all(it["expected"] is None or rel_ok(c, it["expected"]) for it, c in zip(self.items, counts))


def reference_arm(args, rank):
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    items = load_workload(args.min_n, args.max_n)
    arm = CpuArm(items)
    # the instance list is NEVER shortened (both arms run the same instances); if the box is so slow that
    # W + K passes would take more than ~25 minutes, fewer timed passes are taken and the line says so
    t_first, counts = arm.one_pass()
    steps, warmup = args.steps, args.warmup
    truncated = False
    budget = 1500.0
    if t_first * (steps + warmup) > budget:
        truncated = True
        warmup = 1
        steps = max(1, min(steps, int(budget / t_first) - 1))
    for _ in range(max(warmup - 1, 0)):
        arm.one_pass()
    total = 0.0
    for _ in range(steps):
        dt, counts = arm.one_pass()
        total += dt
    per_step = total / steps
    value = per_step / len(items)
    sample = "instances n=%d..%d of the workload (%d of %d), one pass per step" % (
        items[0]["n"], items[-1]["n"], len(items), len(items))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": arm.kind, "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "counts_ok": bool(arm.counts_ok(counts)), "instances": len(items), "truncated": truncated,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def program_cost_s(desc, slices_per_rank):
    """Modelled device seconds of one compiled program on one rank: per op max(flops / 33 TF, bytes / 5 TB/s)
    plus a launch latency; the slice-invariant prologue once, the per-slice list once per slice this rank runs."""
    def ops_cost(ops):
        t = 0.0
        for op in ops:
            t += max(op.get("flops", 0.0) / 33e12, op.get("bytes", 0.0) / 5e12) + 3e-6
        return t
    return ops_cost(desc["invariant_ops"]) + slices_per_rank * ops_cost(desc["slice_ops"])


def measure_fp64_peak(torch, dev, n=8192, reps=6):
    """cuBLAS DGEMM n^3 through torch.matmul, best of `reps` (CUDA events): the FP64 roofline denominator measured in
    this run (MEASURED_PEAKS.json carries only bf16 and HBM figures)."""
    a = torch.rand(n, n, dtype=torch.float64, device=dev)
    b = torch.rand(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize(dev)
    best = None
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def lib_build_id():
    """Identity of the kernel sources the library was built from (tools/ncu_summary.py stamps its captures with it)."""
    h = hashlib.sha256()
    for name in ("tob_kernels.cu", "tob_kernels.cuh", "tob_dispatch_table.h"):
        with open(os.path.join(REPO, "tensororder_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def gemm_traffic():
    """DRAM bytes of one dominant GEMM launch from an `ncu --set full` capture of THIS build
    (tools/ncu_gemm_traffic.sh writes profiles/gemm_traffic.json with the library's build id)."""
    path = os.path.join(REPO, "profiles", "gemm_traffic.json")
    if not os.path.exists(path):
        return None, None
    doc = json.load(open(path))
    return doc.get("dram_bytes_per_launch"), {
        "source": "profiles/gemm_traffic.json (ncu --set full, one launch of the join it names)",
        "join": doc.get("join"), "algorithmic_bytes": doc.get("algorithmic_bytes"),
        "captured_build": doc.get("lib_build_id"), "this_build": lib_build_id(),
        "same_build": doc.get("lib_build_id") == lib_build_id()}


# --------------------------------------------------------------------------------------------------
def extra_weighted150(torch, dev, local_rank, hbm_gbs, fp64_peak):
    """BASELINE config 3: n=150 with random literal weights (mcc), factor-Flow tree, float64, one GPU."""
    from tensororder_b200.api import B200API, CompiledPlan
    from tensororder_b200.flatten import flatten_plan
    from tensororder_b200.plan_format import PortablePlan

    pp = PortablePlan.load(os.path.join(GOLDEN, "vc150_mcc_factorflow.json.gz"))
    want = pp.expected["count"]
    cp = CompiledPlan(flatten_plan(pp.as_execution_plan()), device=local_rank)
    cp.upload()
    for _ in range(3):
        got = cp.run()
    ms = []
    for _ in range(10):
        got = cp.run()
        ms.append(cp.last_ms)
    desc = cp.describe()
    ops = desc["invariant_ops"] + desc["slice_ops"]
    cp.profile(0)
    per_op, _ = cp.profile(0)
    nodes = []
    for t, op in sorted(zip(per_op, ops), key=lambda x: -x[0])[:4]:
        if op["kind"] not in (0, 1) or t <= 0:
            continue
        tf = op["flops"] / (t * 1e-3) / 1e12
        gb = op["bytes"] / (t * 1e-3) / 1e9
        bound = "tensor" if op["flops"] / (fp64_peak * 1e12) > op["bytes"] / (hbm_gbs * 1e9) else "hbm"
        nodes.append({"m": op["m"], "n": op["n"], "k": op["k"], "kernel": "gemm" if op["kind"] == 1 else "generic",
                      "ksplit_log2": op["ksplit_log2"], "ms": t, "tflops": tf, "gbs": gb, "bound": bound,
                      "frac": tf / fp64_peak if bound == "tensor" else gb / hbm_gbs})
    cp.close()
    api = B200API()
    api.add_argument("entry_type", "float64")
    api.add_argument("device", local_rank)
    api.add_argument("distributed", False)
    plan = pp.as_execution_plan()
    api.contract_sliced(plan)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        e2e_got = float(api.contract_sliced(plan))
    e2e_s = (time.perf_counter() - t0) / reps
    ms.sort()
    return {"config": "BASELINE config 3: cubic_vc n=150 + mcc literal weights U(0.5,1.5), factor-Flow, float64, 1 GPU",
            "device_seconds": ms[len(ms) // 2] / 1e3, "e2e_seconds": e2e_s, "count": got, "reference_count": want,
            "rel_err": abs(got - want) / abs(want), "count_ok": bool(rel_ok(got, want) and rel_ok(e2e_got, want)),
            "tolerance": 1e-9, "dominant_nodes": nodes}


def extra_sliced250(torch, dist, dev, local_rank, rank, world, passes=2):
    """BASELINE config 4: n=250 sliced by the reference slicer (minimum_slice=3 -> 8 slices), slices r::N,
    one all-reduce of the count; the SAME plan at every N."""
    from tensororder_b200.api import CompiledPlan
    from tensororder_b200.flatten import flatten_plan
    from tensororder_b200.plan_format import PortablePlan

    pp = PortablePlan.load(os.path.join(GOLDEN, "vc250_lineflow.json.gz")).variant("min3")
    want = pp.expected.get("count")
    cp = CompiledPlan(flatten_plan(pp.as_execution_plan()), device=local_rank)
    cp.upload()
    stream = torch.cuda.Stream(device=dev)
    cp.set_stream(stream.cuda_stream)
    secs, got = [], None
    for i in range(1 + passes):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            part = cp.run(first=rank, stride=world)
            vec = torch.tensor([part], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(vec)
            e1.record(stream)
        stream.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= 1:
            secs.append(float(t.item()) / 1e3)
        got = float(vec.item())
    out = {"config": "BASELINE config 4: cubic_vc n=250, stored line-Flow plan, reference slicer minimum_slice=3 "
                     "(8 slices), slices r::N + one NCCL all-reduce of the count",
           "n_gpus": world, "slices": cp.num_slices, "seconds": min(secs), "seconds_all": secs, "count": got,
           "reference_count": want, "count_ok": bool(rel_ok(got, want)), "peak_gb_per_gpu": cp.peak_bytes / 1e9,
           "launches_per_pass_rank0": cp.last_launches}
    cp.close()
    return out


SWEEP_POINTS = [(28, 8), (28, 12), (28, 16), (30, 8), (30, 16), (32, 2), (32, 4), (32, 8), (32, 16), (34, 2), (34, 4), (34, 12)]


def extra_rank_sweep(torch, dev, hbm_gbs, fp64_peak):
    """BASELINE config 5 digest: single pairwise contractions, T = fL + fR + k total indices, k contracted,
    GEMM-ready operands (`tob_tensordot_device`); the full sweep is tools/rank_sweep.py."""
    import numpy as np

    from tensororder_b200 import cabi

    P32 = ctypes.POINTER(ctypes.c_int32)
    rows = []
    for T, k in SWEEP_POINTS:
        fL = (T - k + 1) // 2
        fR = T - k - fL
        ra, rb, rc = fL + k, fR + k, fL + fR
        a = torch.rand(1 << ra, dtype=torch.float64, device=dev)
        b = torch.rand(1 << rb, dtype=torch.float64, device=dev)
        c = torch.empty(1 << rc, dtype=torch.float64, device=dev)
        ws_bytes = 8 * min(1 << (rc + 8), 1 << 28) + 4096  # split-K partials up to 2^8 splits
        ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device=dev)
        aa = np.arange(ra - k, ra, dtype=np.int32)
        ab = np.arange(rb - k, rb, dtype=np.int32)
        best = None
        # a freshly allocated multi-GB output needs ~6 full passes before its writes run at speed (measured: m=n=15,k=4 takes
        # 2.16 ms for its first five launches, 1.47 ms from then on and on any re-used buffer; profiles/r02i_first_touch.md)
        for rep in range(10 if T >= 32 else 4):
            ms = (ctypes.c_float * 3)()
            torch.cuda.synchronize(dev)
            rc_ = cabi.lib.tob_tensordot_device(a.data_ptr(), ra, b.data_ptr(), rb, aa.ctypes.data_as(P32),
                                                ab.ctypes.data_as(P32), k, c.data_ptr(), ws.data_ptr(), ws_bytes, 0, None, ms)
            if rc_ != 0:
                raise RuntimeError(cabi.last_error())
            if rep and (best is None or ms[1] < best):
                best = ms[1]
        # the same launch back to back (R launches captured into one CUDA graph): a join inside a tree follows its
        # predecessor on the stream, without the ~5 us a lone event-bracketed launch carries
        reps = 20 if T <= 30 else 4
        side = torch.cuda.Stream(device=dev)
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                rc_ = cabi.lib.tob_tensordot_device(a.data_ptr(), ra, b.data_ptr(), rb, aa.ctypes.data_as(P32),
                                                    ab.ctypes.data_as(P32), k, c.data_ptr(), ws.data_ptr(), ws_bytes, 0,
                                                    ctypes.c_void_p(side.cuda_stream), None)
                if rc_ != 0:
                    raise RuntimeError(cabi.last_error())
        b2b = None
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            t = e0.elapsed_time(e1) / reps
            b2b = t if b2b is None or t < b2b else b2b
        del graph
        flops = 2.0 * 2.0 ** T
        byts = 8.0 * (2.0 ** ra + 2.0 ** rb + 2.0 ** rc)
        tf = flops / (best * 1e-3) / 1e12
        gb = byts / (best * 1e-3) / 1e9
        bound = "tensor" if flops / (fp64_peak * 1e12) > byts / (hbm_gbs * 1e9) else "hbm"
        work, peak = (flops / 1e12, fp64_peak) if bound == "tensor" else (byts / 1e9, hbm_gbs)
        rows.append({"T": T, "k": k, "ms": best, "tflops": tf, "gbs": gb, "bound": bound,
                     "frac": tf / fp64_peak if bound == "tensor" else gb / hbm_gbs,
                     "ms_back_to_back": b2b, "frac_back_to_back": work / (b2b * 1e-3) / peak})
        del a, b, c, ws
        torch.cuda.empty_cache()
    return {"config": "BASELINE config 5 digest: single contractions, T total / k contracted indices, GEMM-ready operands",
            "timing": "ms / frac: ONE launch between two CUDA events (best of 3, of 9 for T >= 32); ms_back_to_back / frac_back_to_back: the same "
                      "launch repeated inside one CUDA graph, replay time / repetitions (operands of T <= 30 stay in the 126 MB L2 "
                      "either way, as they do behind their producer in a tree)",
            "points": rows}


# --------------------------------------------------------------------------------------------------
def b200_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    from tensororder_b200.api import PLAN_CACHE, B200API, CollectiveOrder, CompiledPlan
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
    peaks = measured_peaks()
    hbm_gbs = float(peaks.get("hbm_gbs", 6552.6))
    items = load_workload(args.min_n, args.max_n)
    n_inst = len(items)

    # ---- partition: sliced instances are shared by all ranks, unsliced ones go to one rank (longest first), by the
    # modelled cost of the COMPILED programs: a shared instance costs every rank its replicated slice-invariant
    # prologue plus its share of the slices ----
    load = [0.0] * world
    for it in items:
        cp = CompiledPlan(flatten_plan(it["pp"].as_execution_plan()), device=local_rank)  # host only
        it["nsl"] = cp.num_slices
        it["shared"] = world > 1 and it["nsl"] > 1 and it["nsl"] >= world
        it["model_s"] = program_cost_s(cp.describe(), (it["nsl"] + world - 1) // world if it["shared"] else it["nsl"])
        cp.close()
    for it in items:
        if it["shared"]:
            it["owner"] = None
            for r in range(world):
                load[r] += it["model_s"]
    for it in sorted([x for x in items if not x["shared"]], key=lambda x: -x["model_s"]):
        r = min(range(world), key=lambda q: load[q])
        it["owner"] = r
        load[r] += it["model_s"]
    mine = [it for it in items if it["owner"] in (None, rank)]
    mine.sort(key=lambda x: -x["model_s"])  # issue order: the long GEMM-heavy contractions first

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
        if args.sequential:
            cp.set_stream(stream.cuda_stream)
        it["cp"] = cp
    index = {it["name"]: j for j, it in enumerate(items)}
    counts = torch.zeros(n_inst, dtype=torch.float64, device=dev)
    step_ms, launches = [], 0
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
            if args.sequential:
                for it in mine:
                    host[index[it["name"]]] = it["cp"].run(first=rank, stride=world) if it["owner"] is None else it["cp"].run()
            else:
                for it in mine:  # all of this rank's contractions in flight at once, each on its own streams
                    if it["owner"] is None:
                        it["cp"].run_async(first=rank, stride=world, after_stream=stream.cuda_stream)
                    else:
                        it["cp"].run_async(after_stream=stream.cuda_stream)
                for it in mine:
                    it["cp"].join(stream.cuda_stream)
                for it in mine:
                    host[index[it["name"]]] = it["cp"].wait()
            counts.copy_(torch.from_numpy(host))
            reduce_counts(counts)
            e1.record(stream)
        stream.synchronize()
        if timed:
            torch.cuda.nvtx.range_pop()
            launches += sum(it["cp"].last_launches for it in mine)
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

    # ---- roofline pass: the same plans, sequentially, a CUDA-event pair around every DMMA GEMM ----
    gemm_ms = gemm_flops = seq_ms = 0.0
    gemm_launches = 0
    for it in mine:
        it["cp"].set_gemm_timing(True)
        if it["owner"] is None:
            it["cp"].run(first=rank, stride=world)
        else:
            it["cp"].run()
        g = it["cp"].last_gemm
        gemm_ms += g[0]
        gemm_flops += g[1]
        gemm_launches += g[2]
        seq_ms += it["cp"].last_ms
    # ---- where this rank's device time goes: the replicated slice-invariant prologues of the shared instances, its
    # slices, the unsliced instances it owns (per-op CUDA events, tob_plan_profile: sequential, every small op carries
    # its event pair) and the count all-reduce ----
    split = [0.0, 0.0, 0.0, 0.0]
    for it in mine:
        cp = it["cp"]
        cp.set_gemm_timing(False)
        ms, _ = cp.profile(rank if it["owner"] is None else 0)
        n_inv = len(cp.describe()["invariant_ops"])
        pro, per_slice = float(sum(ms[:n_inv])), float(sum(ms[n_inv:]))
        if it["owner"] is None:
            split[0] += pro
            split[1] += per_slice * ((it["nsl"] - rank + world - 1) // world)
        else:
            split[2] += pro + per_slice * it["nsl"]
    if world > 1:
        probe = torch.zeros(4, dtype=torch.float64, device=dev)
        for rep in range(12):
            if rep == 2:
                torch.cuda.synchronize(dev)
                a0 = torch.cuda.Event(enable_timing=True)
                a1 = torch.cuda.Event(enable_timing=True)
                a0.record()
            dist.all_reduce(probe)
        a1.record()
        torch.cuda.synchronize(dev)
        split[3] = a0.elapsed_time(a1) / 10.0
    split_t = torch.tensor(split, dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(split_t) for _ in range(world)]
        dist.all_gather(gathered, split_t)
    else:
        gathered = [split_t]
    rank_split = [{"rank": r, "shared_prologues_ms": float(g_[0]), "own_slices_ms": float(g_[1]), "unsliced_ms": float(g_[2]),
                   "count_allreduce_ms": float(g_[3])} for r, g_ in enumerate(gathered)]
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
    cache_hits = 0
    from concurrent.futures import ThreadPoolExecutor

    order = CollectiveOrder()

    def contract_one(it, ticket=None):
        """One instance through the reference-facing call; returns (count, stats)."""
        api = B200API()
        api.add_argument("entry_type", "float64")
        api.add_argument("device", local_rank)
        api.add_argument("distributed", it["owner"] is None)
        if ticket is not None:
            api.add_argument("collective_ticket", (order, ticket))
        got = float(api.contract_sliced(plans[it["name"]]))
        # sliced instances come back already all-reduced (identical on every rank): count them once
        return (got / world if it["owner"] is None else got), api.last_stats

    def contract_batch(batch, step):
        torch.cuda.set_device(local_rank)
        return [contract_one(it, step * len(ordered) + ordered.index(it) if it in ordered else None) for it in batch]

    # host threads making the public call: the instances of a step are independent objects, so a user contracts them from
    # a small thread pool (ctypes releases the GIL inside the C ABI; every plan runs on its own streams, so the GPU sees
    # several contractions at once and one thread's host work hides behind another's device time).  Instances whose call
    # ends in an all-reduce (N > 1: the sliced ones) get a thread each and a collective ticket (api.CollectiveOrder): they
    # run concurrently on the device and their all-reduces are issued in ticket order on every rank; everything else is
    # packed longest-first over the remaining threads (profiles/r02i_e2e_threads.md).
    ordered = [it for it in mine if world > 1 and it["owner"] is None]
    free = [it for it in mine if it not in ordered]
    n_free = max(1, E2E_THREADS - len(ordered))
    packs, pload = [[] for _ in range(n_free)], [0.0] * n_free
    for it in sorted(free, key=lambda x: -x["model_s"]):
        g = min(range(n_free), key=lambda q: pload[q])
        packs[g].append(it)
        pload[g] += it["model_s"]
    groups = [[it] for it in ordered] + packs
    groups = [g for g in groups if g] or [[]]
    pool = ThreadPoolExecutor(max_workers=max(1, len(groups) - 1))
    E2E_WARM = 2  # untimed passes: the first pays the plan compiles (cached by plan identity afterwards), the second the CUDA-graph
    #               capture of the launch-bound stretches (a resident plan is captured the second time it runs)
    for step in range(E2E_WARM + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        host = np.zeros(n_inst)
        h2d = d2h = 0
        side = [pool.submit(contract_batch, g, step) for g in groups[1:]]
        results = list(zip(groups[0], contract_batch(groups[0], step)))
        for g, fut in zip(groups[1:], side):
            results += list(zip(g, fut.result()))
        timeline = [{"n": it["n"], "thread": next(gi for gi, g in enumerate(groups) if it in g),
                     "ms_from_step_start": [round((x - t0) * 1e3, 2) for x in stats["t_call"]]}
                    for it, (got, stats) in results if it["model_s"] >= 1e-3]
        for it, (got, stats) in results:
            host[index[it["name"]]] = got
            h2d += stats["h2d_bytes"]
            d2h += stats["d2h_bytes"]
            if step >= E2E_WARM:
                cache_hits += int(stats["plan_cache_hit"])
                for key in e2e_parts:
                    e2e_parts[key] += stats[key]
        vec = torch.from_numpy(host).to(dev)
        reduce_counts(vec)
        e2e_counts = vec.cpu().numpy()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if step >= E2E_WARM:
            e2e_total += float(dt.item())
            e2e_step_s.append(float(dt.item()))
    pool.shutdown()
    bytes_t = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(bytes_t, op=dist.ReduceOp.SUM)
    e2e_value = e2e_total / e2e_steps / n_inst
    PLAN_CACHE.clear()

    # =============================== checks, extras, CPU baseline, report ========================
    def counts_ok(vec):
        ok = True
        for it, c in zip(items, vec):
            if it["expected"] is not None:
                ok &= rel_ok(c, it["expected"])
            else:
                ok &= bool(np.isfinite(c) and c > 0)
        return bool(ok)

    ok = counts_ok(value_counts) and counts_ok(e2e_counts) and bool(
        np.allclose(value_counts, e2e_counts, rtol=1e-12, atol=0))

    extra = {}
    if not args.no_extra:
        extra["sliced250"] = extra_sliced250(torch, dist, dev, local_rank, rank, world)
    if rank == 0:
        fp64_live = measure_fp64_peak(torch, dev)
        line = {
            "metric": METRIC, "value": ms_per_step / 1e3 / n_inst, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(bytes_t[0].item()),
                    "d2h_bytes_per_step": int(bytes_t[1].item()), "steps": e2e_steps,
                    "step_seconds": e2e_step_s, "plan_cache_hits_rank0": cache_hits, "host_threads": len(groups),
                    "timeline_last_step_rank0": {"note": "calls of the instances modelled >= 1 ms: entry, plan acquired (cache hit: leaves "
                                                         "re-read), leaves on the device, run done", "calls": timeline,
                                                 "step_ms": round(e2e_step_s[-1] * 1e3, 2)},
                    "rank0_per_step": {k: v / e2e_steps for k, v in e2e_parts.items()}},
            "gpu_launches": int(lt.item()), "clocks": clocks, "counts_ok": ok, "instances": n_inst,
            "issue": "sequential" if args.sequential else "async: all of a rank's instances in flight",
            "rank0": {"instances": [it["n"] for it in mine], "sequential_device_ms": seq_ms,
                      "modelled_load_s": load},
            "rank_split": {"note": "per rank, sequential per-op event times (tob_plan_profile): slice-invariant prologues of the "
                                   "instances every rank shares (replicated), this rank's slices of them, the unsliced instances "
                                   "it owns, one 32-byte count all-reduce (the only collective of a contraction)",
                           "ranks": rank_split},
        }
        if gemm_launches > 0:
            achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
            traffic, traffic_note = gemm_traffic()
            line["roofline"] = {
                "bound": "tensor", "kernel": "k_gemm_dmma* (DMMA.8x8x4 FP64)", "achieved": achieved,
                "peak": fp64_live, "unit": "TFLOP/s", "frac": achieved / fp64_live, "traffic": traffic,
                "launches": int(gemm_launches), "avg_launch_ms": gemm_ms / gemm_launches,
                "flops_per_launch": gemm_flops / gemm_launches,
                "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul, best of 6, CUDA events) measured in THIS run; "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "peak_round1_calibration": FP64_CUBLAS_CALIBRATION, "frac_of_round1_calibration": achieved / FP64_CUBLAS_CALIBRATION,
                "dmma_pipe_peak": FP64_DMMA_PIPE, "frac_of_dmma_pipe": achieved / FP64_DMMA_PIPE,
                "measured": "separate sequential pass over rank 0's instances after the timed steps, one CUDA-event pair per GEMM launch",
                "gemm_ms_rank0": gemm_ms, "share_of_step": gemm_ms / seq_ms if seq_ms > 0 else None,
                "share_note": "GEMM event time / device time of the same sequential pass on rank 0 (comparable with the "
                              "serialised ncu launch list); in the timed (asynchronous) steps everything else overlaps the GEMMs",
                "traffic_note": traffic_note,
            }
        if world == 1 and not args.no_large:
            # beyond what the CPU arm can time: the largest family members, one pass each (device time)
            from tensororder_b200.plan_format import PortablePlan

            large = []
            for n in (230, 240, 250):
                pp = PortablePlan.load(os.path.join(GOLDEN, "vc%d_lineflow.json.gz" % n))
                want = pp.expected.get("count")
                if want is None:  # the reference replayed the 8-slice plan of this instance (tests/golden/ref_replay.py)
                    want = pp.variant("min3").expected.get("count")
                cp = CompiledPlan(flatten_plan(pp.as_execution_plan()), device=local_rank)
                cp.upload()
                cp.set_gemm_timing(True)
                cp.run()
                c = cp.run()
                g = cp.last_gemm
                large.append({"n": n, "seconds": cp.last_ms / 1e3, "count": c, "peak_gb": cp.peak_bytes / 1e9,
                              "gemm_tflops": (g[1] / (g[0] * 1e-3) / 1e12) if g[0] > 0 else None,
                              "gemm_share": g[0] / cp.last_ms if cp.last_ms > 0 else None,
                              "reference_count": want, "count_ok": bool(rel_ok(c, want))})
                cp.close()
            line["large_instances"] = large
        if not args.no_extra:
            extra["weighted150"] = extra_weighted150(torch, dev, local_rank, hbm_gbs, fp64_live)
            if world == 1:
                extra["rank_sweep"] = extra_rank_sweep(torch, dev, hbm_gbs, fp64_live)
        line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            # our own arm on exactly the same sample (apples to apples next to the CPU number); before the CPU
            # pass, whose BLAS threads keep spinning for a while afterwards
            api_s = 0.0
            fresh = {it["name"]: it["pp"].as_execution_plan() for it in items}
            for it in items:
                api = B200API()
                api.add_argument("entry_type", "float64")
                t0 = time.perf_counter()
                api.contract_sliced(fresh[it["name"]])
                api_s += time.perf_counter() - t0
            PLAN_CACHE.clear()
            arm = CpuArm(items)
            cpu_s, cpu_counts = arm.one_pass()
            line["cpu_baseline"] = {
                "value": cpu_s / len(items), "unit": UNIT, "cores": cpu_threads(), "kind": arm.kind,
                "sample": "instances n=%d..%d of the workload (%d of %d), one pass" % (
                    items[0]["n"], items[-1]["n"], len(items), n_inst),
                "counts_ok": bool(arm.counts_ok(cpu_counts)), "host_cpus": os.cpu_count(),
                "b200_e2e_same_sample_cold": api_s / len(items),
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--min-n", type=int, default=50)
    ap.add_argument("--max-n", type=int, default=220)
    ap.add_argument("--no-large", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip BASELINE configs 3, 4, 5 (extra blocks)")
    ap.add_argument("--sequential", action="store_true", help="value arm: one instance after the other on one stream")
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
