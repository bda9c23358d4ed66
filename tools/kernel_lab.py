"""Kernel lab: single joins under dispatch variants, timed BACK TO BACK (R launches captured into one CUDA graph, replayed;
per-launch time = replay time / R — the way a join runs inside a tree, without the event-pair overhead a single timed launch
carries) and as a single launch between two events (what tools/rank_sweep.py and the bench digest report).
Every variant's result is held against the default kernel's.  Usage: python tools/kernel_lab.py [streamk|store|all]"""
import ctypes, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from tensororder_b200 import cabi

HBM = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6551.7
FP64 = 35.47
P32 = ctypes.POINTER(ctypes.c_int32)
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def tset(**kv):
    for key, value in kv.items():
        assert cabi.lib.tob_tuning_set(key.encode(), float(value)) == 0, cabi.last_error()


DEFAULTS = {}
for key in ("streamk", "streamk_min_tiles_log2", "streamk_max_tiles_log2", "streamk_max_steps", "store_group_log2", "store_tile", "ws_min_k", "t256_ctas_log2", "force_ksplit_log2", "persist_max_k"):
    v = ctypes.c_double()
    assert cabi.lib.tob_tuning_get(key.encode(), ctypes.byref(v)) == 0
    DEFAULTS[key] = v.value


def run_join(m, n, k, a, b, c, ws, ws_bytes, stream_ptr, ms=None):
    aa = np.arange(m, m + k, dtype=np.int32)
    ab = np.arange(n, n + k, dtype=np.int32)
    rc = cabi.lib.tob_tensordot_device(a.data_ptr(), m + k, b.data_ptr(), n + k, aa.ctypes.data_as(P32), ab.ctypes.data_as(P32), k,
                                       c.data_ptr(), ws.data_ptr(), ws_bytes, 0, stream_ptr, ms)
    assert rc == 0, cabi.last_error()


def time_join(m, n, k, knobs, reps):
    tset(**DEFAULTS)
    tset(**knobs)
    torch.manual_seed(1000 * m + 10 * n + k)  # the same operands for every variant of a shape
    a = torch.rand(1 << (m + k), dtype=torch.float64, device="cuda")
    b = torch.rand(1 << (n + k), dtype=torch.float64, device="cuda")
    c = torch.empty(1 << (m + n), dtype=torch.float64, device="cuda")
    ws_bytes = 8 * max(min(1 << (m + n + 8), 1 << 27), 300 * 8192) + 8192
    ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device="cuda")
    # single launch between events (best of 5)
    single = None
    for _ in range(6):
        ms = (ctypes.c_float * 3)()
        run_join(m, n, k, a, b, c, ws, ws_bytes, None, ms)
        single = ms[1] if single is None or ms[1] < single else single
    out = c.clone()
    # back to back: R launches in one graph
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(reps):
            run_join(m, n, k, a, b, c, ws, ws_bytes, ctypes.c_void_p(side.cuda_stream), None)
    best = None
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        best = t if best is None or t < best else best
    out2 = c.clone()
    assert torch.equal(out, out2), "run-to-run difference"
    del graph
    tset(**DEFAULTS)
    return single, best, out


def report(m, n, k, label, single, b2b):
    flops = 2.0 * 2 ** (m + n + k)
    byts = 8.0 * (2 ** (m + k) + 2 ** (n + k) + 2 ** (m + n))
    tensor = flops / (FP64 * 1e12) > byts / (HBM * 1e9)
    def frac(ms):
        return (flops / (ms * 1e-3) / 1e12 / FP64) if tensor else (byts / (ms * 1e-3) / 1e9 / HBM)
    print("| %d | %d | %d | %s | %s | %.1f | %.3f | %.1f | %.3f |" % (m, n, k, label, "tensor" if tensor else "hbm", single * 1e3,
                                                                    frac(single), b2b * 1e3, frac(b2b)), flush=True)
    return {"m": m, "n": n, "k": k, "variant": label, "bound": "tensor" if tensor else "hbm", "single_us": single * 1e3,
            "single_frac": frac(single), "b2b_us": b2b * 1e3, "b2b_frac": frac(b2b)}


rows = []
print("| m | n | k | variant | bound | single-launch us | frac | back-to-back us | frac |")
print("|---|---|---|---|---|---|---|---|---|")
if what == "dot":
    for (m, n, k) in [(0, 0, 24), (0, 0, 26), (0, 0, 22), (1, 1, 23), (2, 1, 22), (3, 3, 20)]:
        for v in (10, 11, 12, 13):
            s0, b0, out = time_join(m, n, k, {"t256_ctas_log2": v}, 10)
            rows.append(report(m, n, k, "CTA per (output, chunk), 2^%d CTAs" % v, s0, b0))
if what == "midk":
    for (m, n, k) in [(15, 14, 5), (14, 12, 5), (13, 11, 5), (13, 10, 5), (12, 10, 5), (15, 14, 6), (13, 10, 6), (11, 11, 7)]:
        s0, b0, ref = time_join(m, n, k, {}, 5)
        rows.append(report(m, n, k, "table default", s0, b0))
        if k == 5:
            for lab, kn in (("warp-specialised 3-stage ring (k_gemm_dmma_ws)", {"ws_min_k": 5, "store_tile": 1}), ("one tile per CTA (k_gemm_dmma)", {"store_tile": 1})):
                s1, b1, out = time_join(m, n, k, kn, 5)
                assert torch.equal(out, ref)
                rows.append(report(m, n, k, lab, s1, b1))
        elif k <= 7:
            s1, b1, out = time_join(m, n, k, {"persist_max_k": 7, "store_tile": 0}, 5)
            assert torch.equal(out, ref)
            rows.append(report(m, n, k, "persistent whole-tile kernel (k_gemm_dmma_p)", s1, b1))
if what == "streamk_long":
    for (m, n, k) in [(11, 10, 11), (11, 10, 12), (11, 10, 13), (11, 10, 14), (11, 9, 13), (10, 10, 13), (10, 9, 13), (11, 8, 12), (11, 8, 10)]:
        reps = 5
        s0, b0, ref = time_join(m, n, k, {"streamk": 0}, reps)
        rows.append(report(m, n, k, "data-parallel / split-K", s0, b0))
        s1, b1, out = time_join(m, n, k, {"streamk": 2, "streamk_min_tiles_log2": 0, "streamk_max_tiles_log2": 40, "streamk_max_steps": 1 << 30}, reps)
        assert float(((out - ref).abs() / ref.abs().clamp_min(1e-300)).max()) < 1e-12
        rows.append(report(m, n, k, "stream-K", s1, b1))
if what in ("streamk", "all"):
    shapes = [(11, 10, 10), (11, 10, 9), (10, 10, 10), (10, 10, 8), (11, 11, 8), (11, 11, 10), (11, 11, 12), (12, 11, 10), (12, 12, 8),
              (12, 12, 11), (13, 11, 11), (10, 10, 12), (10, 9, 12), (9, 9, 12), (9, 9, 14), (8, 8, 12), (7, 7, 16), (8, 8, 16), (6, 6, 16)]
    for (m, n, k) in shapes:
        reps = 20 if m + n + k <= 32 else 5
        s0, b0, ref = time_join(m, n, k, {"streamk": 0}, reps)
        rows.append(report(m, n, k, "data-parallel / split-K", s0, b0))
        s1, b1, out = time_join(m, n, k, {"streamk": 2, "streamk_min_tiles_log2": 0, "streamk_max_tiles_log2": 40, "streamk_max_steps": 1 << 30}, reps)
        err = float(((out - ref).abs() / ref.abs().clamp_min(1e-300)).max())
        assert err < 1e-12, (m, n, k, err)
        rows.append(report(m, n, k, "stream-K", s1, b1))
        s2, b2, out = time_join(m, n, k, {}, reps)
        rows.append(report(m, n, k, "table default", s2, b2))
if what in ("store", "all"):
    for (m, n, k) in [(14, 14, 4), (15, 15, 4), (13, 11, 4), (12, 10, 4), (15, 14, 5), (14, 12, 5), (13, 11, 5), (13, 10, 5)]:
        ref = None
        variants = [("whole tile, then stores (k_gemm_dmma_p)", {"store_tile": 0}), ("table default", {})]
        if k == 5:
            variants = [("one tile per CTA (k_gemm_dmma)", {"store_tile": 1, "ws_min_k": 6}), ("warp-specialised 3-stage ring (k_gemm_dmma_ws)", {"store_tile": 1}),
                        ("table default", {})]
        for label, knobs in variants:
            s, b, out = time_join(m, n, k, knobs, 3)
            if ref is None:
                ref = out
            assert torch.equal(out, ref), (m, n, k, label)
            rows.append(report(m, n, k, label, s, b))
        del ref, out
        torch.cuda.empty_cache()
    # write-only and copy ceilings for the store-bound joins
    x = torch.empty(1 << 30, dtype=torch.float64, device="cuda")
    for name, fn, byts in (("fill 8 GiB", lambda: x.zero_(), 8.0 * (1 << 30)), ("copy 4+4 GiB", lambda: x[: 1 << 29].copy_(x[1 << 29:]), 8.0 * (1 << 30))):
        best = None
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1)
            best = t if best is None or t < best else best
        print("ceiling: %s %.3f ms = %.0f GB/s (%.3f of %.0f)" % (name, best, byts / best / 1e6, byts / best / 1e6 / HBM, HBM), flush=True)
        rows.append({"ceiling": name, "ms": best, "gbs": byts / best / 1e6})
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(REPO, "gpurun_out", "kernel_lab_%s.json" % what), "w"), indent=1)
