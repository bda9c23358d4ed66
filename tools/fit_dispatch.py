"""Generates tensororder_b200/csrc/tob_dispatch_table.h from measurements on the GPU (north_star: "the choice
per node made from measured counters").  Runs single joins through `tob_tensordot_device` (GEMM-ready operands)
under `tob_tuning_set` overrides and derives

  A. the split-K time model's constants (alone_frac, gemm_fix_us, reduce_gbs, reduce_fix_us): every power-of-two
     split of a grid of GEMM shapes is timed, the constants are fitted to the measured times (least squares on
     log(model / measured)), and the table prints the split the fitted model picks next to the measured best;
  B. the generic <-> DMMA GEMM crossover as a table by k (gemm_min_out[k]: the smallest m + n from which the GEMM wins);
  C. the generic kernel classes (t1_max_k, t1_small_*, t32_max_k) and the persistent short-K range (persist_max_k);
  D. the stream-K range (streamk_min/max_tiles_log2: tile counts where k_gemm_dmma_sk beats the best split of the
     one-tile-per-CTA kernel) and its fixed cost (streamk_fix_us), and the K = 16 store kernel (store_tile).

Usage (GPU box):  python tools/fit_dispatch.py [--quick]
Writes gpurun_out/dispatch_fit.json (raw measurements) and gpurun_out/tob_dispatch_table.h (copy it over
tensororder_b200/csrc/tob_dispatch_table.h and rebuild)."""
import ctypes
import datetime
import itertools
import json
import math
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tensororder_b200 import cabi  # noqa: E402

lib = cabi.lib
P32 = ctypes.POINTER(ctypes.c_int32)
quick = "--quick" in sys.argv
DEFAULTS = {}


def tset(key, value):
    assert lib.tob_tuning_set(key.encode(), float(value)) == 0, cabi.last_error()


def tget(key):
    v = ctypes.c_double()
    assert lib.tob_tuning_get(key.encode(), ctypes.byref(v)) == 0, cabi.last_error()
    return v.value


KEYS = ["gemm_min_free", "gemm_min_k", "gemm_smallk_min_free", "t1_max_k", "t1_small_out", "t1_small_max_k",
        "t32_max_k", "t32_min_out", "persist_max_k", "sm_gflops", "alone_frac", "gemm_fix_us", "reduce_gbs",
        "reduce_fix_us", "max_ksplit_log2", "min_k_per_split_log2", "force_ksplit_log2", "streamk", "streamk_min_tiles_log2",
        "streamk_max_tiles_log2", "streamk_max_steps", "streamk_fix_us", "store_tile", "ws_min_k"] + ["gemm_min_out.%d" % i for i in range(17)]
for k_ in KEYS:
    DEFAULTS[k_] = tget(k_)


def restore():
    for k_, v in DEFAULTS.items():
        tset(k_, v)


_bufs = {}


def buf(n_doubles, tag):
    key = (tag, n_doubles)
    if key not in _bufs:
        for old in [k2 for k2 in _bufs if k2[0] == tag]:
            del _bufs[old]
        torch.cuda.empty_cache()
        _bufs[key] = torch.rand(n_doubles, dtype=torch.float64, device="cuda")
    return _bufs[key]


def time_join(m, n, k, policy=0, reps=5, ws_log2=None):
    """Best-of-reps contraction time (us) and the kernel kind of C[2^(m+n)] = A[2^(m+k)] . B[2^(n+k)]."""
    ra, rb, rc = m + k, n + k, m + n
    a, b, c = buf(1 << ra, "a"), buf(1 << rb, "b"), buf(1 << rc, "c")
    ws_n = max(1 << (ws_log2 if ws_log2 is not None else min(rc + 8, 28)), 300 * 8192)  # >= one stream-K slot per CTA
    ws = buf(ws_n + 512, "ws")
    aa = np.arange(ra - k, ra, dtype=np.int32)
    ab = np.arange(rb - k, rb, dtype=np.int32)
    best, kind = None, None
    for rep in range(reps + 1):
        ms = (ctypes.c_float * 3)()
        torch.cuda.synchronize()
        rc_ = lib.tob_tensordot_device(a.data_ptr(), ra, b.data_ptr(), rb, aa.ctypes.data_as(P32), ab.ctypes.data_as(P32), k,
                                       c.data_ptr(), ws.data_ptr(), 8 * ws_n, policy, None, ms)
        assert rc_ == 0, cabi.last_error()
        if rep and (best is None or ms[1] < best):
            best = ms[1]
        kind = int(ms[2])
    return best * 1e3, kind


def model_us(m, n, k, c):
    lib.tob_gemm_time_model_us.restype = ctypes.c_double
    lib.tob_gemm_time_model_us.argtypes = [ctypes.c_int32] * 6
    return lib.tob_gemm_time_model_us(m, n, k, min(m, 7), min(n, 6), c)


out = {"when": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ"), "gpu": torch.cuda.get_device_name(0)}

# ---------------------------------------------------------------------------------------------------
# A. split-K
# ---------------------------------------------------------------------------------------------------
shapes = []
for T in ([28, 32] if quick else [26, 28, 30, 32, 34]):
    for k in ([8, 12, 16] if quick else [8, 10, 12, 14, 16]):
        m = (T - k + 1) // 2
        n = T - k - m
        if n >= 6:
            shapes.append((m, n, k))
shapes += [(11, 10, 10), (11, 10, 9), (10, 10, 9), (12, 11, 12), (14, 13, 10)]
rows = []
for (m, n, k) in shapes:
    for c in range(0, 9):
        if k - c < 6 or (m + n + c) > 28:
            break
        tset("force_ksplit_log2", c)
        us, kind = time_join(m, n, k)
        rows.append({"m": m, "n": n, "k": k, "c": c, "us": us, "kind": kind})
restore()
out["splitk"] = rows
print("split-K calibration: %d timings" % len(rows))

# fit: grid search, then report
grid = {"alone_frac": [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.9, 1.0], "gemm_fix_us": [2, 3, 4, 5, 6, 8, 10, 12],
        "reduce_gbs": [1500, 2000, 2500, 3000, 3500, 4000, 4500, 5500], "reduce_fix_us": [1, 2, 3, 4, 6, 8]}
best_fit = None
for vals in itertools.product(*grid.values()):
    for key, v in zip(grid.keys(), vals):
        tset(key, v)
    err = 0.0
    for r in rows:
        err += math.log(model_us(r["m"], r["n"], r["k"], r["c"]) / r["us"]) ** 2
    if best_fit is None or err < best_fit[0]:
        best_fit = (err, dict(zip(grid.keys(), vals)))
restore()
fit = best_fit[1]
out["fit"] = {"rms_log_error": math.sqrt(best_fit[0] / len(rows)), **fit}
print("fitted:", out["fit"])
for key, v in fit.items():
    tset(key, v)
    DEFAULTS[key] = v
print("| m | n | k | measured us by split c | measured best c | model picks c | loss vs best |")
print("|---|---|---|---|---|---|---|")
picks = []
for (m, n, k) in shapes:
    rs = [r for r in rows if (r["m"], r["n"], r["k"]) == (m, n, k)]
    meas_best = min(rs, key=lambda r: r["us"])
    pick, bt = 0, model_us(m, n, k, 0)
    for r in rs[1:]:
        t = model_us(m, n, k, r["c"])
        if t < bt * 0.97:
            pick, bt = r["c"], t
    got = [r for r in rs if r["c"] == pick][0]["us"]
    picks.append({"m": m, "n": n, "k": k, "best_c": meas_best["c"], "best_us": meas_best["us"], "model_c": pick, "model_c_us": got,
                  "tflops_at_pick": 2.0 * 2.0 ** (m + n + k) / got / 1e6})
    print("| %d | %d | %d | %s | %d (%.1f) | %d (%.1f) | %.1f%% |" % (
        m, n, k, " ".join("%d:%.1f" % (r["c"], r["us"]) for r in rs), meas_best["c"], meas_best["us"], pick, got,
        100 * (got / meas_best["us"] - 1)))
out["splitk_picks"] = picks

# ---------------------------------------------------------------------------------------------------
# B. generic <-> GEMM crossover, by k: the smallest m + n from which the GEMM kernel is at least as fast as the best
#    generic kernel for that and every larger measured output
# ---------------------------------------------------------------------------------------------------
cross = []
tset("gemm_min_out", 0)  # every k: GEMM whenever the free sides allow a tile
for k in ([2, 4, 8] if quick else [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16]):
    lo = 14 if k < 4 else 12
    for outs in range(lo, 23 if k <= 8 else 21):
        m = (outs + 1) // 2
        n = outs - m
        if n < (7 if k < 4 else 6):
            continue
        g, kind_g = time_join(m, n, k, policy=0)
        s_, _ = time_join(m, n, k, policy=1)
        cross.append({"m": m, "n": n, "k": k, "outs": outs, "gemm_us": g, "generic_us": s_, "gemm_kind": kind_g})
restore()
out["crossover"] = cross
min_out = [99] + [int(DEFAULTS["gemm_min_out.%d" % i]) for i in range(1, 17)]
for k in sorted(set(r["k"] for r in cross)):
    rs = sorted([r for r in cross if r["k"] == k and r["gemm_kind"] == 1], key=lambda r: -r["outs"])
    best = None
    for r in rs:
        if r["gemm_us"] <= r["generic_us"] * 1.03:
            best = r["outs"]
        else:
            break
    if best is not None:
        min_out[k] = best
    elif rs:
        min_out[k] = rs[0]["outs"] + 1
for k in range(1, 17):  # unmeasured k: the nearest measured neighbour below
    if k not in set(r["k"] for r in cross):
        below = [q for q in set(r["k"] for r in cross) if q < k]
        if below:
            min_out[k] = min_out[max(below)]
print("crossover: gemm_min_out by k ->", min_out)
for r in cross:
    print("  m=%d n=%d k=%d  gemm %.1f us  generic %.1f us  (%s)" % (r["m"], r["n"], r["k"], r["gemm_us"], r["generic_us"],
                                                                  "gemm kernel" if r["gemm_kind"] == 1 else "generic both"))

# ---------------------------------------------------------------------------------------------------
# C. generic classes and the persistent short-K range
# ---------------------------------------------------------------------------------------------------
classes = []
for outs in (12, 16, 20):
    for k in range(4, 13):
        m = (outs + 1) // 2
        n = outs - m
        res = {}
        for name, t1, t32 in (("t1", 99, 99), ("t32", -1, 99), ("t256", -1, -1)):
            if name == "t1" and k > 6:
                continue  # the one-thread kernel walks K <= 64 only
            tset("t1_max_k", t1)
            tset("t32_max_k", t32)
            tset("t32_min_out", 0)
            res[name], _ = time_join(m, n, k, policy=1)
        restore()
        classes.append({"outs": outs, "k": k, **res})
        print("  generic classes outs=2^%d k=%d: %s" % (outs, k, {a: round(b, 1) for a, b in res.items()}))
out["generic_classes"] = classes
t1_max_k = max([r["k"] for r in classes if r["outs"] >= 16 and "t1" in r and r["t1"] <= min(r["t32"], r["t256"]) * 1.02] or [int(DEFAULTS["t1_max_k"])])
small = [r for r in classes if r["outs"] == 12 and "t1" in r]
t1_small_max_k = max([r["k"] for r in small if r["t1"] <= min(r["t32"], r["t256"]) * 1.02] or [3])
t32_max_k = max([r["k"] for r in classes if r["outs"] >= 12 and r["t32"] <= r["t256"] * 1.02] or [int(DEFAULTS["t32_max_k"])])
persist = []
for k in range(1, 8):
    m, n = (32 - k + 1) // 2, (32 - k) // 2
    tset("persist_max_k", 99)
    p_us, _ = time_join(m, n, k)
    tset("persist_max_k", -1)
    o_us, _ = time_join(m, n, k)
    restore()
    persist.append({"m": m, "n": n, "k": k, "persistent_us": p_us, "one_tile_per_cta_us": o_us})
    print("  persistent short-K m=%d n=%d k=%d: %.1f vs %.1f us" % (m, n, k, p_us, o_us))
out["persistent"] = persist
persist_max_k = max([r["k"] for r in persist if r["persistent_us"] <= r["one_tile_per_cta_us"] * 1.01] or [int(DEFAULTS["persist_max_k"])])

# ---------------------------------------------------------------------------------------------------
# D. stream-K range and fixed cost; the K = 16 store kernel
# ---------------------------------------------------------------------------------------------------
sk_rows = []
for tl in range(4, 11):  # 16 .. 1024 tiles of 128x64
    m, n = 7 + (tl + 1) // 2, 6 + tl // 2
    for k in (10, 11, 12, 13):
        if m + n + k > 34:
            continue
        tset("streamk", 0)
        dp_us, _ = time_join(m, n, k)
        tset("streamk", 2); tset("streamk_min_tiles_log2", 0); tset("streamk_max_tiles_log2", 40); tset("streamk_max_steps", 1 << 30)
        sk_us, _ = time_join(m, n, k)
        restore()
        steps = math.ceil((1 << (tl + k - 4)) / 296.0)
        fix = sk_us - steps * (2.0 * 2 ** 17 / (DEFAULTS["sm_gflops"] * 1e3 / 2.0)) - DEFAULTS["gemm_fix_us"]
        sk_rows.append({"m": m, "n": n, "k": k, "tiles_log2": tl, "steps_per_cta": steps, "best_split_us": dp_us, "streamk_us": sk_us,
                        "residual_fix_us": fix})
        print("  stream-K m=%d n=%d k=%d (%d tiles): best split %.1f us, stream-K %.1f us" % (m, n, k, 1 << tl, dp_us, sk_us))
out["streamk"] = sk_rows
# the range: from the first to the last tile count where stream-K wins by >= 3 % at some K, as long as it is no worse than
# 1 % everywhere in between (inside the range the time model decides per join); the fixed cost: the smallest residual
# measured time - K steps * step time - launch cost over the winning joins (optimistic, the 3 % margin covers the rest)
wins = sorted({r["tiles_log2"] for r in sk_rows if r["streamk_us"] < 0.97 * r["best_split_us"]})
tie_or_win = {r["tiles_log2"] for r in sk_rows if r["streamk_us"] <= 1.01 * r["best_split_us"]}
sk_min = wins[0] if wins else 99
sk_max = wins[-1] if wins else 99
while wins and not all(t in tie_or_win for t in range(sk_min, sk_max + 1)):
    sk_max -= 1
win_fix = [r["residual_fix_us"] for r in sk_rows if sk_min <= r["tiles_log2"] <= sk_max and r["streamk_us"] < 0.97 * r["best_split_us"]]
sk_fix = int(math.ceil(min(win_fix))) if win_fix else int(DEFAULTS["streamk_fix_us"])
# K steps per CTA up to which it still wins (long K: the one-tile-per-CTA grid catches up), as the next power of two
win_steps = [r["steps_per_cta"] for r in sk_rows if sk_min <= r["tiles_log2"] <= sk_max and r["streamk_us"] < 0.97 * r["best_split_us"]]
sk_steps = 1 << int(math.ceil(math.log2(max(win_steps)))) if win_steps else int(DEFAULTS["streamk_max_steps"])
store = {}
for v in (0, 1, 2):  # 0: whole tile then stores; 1: rows streamed at K = 16; 2: also at K = 32 (>= 2048 tiles)
    tset("store_tile", v)
    store[v] = {"k4": [time_join(m, n, 4)[0] for (m, n) in ((14, 14), (15, 14), (13, 12))],
                "k5": [time_join(m, n, 5)[0] for (m, n) in ((15, 14), (14, 12), (13, 11))]}
    restore()
out["store_tile"] = store
# K = 32 below 2048 tiles: warp-specialised ring (ws_min_k = 5) or one CTA barrier per K step (6)
wsk = {}
for v in (5, 6):
    tset("ws_min_k", v)
    wsk[v] = [time_join(m, n, 5)[0] for (m, n) in ((13, 10), (12, 10), (12, 11))]
    restore()
out["ws_min_k"] = wsk
ws_min_k = 5 if sum(wsk[5]) < sum(wsk[6]) else 6
store_tile = 0
if sum(store[1]["k4"]) < sum(store[0]["k4"]):
    store_tile = 2 if sum(store[2]["k5"]) < 0.99 * sum(store[1]["k5"]) else 1
print("  store kernels: K = 16 whole tile %s us, row-streamed %s us; K = 32 one tile per CTA %s us, row-streamed %s us" % (
    ["%.1f" % x for x in store[0]["k4"]], ["%.1f" % x for x in store[1]["k4"]], ["%.1f" % x for x in store[1]["k5"]], ["%.1f" % x for x in store[2]["k5"]]))

# ---------------------------------------------------------------------------------------------------
table = {
    "GEMM_MIN_FREE": int(DEFAULTS["gemm_min_free"]), "GEMM_MIN_K": int(DEFAULTS["gemm_min_k"]),
    "GEMM_SMALLK_MIN_FREE": int(DEFAULTS["gemm_smallk_min_free"]),
    "GEMM_MIN_OUT_BY_K": "{" + ", ".join(str(v) for v in min_out) + "}",
    "T1_MAX_K": min(t1_max_k, 6), "T1_SMALL_OUT": 13, "T1_SMALL_MAX_K": min(t1_small_max_k, 6),
    "T32_MAX_K": t32_max_k, "T32_MIN_OUT": int(DEFAULTS["t32_min_out"]),
    "PERSIST_MAX_K": persist_max_k, "SM_GFLOPS": DEFAULTS["sm_gflops"], "ALONE_FRAC": fit["alone_frac"],
    "GEMM_FIX_US": float(fit["gemm_fix_us"]), "REDUCE_GBS": float(fit["reduce_gbs"]), "REDUCE_FIX_US": float(fit["reduce_fix_us"]),
    "MAX_KSPLIT_LOG2": int(DEFAULTS["max_ksplit_log2"]), "MIN_K_PER_SPLIT_LOG2": int(DEFAULTS["min_k_per_split_log2"]),
    "STREAMK": 1 if wins else 0, "STREAMK_MIN_TILES_LOG2": sk_min if wins else int(DEFAULTS["streamk_min_tiles_log2"]),
    "STREAMK_MAX_TILES_LOG2": sk_max if wins else int(DEFAULTS["streamk_max_tiles_log2"]), "STREAMK_FIX_US": sk_fix, "STREAMK_MAX_STEPS": sk_steps,
    "STORE_TILE": store_tile, "WS_MIN_K": ws_min_k,
}
out["table"] = table
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(REPO, "gpurun_out", "dispatch_fit.json"), "w"), indent=1)
src = open(os.path.join(REPO, "tensororder_b200", "csrc", "tob_dispatch_table.h")).read().split("\n")
lines = []
for ln in src:
    if ln.startswith("// generated:"):
        ln = "// generated: %s on %s by tools/fit_dispatch.py (fit rms log error %.3f over %d timings; raw: profiles/r02i_dispatch_fit.json)" % (
            out["when"], out["gpu"], out["fit"]["rms_log_error"], len(rows))
    if ln.startswith("#define TOB_TUNE_"):
        name = ln.split()[1][len("TOB_TUNE_"):]
        if name in table:
            comment = ln[ln.index("//"):] if "//" in ln else ""
            v = table[name]
            txt = v if isinstance(v, str) else (("%d" % v) if isinstance(v, int) else ("%.4g" % v if v < 100 else "%.1f" % v))
            ln = ("#define TOB_TUNE_%s %s" % (name, txt)).ljust(42) + ("  " + comment if isinstance(v, str) else comment)
    lines.append(ln)
open(os.path.join(REPO, "gpurun_out", "tob_dispatch_table.h"), "w").write("\n".join(lines))
print("wrote gpurun_out/dispatch_fit.json and gpurun_out/tob_dispatch_table.h")
