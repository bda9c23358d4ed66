"""BASELINE config 5 — rank sweep of single pairwise contractions of binary-index tensors.

T = fL + fR + k distinct indices (20..34), k contracted (2..16), fL = ceil((T-k)/2).  Two axis placements:
"ready" (contracted axes trailing in both operands = already canonical, no permutation) and "random"
(seeded random axis order of each operand -> the stand-alone permute kernel runs first).  Reports the
permute kernel (16 B moved per element), the contraction kernel (DMMA GEMM or generic) and their
roofline fractions.  Usage: python tools/rank_sweep.py [--quick] [--T=28,30] [--K=2,4]"""
import ctypes, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from tensororder_b200 import cabi

HBM = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6650.0
FP64 = 35.49
quick = "--quick" in sys.argv
Ts = [20, 24, 28] if quick else list(range(20, 35, 2))
Ks = [2, 8, 16] if quick else [2, 4, 8, 12, 16]
for arg in sys.argv[1:]:  # --T=28,30  --K=2,4 : sub-grid
    if arg.startswith("--T="):
        Ts = [int(x) for x in arg[4:].split(",")]
    if arg.startswith("--K="):
        Ks = [int(x) for x in arg[4:].split(",")]
P32 = ctypes.POINTER(ctypes.c_int32)
rows = []
for T in Ts:
    for k in Ks:
        if k > T - 2:
            continue
        fL = (T - k + 1) // 2
        fR = T - k - fL
        ra, rb, rc = fL + k, fR + k, fL + fR
        a = torch.rand(1 << ra, dtype=torch.float64, device="cuda")
        b = torch.rand(1 << rb, dtype=torch.float64, device="cuda")
        c = torch.empty(1 << rc, dtype=torch.float64, device="cuda")
        ws_bytes = 8 * ((1 << ra) + (1 << rb) + min(1 << (rc + 8), 1 << 28)) + 4096  # permuted operands + split-K partials (up to 2^8 splits)
        ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device="cuda")
        rng = np.random.default_rng(1000 * T + k)
        for placement in ("ready", "random"):
            if placement == "ready":
                axes_a, axes_b = list(range(ra - k, ra)), list(range(rb - k, rb))
            else:
                axes_a = [int(x) for x in rng.permutation(ra)[:k]]
                axes_b = [int(x) for x in rng.permutation(rb)[:k]]
            aa, ab = np.asarray(axes_a, dtype=np.int32), np.asarray(axes_b, dtype=np.int32)
            best = None
            for rep in range(10 if T >= 32 else 4):  # fresh multi-GB outputs: ~6 passes until first-touch effects are gone
                ms = (ctypes.c_float * 3)()
                torch.cuda.synchronize()
                rc_ = cabi.lib.tob_tensordot_device(a.data_ptr(), ra, b.data_ptr(), rb, aa.ctypes.data_as(P32),
                                                    ab.ctypes.data_as(P32), k, c.data_ptr(), ws.data_ptr(), ws_bytes, 0, None, ms)
                assert rc_ == 0, cabi.last_error()
                if rep and (best is None or ms[0] + ms[1] < best[0] + best[1]):
                    best = (ms[0], ms[1], ms[2])
            flops = 2.0 * 2 ** (fL + fR + k)
            byts = 8.0 * (2 ** ra + 2 ** rb + 2 ** rc)
            tf = flops / (best[1] * 1e-3) / 1e12
            gb = byts / (best[1] * 1e-3) / 1e9
            bound = "tensor" if flops / (FP64 * 1e12) > byts / (HBM * 1e9) else "hbm"
            frac = tf / FP64 if bound == "tensor" else gb / HBM
            perm_gb = (16.0 * (2 ** ra + 2 ** rb) / (best[0] * 1e-3) / 1e9) if placement == "random" and best[0] > 0 else None
            rows.append({"T": T, "k": k, "fL": fL, "fR": fR, "placement": placement, "kernel": "gemm" if best[2] == 1 else "generic",
                         "contract_ms": best[1], "tflops": tf, "gbs": gb, "bound": bound, "frac": frac,
                         "permute_ms": best[0] if placement == "random" else 0.0, "permute_gbs": perm_gb,
                         "permute_frac": (perm_gb / HBM) if perm_gb else None})
        del a, b, c, ws
        torch.cuda.empty_cache()
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(REPO, "gpurun_out", "rank_sweep.json"), "w"), indent=1)
print("| T | k | fL | fR | placement | kernel | contract ms | TF/s | GB/s | bound | frac | permute ms | permute GB/s (frac of HBM) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print("| %d | %d | %d | %d | %s | %s | %.4f | %.2f | %.0f | %s | %.3f | %s | %s |" % (
        r["T"], r["k"], r["fL"], r["fR"], r["placement"], r["kernel"], r["contract_ms"], r["tflops"], r["gbs"], r["bound"], r["frac"],
        ("%.4f" % r["permute_ms"]) if r["placement"] == "random" else "-",
        ("%.0f (%.2f)" % (r["permute_gbs"], r["permute_frac"])) if r["permute_gbs"] else "-"))
