"""Kernel variants side by side on single joins (GEMM-ready operands, `tob_tensordot_device` under
`tob_tuning_set` overrides): the long-K DMMA GEMM with the LDGSTS feed vs the 2-D tensor-map (TMA) feed.  (Round 2 also
compared a shared-memory-staged epilogue of the persistent short-K kernel here: profiles/r02b_kernel_lab_tma_staged.md.)
Usage (GPU box): python tools/kernel_lab.py  -> table on stdout, gpurun_out/kernel_lab.json"""
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tensororder_b200 import cabi  # noqa: E402

lib = cabi.lib
P32 = ctypes.POINTER(ctypes.c_int32)
peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6552.6))
FP64 = 35.49


def tset(key, value):
    assert lib.tob_tuning_set(key.encode(), float(value)) == 0, cabi.last_error()


def time_join(m, n, k, reps=5):
    ra, rb, rc = m + k, n + k, m + n
    a = torch.rand(1 << ra, dtype=torch.float64, device="cuda")
    b = torch.rand(1 << rb, dtype=torch.float64, device="cuda")
    c = torch.empty(1 << rc, dtype=torch.float64, device="cuda")
    ws_n = min(1 << (rc + 4), 1 << 28)
    ws = torch.empty(ws_n + 512, dtype=torch.float64, device="cuda")
    aa = np.arange(ra - k, ra, dtype=np.int32)
    ab = np.arange(rb - k, rb, dtype=np.int32)
    times = []
    for rep in range(reps + 1):
        ms = (ctypes.c_float * 3)()
        torch.cuda.synchronize()
        rc_ = lib.tob_tensordot_device(a.data_ptr(), ra, b.data_ptr(), rb, aa.ctypes.data_as(P32), ab.ctypes.data_as(P32), k,
                                       c.data_ptr(), ws.data_ptr(), 8 * ws_n, 0, None, ms)
        assert rc_ == 0, cabi.last_error()
        if rep:
            times.append(ms[1])
    chk = float(c[:: max(1, c.numel() // 4096)].sum().item())
    del a, b, c, ws
    torch.cuda.empty_cache()
    return min(times), sorted(times)[len(times) // 2], chk


rows = []
print("| join (m,n,k) | variant | best ms | median ms | TF/s | GB/s | frac |")
print("|---|---|---|---|---|---|---|")
for (m, n, k) in [(11, 11, 12), (12, 11, 12), (13, 12, 8), (12, 12, 10), (14, 13, 10), (10, 10, 14), (13, 12, 16), (11, 10, 10), (9, 9, 16)]:
    for feed in (0, 1):
        tset("gemm_feed", feed)
        best, med, chk = time_join(m, n, k)
        fl = 2.0 * 2.0 ** (m + n + k)
        by = 8.0 * (2.0 ** (m + k) + 2.0 ** (n + k) + 2.0 ** (m + n))
        tf, gb = fl / best / 1e9, by / best / 1e6
        rows.append({"m": m, "n": n, "k": k, "variant": "tma" if feed else "ldgsts", "ms": best, "median_ms": med, "tflops": tf,
                     "gbs": gb, "frac": tf / FP64, "checksum": chk})
        print("| (%d,%d,%d) | %s | %.4f | %.4f | %.2f | %.0f | %.3f of FP64 |" % (m, n, k, "TMA feed" if feed else "LDGSTS feed", best, med, tf, gb, tf / FP64))
tset("gemm_feed", 0)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(REPO, "gpurun_out", "kernel_lab.json"), "w"), indent=1)
