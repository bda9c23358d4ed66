"""Runs ONE pairwise contraction a few times (GEMM-ready operands) — the command `ncu` wraps for a per-kernel capture.
Usage: python tools/one_join.py M N K [reps] [key=value ...]   (key=value: tob_tuning_set overrides)"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tensororder_b200 import cabi  # noqa: E402

m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 and "=" not in sys.argv[4] else 3
for kv in sys.argv[4:]:
    if "=" in kv:
        key, value = kv.split("=")
        assert cabi.lib.tob_tuning_set(key.encode(), float(value)) == 0, cabi.last_error()
ra, rb, rc = m + k, n + k, m + n
a = torch.rand(1 << ra, dtype=torch.float64, device="cuda")
b = torch.rand(1 << rb, dtype=torch.float64, device="cuda")
c = torch.empty(1 << rc, dtype=torch.float64, device="cuda")
ws_n = min(1 << (rc + 4), 1 << 28)
ws = torch.empty(ws_n + 512, dtype=torch.float64, device="cuda")
aa = np.arange(ra - k, ra, dtype=np.int32)
ab = np.arange(rb - k, rb, dtype=np.int32)
P32 = ctypes.POINTER(ctypes.c_int32)
for _ in range(reps):
    ms = (ctypes.c_float * 3)()
    rc_ = cabi.lib.tob_tensordot_device(a.data_ptr(), ra, b.data_ptr(), rb, aa.ctypes.data_as(P32), ab.ctypes.data_as(P32), k,
                                        c.data_ptr(), ws.data_ptr(), 8 * ws_n, 0, None, ms)
    assert rc_ == 0, cabi.last_error()
torch.cuda.synchronize()
print("join m=%d n=%d k=%d: %.4f ms (kernel kind %d)" % (m, n, k, ms[1], int(ms[2])))
