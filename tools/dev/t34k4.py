import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import numpy as np, torch
from tensororder_b200 import cabi
P32 = ctypes.POINTER(ctypes.c_int32)
def run(m, n, k, a, b, c, ws, ws_bytes, ms):
    aa = np.arange(m, m + k, dtype=np.int32); ab = np.arange(n, n + k, dtype=np.int32)
    rc = cabi.lib.tob_tensordot_device(a.data_ptr(), m + k, b.data_ptr(), n + k, aa.ctypes.data_as(P32), ab.ctypes.data_as(P32), k,
                                       c.data_ptr(), ws.data_ptr(), ws_bytes, 0, None, ms)
    assert rc == 0, cabi.last_error()
for trial in range(2):
    for (m, n, k) in [(16, 16, 2), (15, 15, 4), (14, 14, 4)]:
        a = torch.rand(1 << (m + k), dtype=torch.float64, device="cuda")
        b = torch.rand(1 << (n + k), dtype=torch.float64, device="cuda")
        c = torch.empty(1 << (m + n), dtype=torch.float64, device="cuda")
        ws_bytes = 8 * min(1 << (m + n + 8), 1 << 28) + 4096
        ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device="cuda")
        for st in (1, 0, 1):
            cabi.lib.tob_tuning_set(b"store_tile", float(st))
            out = []
            for rep in range(6):
                ms = (ctypes.c_float * 3)()
                torch.cuda.synchronize()
                run(m, n, k, a, b, c, ws, ws_bytes, ms)
                out.append(round(ms[1], 4))
            print("trial", trial, (m, n, k), "store_tile", st, "c ptr %x" % c.data_ptr(), "a ptr %x" % a.data_ptr(), out, flush=True)
        del a, b, c, ws
        torch.cuda.empty_cache()
