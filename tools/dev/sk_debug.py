import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import numpy as np, torch
from tensororder_b200 import cabi
P32 = ctypes.POINTER(ctypes.c_int32)
m, n, k = [int(x) for x in sys.argv[1:4]]
cabi.lib.tob_tuning_set(b"streamk", 2.0); cabi.lib.tob_tuning_set(b"streamk_min_tiles_log2", 0.0)
torch.manual_seed(1)
a = torch.randint(0, 3, (1 << m, 1 << k), device="cuda").double()
b = torch.randint(0, 3, (1 << n, 1 << k), device="cuda").double()
want = a @ b.T
c = torch.full((1 << (m + n),), -7.0, dtype=torch.float64, device="cuda")
ws_bytes = 8 * 300 * 8192 + (1 << 20)
ws = torch.zeros(ws_bytes // 8, dtype=torch.float64, device="cuda")
aa = np.arange(m, m + k, dtype=np.int32); ab = np.arange(n, n + k, dtype=np.int32)
for rep in range(2):
    ms = (ctypes.c_float * 3)()
    rc = cabi.lib.tob_tensordot_device(a.data_ptr(), m + k, b.data_ptr(), n + k, aa.ctypes.data_as(P32), ab.ctypes.data_as(P32), k,
                                       c.data_ptr(), ws.data_ptr(), ws_bytes, 0, None, ms)
    assert rc == 0, cabi.last_error()
    torch.cuda.synchronize()
    got = c.view(1 << m, 1 << n)
    bad = (got != want)
    print("rep", rep, "ms", ms[1], "kind", ms[2], "wrong elements", int(bad.sum()), "of", bad.numel())
    tiles = bad.view((1 << m) // 128, 128, (1 << n) // 64, 64).any(dim=3).any(dim=1)
    print("wrong tiles:", int(tiles.sum()), "of", tiles.numel())
    idx = tiles.nonzero()[:12].tolist()
    for tm, tn in idx:
        g = got[tm * 128:(tm + 1) * 128, tn * 64:(tn + 1) * 64]; w = want[tm * 128:(tm + 1) * 128, tn * 64:(tn + 1) * 64]
        print("  tile", tm, tn, "tile id", tn * min(16, (1 << m) // 128) + tm, "sum got/want %.4f" % float(g.sum() / w.sum()), "untouched", int((g == -7).sum()))
    flags = ws.view(torch.int32)[: 1024]
    print("flags nonzero:", flags.nonzero().flatten().tolist()[:20])
