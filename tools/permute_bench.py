"""Stand-alone index-permutation kernel: GB/s (16 B moved per element) for random and structured permutations."""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from tensororder_b200 import cabi
P32 = ctypes.POINTER(ctypes.c_int32)
for rank in (20, 24, 28, 30):
    x = torch.rand(1 << rank, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    rng = np.random.default_rng(rank)
    perms = {"random": [int(p) for p in rng.permutation(rank)], "reverse": list(range(rank))[::-1],
             "swap_halves": list(range(rank // 2, rank)) + list(range(rank // 2)),
             "rotate3": list(range(3, rank)) + [0, 1, 2]}
    for name, perm in perms.items():
        pa = np.asarray(perm, dtype=np.int32); best = 1e9
        for rep in range(4):
            ms = ctypes.c_float(0)
            rc = cabi.lib.tob_permute_device(x.data_ptr(), y.data_ptr(), rank, pa.ctypes.data_as(P32), None, ctypes.byref(ms))
            assert rc == 0, cabi.last_error()
            if rep: best = min(best, ms.value)
        print("rank %d %-12s %8.4f ms  %7.1f GB/s" % (rank, name, best, 16.0 * (1 << rank) / best / 1e6))
