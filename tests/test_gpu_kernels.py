"""Kernel-level parity: single pairwise contractions and index permutations against numpy
(the oracle's `tensordot`), integer-valued inputs so results are exact in float64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tensordot_device(a, b, axes_a, axes_b, policy=0):
    import ctypes

    import torch

    from tensororder_b200 import cabi

    ta = torch.from_numpy(a.reshape(-1).copy()).cuda()
    tb = torch.from_numpy(b.reshape(-1).copy()).cuda()
    rank_c = a.ndim + b.ndim - 2 * len(axes_a)
    tc = torch.empty(1 << rank_c, dtype=torch.float64, device="cuda")
    ws_bytes = 8 * (a.size + b.size + max(tc.numel() << 4, min(tc.numel() << 8, 1 << 27))) + 1024
    ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device="cuda")
    aa = np.asarray(axes_a, dtype=np.int32)
    ab = np.asarray(axes_b, dtype=np.int32)
    ms = (ctypes.c_float * 3)()
    torch.cuda.synchronize()
    rc = cabi.lib.tob_tensordot_device(
        ta.data_ptr(), a.ndim, tb.data_ptr(), b.ndim,
        aa.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ab.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
        len(axes_a), tc.data_ptr(), ws.data_ptr(), ws_bytes, policy, None, ms)
    assert rc == 0, cabi.last_error()
    torch.cuda.synchronize()
    return tc.cpu().numpy().reshape((2,) * rank_c), list(ms)


CASES = [
    # (rank_a, rank_b, k)   covers: outer product, scalar operands, mat-vec, full dot, thread/warp/CTA generic, GEMM tiles
    (0, 0, 0), (3, 0, 0), (0, 4, 0), (3, 2, 0), (5, 5, 5), (10, 3, 1), (12, 2, 2), (3, 12, 2),
    (9, 9, 5), (14, 14, 14), (20, 20, 20), (16, 10, 8), (15, 13, 7), (13, 14, 6), (14, 12, 4),
    (18, 16, 9), (17, 17, 10), (20, 12, 6), (21, 18, 9), (13, 13, 1), (12, 12, 0), (22, 8, 8),
    (12, 11, 2), (11, 13, 3), (14, 10, 3), (16, 3, 1), (17, 2, 0),  # small-k GEMM (zero-filled K step) and x4 streaming kernel
    (14, 14, 8), (15, 14, 8), (16, 16, 10), (22, 22, 16),  # 64x64 tiles, 128x64 tiles, split-K with few tiles
    # short K with more tiles than CTA slots: the persistent kernel (partial K step, one and two K steps per tile,
    # one tile column, swapped operands)
    (15, 13, 2), (16, 14, 4), (17, 15, 5), (19, 9, 3), (11, 20, 4), (14, 13, 1),
]


@pytest.mark.parametrize("ra,rb,k", CASES)
@pytest.mark.parametrize("ready", [True, False])
def test_tensordot_matches_numpy(ra, rb, k, ready):
    rng = np.random.default_rng(1000 * ra + 10 * rb + k)
    a = rng.integers(0, 3, size=(2,) * ra).astype(np.float64)
    b = rng.integers(0, 3, size=(2,) * rb).astype(np.float64)
    if ready:  # contracted axes trailing, in pair order: no permutation needed
        axes_a = list(range(ra - k, ra))
        axes_b = list(range(rb - k, rb))
    else:
        axes_a = [int(x) for x in rng.permutation(ra)[:k]]
        axes_b = [int(x) for x in rng.permutation(rb)[:k]]
    want = np.tensordot(a, b, (axes_a, axes_b))
    got, ms = _tensordot_device(a, b, axes_a, axes_b)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    got2, _ = _tensordot_device(a, b, axes_a, axes_b, policy=1)  # generic kernels must agree with the GEMM
    assert np.array_equal(got2, want)


def test_tensordot_random_values_tolerance():
    rng = np.random.default_rng(7)
    a = rng.random((2,) * 18)
    b = rng.random((2,) * 17)
    axes_a = [3, 0, 11, 7, 16, 5, 9, 13, 1]
    axes_b = [2, 8, 0, 15, 4, 10, 6, 12, 16]
    want = np.tensordot(a, b, (axes_a, axes_b))
    got, ms = _tensordot_device(a, b, axes_a, axes_b)
    assert ms[2] == 1.0  # DMMA GEMM path
    assert np.allclose(got, want, rtol=1e-12, atol=0)  # float64 tolerance, stated: 1e-12 relative


def test_host_tensordot_entry():
    from tensororder_b200.api import B200API

    api = B200API()
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2, size=(2,) * 9).astype(np.float64)
    b = rng.integers(0, 2, size=(2,) * 3).astype(np.float64)
    # the shapes identify() really produces for the shipped instance
    assert np.array_equal(api.tensordot(a, b, ([6], [2])), np.tensordot(a, b, ([6], [2])))
    assert np.array_equal(api.tensordot(a, a, ([0, 4, 7], [5, 2, 1])), np.tensordot(a, a, ([0, 4, 7], [5, 2, 1])))


@pytest.mark.parametrize("rank", [0, 1, 5, 8, 11, 12, 16, 22])
def test_permute_matches_numpy_transpose(rank):
    import ctypes

    import torch

    from tensororder_b200 import cabi

    rng = np.random.default_rng(rank)
    x = rng.random((2,) * rank)
    for trial in range(3):
        perm = [int(p) for p in rng.permutation(rank)] if trial else list(range(rank))[::-1]
        tin = torch.from_numpy(x.reshape(-1).copy()).cuda()
        tout = torch.empty_like(tin)
        pa = np.asarray(perm, dtype=np.int32)
        rc = cabi.lib.tob_permute_device(tin.data_ptr(), tout.data_ptr(), rank,
                                         pa.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None, None)
        assert rc == 0, cabi.last_error()
        torch.cuda.synchronize()
        want = np.transpose(x, perm) if rank else x
        assert np.array_equal(tout.cpu().numpy().reshape(want.shape), want)


@pytest.fixture
def tuning():
    """Overrides single dispatch parameters (tob_tuning_set) for one test and restores the measured table afterwards."""
    import ctypes

    from tensororder_b200 import cabi

    saved = {}

    def set_(key, value):
        for one in (["gemm_min_out.%d" % k for k in range(1, 17)] if key == "gemm_min_out" else [key]):  # "gemm_min_out": every k
            if one not in saved:
                v = ctypes.c_double()
                assert cabi.lib.tob_tuning_get(one.encode(), ctypes.byref(v)) == 0, cabi.last_error()
                saved[one] = v.value
        assert cabi.lib.tob_tuning_set(key.encode(), float(value)) == 0, cabi.last_error()

    yield set_
    for key, value in saved.items():
        cabi.lib.tob_tuning_set(key.encode(), value)


# stream-K forced on every eligible join (the table restricts it to 64..256 tiles)
SK = {"streamk": 2, "streamk_min_tiles_log2": 0, "streamk_max_tiles_log2": 40, "streamk_max_steps": 1 << 30}

VARIANT_CASES = [
    # long-K joins (K >= 256 per split) on the warp-specialised kernel under forced splits, swapped operands
    ({"force_ksplit_log2": 0}, (18, 16, 9)), ({"force_ksplit_log2": 2}, (19, 18, 11)), ({"force_ksplit_log2": 1}, (16, 21, 10)),
    # short-K persistent kernel under a forced split
    ({"force_ksplit_log2": 1}, (18, 17, 6)), ({"persist_max_k": -1}, (16, 14, 4)),
    # deep split-K on few tiles (the mid-size class of the rank sweep)
    ({"force_ksplit_log2": 5}, (20, 20, 12)), ({"force_ksplit_log2": 7}, (22, 22, 16)), ({"force_ksplit_log2": 8}, (23, 22, 16)),
    # stream-K (k_gemm_dmma_sk) forced on every eligible join: ranges inside one tile (many partials per owner), ranges over
    # two tiles (config 3's dominant join), ranges over many tiles, swapped operands, a store raster order of the short-K kernel
    (SK, (19, 17, 10)), (SK, (21, 20, 10)),
    (SK, (19, 18, 8)), (SK, (22, 21, 9)),
    (SK, (16, 21, 10)), (SK, (20, 20, 12)),
    ({"store_group_log2": 0}, (16, 14, 4)), ({"store_group_log2": 2}, (17, 15, 5)),
    # K = 16: the row-streamed persistent kernel (table default) under a forced split, and the whole-tile kernel it replaced
    ({"force_ksplit_log2": 1}, (20, 19, 5)), ({"store_tile": 0}, (16, 14, 4)), ({"store_tile": 1}, (14, 17, 4)),
    # K = 32 on the row-streamed kernel (2048 tiles and more), under a forced split, and switched off
    ({"store_tile": 2}, (18, 17, 5)), ({"store_tile": 2, "force_ksplit_log2": 1}, (19, 17, 6)), ({"store_tile": 1}, (18, 17, 5)),
]


@pytest.mark.parametrize("knobs,shape", VARIANT_CASES)
@pytest.mark.parametrize("ready", [True, False])
def test_kernel_variants_match_numpy(knobs, shape, ready, tuning):
    ra, rb, k = shape
    for key, value in knobs.items():
        tuning(key, value)
    rng = np.random.default_rng(77 * ra + 5 * rb + k)
    a = rng.integers(0, 3, size=(2,) * ra).astype(np.float64)
    b = rng.integers(0, 3, size=(2,) * rb).astype(np.float64)
    if ready:
        axes_a, axes_b = list(range(ra - k, ra)), list(range(rb - k, rb))
    else:  # random axis order: after the operand permutation the OUTPUT interleave (mask_m) is still the plain one,
        axes_a = [int(x) for x in rng.permutation(ra)[:k]]  # interleaved outputs are covered by the whole-plan tests below
        axes_b = [int(x) for x in rng.permutation(rb)[:k]]
    want = np.tensordot(a, b, (axes_a, axes_b))
    got, ms = _tensordot_device(a, b, axes_a, axes_b)
    assert ms[2] == 1.0  # DMMA GEMM path
    assert np.array_equal(got, want)


@pytest.mark.parametrize("knobs", [{"gemm_min_out": 12}, {"gemm_min_out": 12, "persist_max_k": -1}, {"max_ksplit_log2": 0},
                                   SK, {"streamk": 0, "store_tile": 0}])
@pytest.mark.parametrize("name", ["vc150_lineflow", "vc170_lineflow", "vc150_mcc_factorflow", "vc200_lineflow"])
def test_kernel_variants_on_whole_plans(knobs, name, tuning):
    """Whole contraction trees under the variant kernels: the joins' outputs use arbitrary interleaves of the two
    operands' free indices (mask_m), which single tensordot calls do not produce; sliced variants run two slice lanes
    concurrently.  Three runs must agree bit for bit (a race between lanes shows up as run-to-run noise)."""
    from conftest import load_golden
    from tensororder_b200.api import CompiledPlan
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden(name)
    for key, value in knobs.items():
        tuning(key, value)
    for plan in [pp] + ([pp.variant(0)] if pp.variants else []):
        cp = CompiledPlan(flatten_plan(plan.as_execution_plan()))
        cp.upload()
        got = [cp.run() for _ in range(3)]
        cp.close()
        want = plan.expected.get("count", pp.expected.get("count"))
        assert got[0] == got[1] == got[2], (plan.name, got)
        assert abs(got[0] - want) <= 1e-9 * abs(want), (plan.name, got, want)
