"""Per-join verification on the GPU: every op of a compiled plan is executed (tob_plan_debug_run) and its output
read back (tob_plan_debug_read) and compared with the numpy interpreter of the same program (tests/program_sim.py
semantics), so a wrong kernel is pinned to one join.  Usage:
    python tools/verify_ops.py NAME[:VARIANT] [slice] [key=value ...]      (key=value: tob_tuning_set overrides)"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402

from program_sim import pdep_table, upload_leaves  # noqa: E402
from tensororder_b200 import cabi  # noqa: E402
from tensororder_b200.api import CompiledPlan  # noqa: E402
from tensororder_b200.flatten import flatten_plan  # noqa: E402
from tensororder_b200.plan_format import PortablePlan  # noqa: E402


def verify(pp, slice_id=0, verbose=True, rtol=1e-12):
    flat = flatten_plan(pp.as_execution_plan())
    cp = CompiledPlan(flat)
    cp.upload()
    desc = cp.describe()
    ops = desc["invariant_ops"] + desc["slice_ops"]
    leaves = upload_leaves(desc, flat)
    arena = np.full(max(desc["arena_doubles"], 1), np.nan)
    leaf_off = []
    for L in desc["leaves"]:
        off = 0
        for ib, ab in zip(L["slice_id_bit"], L["slice_addr_bit"]):
            off |= ((slice_id >> ib) & 1) << ab
        leaf_off.append(off)

    def operand(ref, size):
        if ref["space"] == 0:
            off = ref["offset"] + (leaf_off[ref["leaf"]] if ref["leaf"] >= 0 else 0)
            return leaves[off: off + size]
        return arena[ref["offset"]: ref["offset"] + size]

    def sim(op):
        m, n, k = op["m"], op["n"], op["k"]
        A = operand(op["a"], 1 << (m + k)).reshape(1 << m, 1 << k)
        B = operand(op["b"], 1 << (n + k)).reshape(1 << n, 1 << k)
        Cm = A @ B.T
        mask_n = ~op["mask_m"] & ((1 << (m + n)) - 1)
        addr = pdep_table(m, op["mask_m"])[:, None] | pdep_table(n, mask_n)[None, :]
        out = arena[op["c_offset"]: op["c_offset"] + (1 << (m + n))]
        out[addr.reshape(-1)] = Cm.reshape(-1)

    def read(offset, n):
        buf = np.empty(n)
        rc = cabi.lib.tob_plan_debug_read(cp._handle, 1, offset, n, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        assert rc == 0, cabi.last_error()
        return buf

    bad = []
    for j, op in enumerate(ops):
        if op["kind"] == 2:
            continue
        subs = op["micro"] if op["kind"] == 3 else [op]
        for sub in subs:
            sim(sub)
        rc = cabi.lib.tob_plan_debug_run(cp._handle, slice_id, j + 1)
        assert rc == 0, cabi.last_error()
        for sub in subs:
            n_out = 1 << (sub["m"] + sub["n"])
            want = arena[sub["c_offset"]: sub["c_offset"] + n_out]
            got = read(sub["c_offset"], n_out)
            scale = max(float(np.max(np.abs(want))), 1e-300)
            err = float(np.max(np.abs(got - want))) / scale
            if not (err <= rtol):
                bad.append((j, sub, err))
                if verbose:
                    wrong = np.nonzero(np.abs(got - want) > rtol * scale)[0]
                    print("MISMATCH op %d kind=%d m=%d n=%d k=%d ksplit=%d tm=%d tn=%d mask_m=%#x a=%s b=%s: max rel err %.3e, %d of %d outputs wrong, first at %s"
                          % (j, sub["kind"], sub["m"], sub["n"], sub["k"], sub.get("ksplit_log2", 0), sub.get("tm_log2", 0), sub.get("tn_log2", 0),
                             sub["mask_m"], sub["a"], sub["b"], err, len(wrong), n_out, wrong[:8]))
                # keep following the DEVICE's values so later ops are judged on their own
                arena[sub["c_offset"]: sub["c_offset"] + n_out] = got
    cp.close()
    return bad, len(ops)


if __name__ == "__main__":
    spec = sys.argv[1]
    slice_id = 0
    for a in sys.argv[2:]:
        if "=" in a:
            key, value = a.split("=")
            assert cabi.lib.tob_tuning_set(key.encode(), float(value)) == 0, cabi.last_error()
        else:
            slice_id = int(a)
    name, _, variant = spec.partition(":")
    pp = PortablePlan.load(os.path.join(REPO, "tests", "golden", name + ".json.gz"))
    if variant:
        pp = pp.variant(variant)
    bad, n = verify(pp, slice_id)
    print("%s slice %d: %d ops checked, %d mismatching joins" % (spec, slice_id, n, len(bad)))
