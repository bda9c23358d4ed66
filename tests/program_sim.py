"""Test-side numpy interpreter of the compiled program (`tob_plan_describe`).  It executes exactly
the canonical-form semantics the CUDA kernels implement, against one shared arena array, so the
host-side compiler (layouts, operand swap, hoisting, arena liveness, slice offsets) is checked on
CPU.  Test infrastructure only."""
import numpy as np


def pdep(x, mask):
    out = 0
    i = 0
    p = 0
    while mask >> p:
        if (mask >> p) & 1:
            out |= ((x >> i) & 1) << p
            i += 1
        p += 1
    return out


def pdep_table(nbits, mask):
    idx = np.arange(1 << nbits, dtype=np.int64)
    out = np.zeros_like(idx)
    i = 0
    p = 0
    while mask >> p:
        if (mask >> p) & 1:
            out |= ((idx >> i) & 1) << p
            i += 1
        p += 1
    return out


def upload_leaves(desc, flat):
    dev = np.zeros(max(desc["leaf_doubles"], 1), dtype=np.float64)
    for L in desc["leaves"]:
        n = 1 << L["rank"]
        d = np.arange(n, dtype=np.int64)
        s = np.zeros(n, dtype=np.int64)
        for q, sb in enumerate(L["src_bit"]):
            s |= ((d >> q) & 1) << sb
        dev[L["dev_offset"]: L["dev_offset"] + n] = flat.leaf_data[L["src_offset"] + s]
    return dev


def run_program(desc, flat, first=0, count=None, stride=1, modulus=0):
    """modulus > 0: the exact mode — every op reduces modulo the prime (done here in int64, exact)."""
    S = desc["n_slice_groups"]
    if count is None:
        count = ((1 << S) - first + stride - 1) // stride
    leaves = upload_leaves(desc, flat)
    arena = np.full(max(desc["arena_doubles"], 1), np.nan, dtype=np.float64)
    leaf_off = [0] * len(desc["leaves"])

    def operand(ref, size):
        if ref["space"] == 0:
            off = ref["offset"] + (leaf_off[ref["leaf"]] if ref["leaf"] >= 0 else 0)
            return leaves[off: off + size]
        if ref["space"] == 2:  # slice-invariant tensor: lane 0's arena (the simulator has one arena)
            assert ref["node"] in invariant_nodes, "space 2 must name a hoisted result"
        return arena[ref["offset"]: ref["offset"] + size]

    def do(op):
        nonlocal acc
        if op["kind"] == 2:
            acc += operand(op["a"], 1)[0]
            if modulus:
                acc %= modulus
            return
        if op["kind"] == 3:  # micro-subtree launch: every join of every CTA, in CTA order
            assert op["cta_start"][0] == 0 and op["cta_start"][-1] == len(op["micro"])
            for sub in op["micro"]:
                assert sub["m"] + sub["n"] <= 12 and sub["m"] + sub["k"] <= 12 and sub["n"] + sub["k"] <= 12
                do(sub)
            return
        m, n, k = op["m"], op["n"], op["k"]
        A = operand(op["a"], 1 << (m + k)).reshape(1 << m, 1 << k)
        B = operand(op["b"], 1 << (n + k)).reshape(1 << n, 1 << k)
        assert not np.isnan(A).any() and not np.isnan(B).any(), "operand read before written / after freed"
        if modulus:
            Cm = ((A.astype(np.int64) @ B.T.astype(np.int64)) % modulus).astype(np.float64)
        else:
            Cm = A @ B.T
        mask_m = op["mask_m"]
        mask_n = ~mask_m & ((1 << (m + n)) - 1)
        addr = pdep_table(m, mask_m)[:, None] | pdep_table(n, mask_n)[None, :]
        out = arena[op["c_offset"]: op["c_offset"] + (1 << (m + n))]
        out[addr.reshape(-1)] = Cm.reshape(-1)

    acc = 0.0
    invariant_nodes = set()
    for op in desc["invariant_ops"]:
        invariant_nodes.add(op.get("node"))
        for sub in op.get("micro", []):
            invariant_nodes.add(sub["node"])
    for op in desc["invariant_ops"]:
        do(op)
    sid = first
    for _ in range(count):
        for l, L in enumerate(desc["leaves"]):
            off = 0
            for ib, ab in zip(L["slice_id_bit"], L["slice_addr_bit"]):
                off |= ((sid >> ib) & 1) << ab
            leaf_off[l] = off
        for op in desc["slice_ops"]:
            do(op)
        sid += stride
    return acc
