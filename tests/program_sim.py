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


def dag_order(ops, rng):
    """A random order of one op list that respects ONLY what the executor enforces (tob_exec.cu run_list):
    list order inside a branch, the declared cross-branch waits, and the barrier ops (micro launch,
    accumulate) that join every branch.  If the compiler's dependency analysis misses a hazard, some
    seed reorders the two ops and the result changes."""
    n = len(ops)
    preds = [set() for _ in range(n)]
    last_on_branch = {}
    barrier = None
    since_barrier = []
    for j, op in enumerate(ops):
        if op["kind"] in (2, 3):  # barrier
            preds[j].update(since_barrier)
            if barrier is not None:
                preds[j].add(barrier)
            barrier, since_barrier, last_on_branch = j, [], {}
            continue
        if barrier is not None:
            preds[j].add(barrier)
        b = op.get("branch", 0)
        if b in last_on_branch:
            preds[j].add(last_on_branch[b])
        last_on_branch[b] = j
        for w in op.get("waits", []):
            assert w < j and ops[w].get("signal") == 1 and ops[w].get("branch", 0) != b
            preds[j].add(w)
        since_barrier.append(j)
    done, order = set(), []
    ready = [j for j in range(n) if not preds[j]]
    while ready:
        j = ready.pop(int(rng.integers(len(ready))))
        order.append(j)
        done.add(j)
        for q in range(n):
            if q not in done and q not in ready and preds[q] <= done:
                ready.append(q)
    assert len(order) == n
    return order


def run_program(desc, flat, first=0, count=None, stride=1, modulus=0, dag_seed=None):
    """modulus > 0: the exact mode — every op reduces modulo the prime (done here in int64, exact).
    dag_seed: execute every op list in a random order allowed by its DAG schedule instead of list order."""
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
                assert sub["m"] + sub["n"] <= 14 and sub["m"] + sub["k"] <= 14 and sub["n"] + sub["k"] <= 14
                assert sub["m"] + sub["n"] + sub["k"] <= 17 and sub["k"] <= 4 and (op["threads"] == 1024) == any(
                    q["m"] + q["n"] > 10 for q in op["micro"])
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
    rng = np.random.default_rng(dag_seed) if dag_seed is not None else None

    def run_list(ops):
        for j in (dag_order(ops, rng) if rng is not None else range(len(ops))):
            do(ops[j])

    run_list(desc["invariant_ops"])
    sid = first
    for _ in range(count):
        for l, L in enumerate(desc["leaves"]):
            off = 0
            for ib, ab in zip(L["slice_id_bit"], L["slice_addr_bit"]):
                off |= ((sid >> ib) & 1) << ab
            leaf_off[l] = off
        run_list(desc["slice_ops"])
        sid += stride
    return acc


def install_fake_device(compiled_plan_cls, setattr_fn=setattr, record=None):
    """CPU tests of the host logic: replaces the device stage of `CompiledPlan` (leaf upload, modulus, run)
    by this interpreter of the compiled program.  `record` (a dict) receives the slice range each run was given."""
    import dataclasses

    import numpy as np

    state = {}

    def fake_update_leaves(self, leaf_data=None):
        if leaf_data is not None:
            self.flat = dataclasses.replace(self.flat, leaf_data=np.ascontiguousarray(leaf_data, dtype=np.float64))
        self.uploaded = True

    def fake_upload(self):
        self.uploaded = True

    def fake_release(self):
        self.uploaded = False

    def fake_set_modulus(self, modulus):
        state["modulus"] = int(modulus)

    def fake_run(self, first=0, count=None, stride=1, initial=0.0, skip_invariant=False):
        if record is not None:
            a = record.setdefault("args", [first, 0, stride])
            a[1] += count
        m = state.get("modulus", 0)
        r = run_program(self.describe(), self.flat, first=first, count=count, stride=stride, modulus=m) if count else 0.0
        return (initial + r) % m if m else initial + r

    setattr_fn(compiled_plan_cls, "update_leaves", fake_update_leaves)
    setattr_fn(compiled_plan_cls, "upload", fake_upload)
    setattr_fn(compiled_plan_cls, "release_device", fake_release)
    setattr_fn(compiled_plan_cls, "set_modulus", fake_set_modulus)
    setattr_fn(compiled_plan_cls, "run", fake_run)
    setattr_fn(compiled_plan_cls, "last_ms", 0.0)
    setattr_fn(compiled_plan_cls, "last_launches", 0)
    setattr_fn(compiled_plan_cls, "last_gemm", (0.0, 0.0, 0))
    return state
