#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference).  Nothing here is used at test
time: the tests read the committed outputs.  What it does:

1. builds the reference into the git-ignored oracle/_ref/ (oracle/reference.py: the five Cython
   extensions with the reference's own `setup.py build_ext --inplace`, FlowCutter with g++,
   SURVEY.md Appendix A);
2. (re)generates the cubic vertex-cover CNF family the reference ships one member of
   (`benchmarks/cubic_vertex_cover/cubic_vc_50_0.cnf`): networkx.random_regular_graph(3, n, seed)
   -> one clause "u v 0" per edge;
3. for every job: reference `cnf_count` -> reference planner (`planning.run`, fixed timeout) ->
   reference slicer -> reference numpy backend `NumpyAPI.contract_sliced` (float64) and, for
   small networks, `contract_einsum`; the plan is exported with
   `tensororder_b200.plan_format.export_plan` and the reference's results are stored in
   `expected`.

Usage:  python tests/golden/make_golden.py [--only PATTERN] [--max-n N] [--force]
"""
import argparse
import fnmatch
import io
import itertools
import json
import os
import random
import shutil
import subprocess
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference"
sys.path.insert(0, REPO)


def ensure_reference_build():
    """The reference is built once into the git-ignored oracle/_ref/ (oracle/reference.py)."""
    from oracle import reference
    if not reference.build():
        raise RuntimeError("the reference could not be built into " + reference.REF_DIR)


def import_reference():
    from oracle import reference
    return reference.import_reference(chdir=True)  # the planners resolve solvers/... relative to the cwd


# --------------------------------------------------------------------------------------
# CNF family
# --------------------------------------------------------------------------------------
def cubic_vc_cnf(n, seed):
    import networkx
    g = networkx.random_regular_graph(3, n, seed=seed)
    edges = sorted((min(u, v) + 1, max(u, v) + 1) for u, v in g.edges())
    lines = ["p cnf %d %d" % (n, len(edges))] + ["%d %d 0" % e for e in edges]
    return "\n".join(lines) + "\n"


def random_kcnf(n_vars, n_clauses, k, seed):
    """Uniform random k-CNF (distinct variables per clause, random signs): leaves of higher rank than the
    cubic family has (a variable tensor's rank is its number of occurrences)."""
    import numpy
    rng = numpy.random.default_rng(seed)
    lines = ["p cnf %d %d" % (n_vars, n_clauses)]
    for _ in range(n_clauses):
        vs = rng.choice(n_vars, size=k, replace=False) + 1
        signs = rng.integers(0, 2, size=k) * 2 - 1
        lines.append(" ".join(str(int(v * s)) for v, s in zip(vs, signs)) + " 0")
    return "\n".join(lines) + "\n"


def mcc_weight_lines(n, rng_seed):
    import numpy
    rng = numpy.random.default_rng(rng_seed)
    lines = []
    for v in range(1, n + 1):
        for lit in (v, -v):
            lines.append("w %d %r" % (lit, float(rng.uniform(0.5, 1.5))))
    return "\n".join(lines) + "\n"


def cachet_weight_lines(n):
    return "".join("w %d %r\n" % (i, ((37 * i) % 89 + 5) / 100) for i in range(1, n + 1))


TOY_CNFS = {
    # unit clause, negative literals, a variable that occurs nowhere (rank-0 variable tensor)
    "toy_unit_neg": "p cnf 5 4\n1 -2 0\n-1 3 0\n2 -3 4 0\n-4 0\n",
    # 3-literal clauses, every variable in >= 4 clauses so factor-* really factors
    "toy_3cnf": "p cnf 6 8\n1 2 3 0\n-1 2 4 0\n1 -3 5 0\n2 -4 6 0\n-2 3 -5 0\n1 4 -6 0\n-1 -2 6 0\n3 4 5 0\n",
    # duplicate-free 2-CNF path
    "toy_path": "p cnf 6 5\n1 2 0\n2 3 0\n3 4 0\n4 5 0\n5 6 0\n",
}


# --------------------------------------------------------------------------------------
# One job
# --------------------------------------------------------------------------------------
def plan_with_reference(R, cnf_text, weights, planner_name, seed, timeout):
    """reference reduction + planning (`tensororder.py:231-252` without the performance-factor stop)."""
    random.seed(seed)
    network = R["tensor_network"].ALL_CONSTRUCTIONS["wmc"](io.StringIO(cnf_text), R["WeightFormat"][weights])
    planner = R["contraction_methods"].ALL_SOLVERS[planner_name]
    stopwatch = R["util"].Stopwatch()
    with R["util"].TimeoutTimer(timeout) as timer:
        best_plan, log = R["planning"].run(planner, network, seed, timer, None, rank_limit=None,
                                           performance_factor=None, mem_limit=None, slicer=None,
                                           stopwatch=stopwatch)
    improving = []
    best = None
    for elapsed, plan in log:
        if best is None or plan.tree.maxrank < best:
            best = plan.tree.maxrank
            improving.append(plan)
    return best_plan, improving


def reference_contract(R, plan, num_slice_limit=None, want_per_slice=False):
    api = R["tensor_network"].ALL_APIS["numpy"]()
    api.add_argument("entry_type", "float64")
    t0 = time.perf_counter()
    result = api.contract_sliced(plan, num_slice_limit)
    dt = time.perf_counter() - t0
    per_slice = None
    if want_per_slice:
        per_slice = []
        slices = plan.network.slice_groups(plan.groups_to_slice)
        if num_slice_limit is not None:
            slices = itertools.islice(slices, num_slice_limit)
        for net in slices:
            per_slice.append(float(api.contract(net, plan.tree)[tuple()]))
    return float(result), dt, per_slice


def expected_block(R, plan, count, seconds, per_slice=None, extra=None):
    exp = {
        "count": count,
        "count_hex": float(count).hex(),
        "entry_type": "float64",
        "num_slices": 2 ** len(plan.groups_to_slice),
        "maxrank": int(plan.maxrank),
        "estimated_memory": float(plan.memory),
        "estimated_flops": float(plan.total_FLOPs),
        "numpy_seconds_buildbox_8c": seconds,
    }
    if per_slice is not None:
        exp["per_slice"] = per_slice
    if extra:
        exp.update(extra)
    return exp


def fresh_plan(R, base_plan):
    return R["sliced_execution_plan"].SlicedExecutionPlan(base_plan.tree, base_plan.network)


def add_params(R, pp, network):
    """Record constructor parameters of the reference leaf classes (for the oracle's restated builders)."""
    for doc, t in zip(pp.tensors, network.tensors):
        name = type(t).__name__
        if name == "OrTensor":
            doc["params"] = {"literals_positive": [bool(x) for x in t._OrTensor__literals_positive],
                             "output_index": t._OrTensor__output_index}
        elif name == "VariableTensor":
            doc["params"] = {"rank": t.rank, "positive_weight": float(t._VariableTensor__positive_weight),
                             "negative_weight": float(t._VariableTensor__negative_weight)}


def make_fixture(R, name, cnf_text, weights, planner, seed, timeout, slicings, contract=True, tree_check=True,
                 which_tree="best", einsum=False, meta_extra=None):
    from tensororder_b200.plan_format import export_plan
    best_plan, improving = plan_with_reference(R, cnf_text, weights, planner, seed, timeout)
    if best_plan is None:
        raise RuntimeError("no plan for " + name)
    base = improving[0] if which_tree == "first" else best_plan
    meta = {"cnf_chars": len(cnf_text), "weights": weights, "planner": planner, "seed": seed,
            "planner_timeout": timeout, "which_tree": which_tree, "generator": "tests/golden/make_golden.py",
            "reference": "vardigroup/TensorOrder numpy backend, float64"}
    meta.update(meta_extra or {})
    plan = fresh_plan(R, base)
    if contract and plan.memory > 3e8:
        raise RuntimeError("%s: best plan needs %.3g entries; refusing to contract it with numpy" % (name, plan.memory))
    pp = export_plan(plan, name, meta=meta, with_tree_check=tree_check)
    add_params(R, pp, plan.network)
    if contract:
        count, dt, per = reference_contract(R, plan, want_per_slice=False)
        extra = {}
        if einsum:
            api = R["tensor_network"].ALL_APIS["numpy"]()
            api.add_argument("entry_type", "float64")
            try:
                extra["einsum"] = float(plan.network.contract_einsum(api))
            except ValueError:  # more than 26 indices (factored networks grow)
                pass
        pp.expected = expected_block(R, plan, count, dt, extra=extra)
    else:
        pp.expected = {"num_slices": 1, "maxrank": int(plan.maxrank), "estimated_memory": float(plan.memory),
                       "estimated_flops": float(plan.total_FLOPs), "entry_type": "float64"}
    for sl in slicings:
        p = fresh_plan(R, base)
        random.seed(seed)
        slicer = R["tensor_network"].ALL_SLICERS[sl.get("slicer", "greedy_mem")]
        mem = sl.get("mem_limit_bytes")
        slicer.slice_until(p, memory=(mem / 8 if mem is not None else None), rank=sl.get("rank_limit"),
                           slices=sl.get("minimum_slice"))
        var = {"name": sl["name"], "groups_to_slice": [sorted(int(e) for e in g) for g in p.groups_to_slice],
               "options": {k: v for k, v in sl.items() if k not in ("name", "contract")}}
        if sl.get("contract", True):
            count, dt, per = reference_contract(R, p, want_per_slice=(2 ** len(p.groups_to_slice) <= 128))
            var["expected"] = expected_block(R, p, count, dt, per_slice=per)
            cutoff = sl.get("slice_cutoff")
            if cutoff is not None:
                c2, _, _ = reference_contract(R, p, num_slice_limit=cutoff)
                var["expected"]["slice_cutoff"] = cutoff
                var["expected"]["count_cutoff"] = c2
        else:
            var["expected"] = {"num_slices": 2 ** len(p.groups_to_slice), "maxrank": int(p.maxrank),
                               "estimated_memory": float(p.memory), "estimated_flops": float(p.total_FLOPs),
                               "entry_type": "float64"}
        pp.variants.append(var)
    return pp


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="*")
    ap.add_argument("--max-n", type=int, default=200, help="largest family member to CONTRACT with the reference")
    ap.add_argument("--max-plan-n", type=int, default=250, help="largest family member to plan")
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()

    ensure_reference_build()
    R = import_reference()

    cnf_dir = os.path.join(HERE, "cnf")
    os.makedirs(cnf_dir, exist_ok=True)
    shipped = open(os.path.join(REF_SRC, "benchmarks/cubic_vertex_cover/cubic_vc_50_0.cnf")).read()

    def family_cnf(n, seed=0):
        path = os.path.join(cnf_dir, "cubic_vc_%d_%d.cnf" % (n, seed))
        if n == 50 and seed == 0:
            text = shipped  # the one member the reference ships
        elif os.path.exists(path):
            return open(path).read()
        else:
            text = cubic_vc_cnf(n, seed)
        with open(path, "w") as f:
            f.write(text)
        return text

    jobs = []

    def job(name, **kw):
        jobs.append((name, kw))

    std_slicings = [
        {"name": "min3", "minimum_slice": 3, "slice_cutoff": 3},
        {"name": "rank8", "rank_limit": 8},
        {"name": "mem20000", "mem_limit_bytes": 20000},
        {"name": "most_min2", "slicer": "greedy_most", "minimum_slice": 2},
    ]
    # --- config 1: the shipped instance, README command ---
    job("vc50_lineflow", cnf=family_cnf(50), weights="unweighted", planner="line-Flow", seed=1, timeout=3,
        slicings=std_slicings)
    job("vc50_lineflow_first", cnf=family_cnf(50), weights="unweighted", planner="line-Flow", seed=1, timeout=3,
        slicings=[{"name": "min4", "minimum_slice": 4}], which_tree="first")
    job("vc50_factorflow", cnf=family_cnf(50), weights="unweighted", planner="factor-Flow", seed=1, timeout=3,
        slicings=[{"name": "min2", "minimum_slice": 2}, {"name": "most_min2", "slicer": "greedy_most", "minimum_slice": 2}])
    job("vc50_mcc_lineflow", cnf=family_cnf(50) + mcc_weight_lines(50, 50), weights="mcc", planner="line-Flow",
        seed=1, timeout=3, slicings=[{"name": "min3", "minimum_slice": 3}])
    job("vc50_mcc_factorflow", cnf=family_cnf(50) + mcc_weight_lines(50, 50), weights="mcc", planner="factor-Flow",
        seed=1, timeout=3, slicings=[{"name": "min3", "minimum_slice": 3}])
    job("vc50_cachet_lineflow", cnf=family_cnf(50) + cachet_weight_lines(50), weights="cachet", planner="line-Flow",
        seed=1, timeout=3, slicings=[{"name": "min2", "minimum_slice": 2}])
    # --- toy networks with the einsum cross-check ---
    for tname, text in TOY_CNFS.items():
        for planner in ("line-Flow", "factor-Flow"):
            job("%s_%s" % (tname, planner.replace("-", "").lower()), cnf=text, weights="unweighted", planner=planner,
                seed=1, timeout=2, slicings=[{"name": "min1", "minimum_slice": 1}, {"name": "min2", "minimum_slice": 2}],
                einsum=True)
    job("toy_3cnf_mcc", cnf=TOY_CNFS["toy_3cnf"] + mcc_weight_lines(6, 6), weights="mcc", planner="factor-Flow",
        seed=1, timeout=2, slicings=[{"name": "min2", "minimum_slice": 2}], einsum=True)
    # --- other CNF shapes: random 3-CNF / 4-CNF (variable tensors of rank ~6-10, clause tensors of rank 3-4) ---
    job("rand3cnf_24_lineflow", cnf=random_kcnf(24, 40, 3, 30), weights="unweighted", planner="line-Flow", seed=1,
        timeout=3, slicings=[{"name": "min3", "minimum_slice": 3}], tree_check=False)
    job("rand3cnf_24_factorflow", cnf=random_kcnf(24, 40, 3, 30), weights="unweighted", planner="factor-Flow", seed=1,
        timeout=3, slicings=[{"name": "min3", "minimum_slice": 3}], tree_check=False)
    job("rand3cnf_16_lineflow", cnf=random_kcnf(16, 30, 3, 5), weights="unweighted", planner="line-Flow", seed=1,
        timeout=3, slicings=[{"name": "min4", "minimum_slice": 4}], tree_check=False)
    job("rand4cnf_18_mcc_lineflow", cnf=random_kcnf(18, 24, 4, 24) + mcc_weight_lines(18, 24), weights="mcc",
        planner="line-Flow", seed=1, timeout=3, slicings=[{"name": "min2", "minimum_slice": 2}], tree_check=False)
    # --- config 2: the family ---
    for n in range(60, args.max_plan_n + 1, 10):
        timeout = 5 if n <= 100 else (10 if n <= 150 else 20)
        slicings = []
        if n in (100, 150):
            slicings = [{"name": "min4", "minimum_slice": 4}]
        if n >= 200:
            slicings = [{"name": "min3", "minimum_slice": 3, "contract": False},
                        {"name": "min6", "minimum_slice": 6, "contract": False}]
        job("vc%d_lineflow" % n, cnf=family_cnf(n), weights="unweighted", planner="line-Flow", seed=1,
            timeout=timeout, slicings=slicings, contract=(n <= args.max_n), tree_check=(n <= 100))
    # --- config 3: weighted factor-Flow ---
    for n in (100, 150):
        job("vc%d_mcc_factorflow" % n, cnf=family_cnf(n) + mcc_weight_lines(n, n), weights="mcc",
            planner="factor-Flow", seed=1, timeout=10, slicings=[{"name": "min3", "minimum_slice": 3}],
            tree_check=False)

    for name, kw in jobs:
        if not fnmatch.fnmatch(name, args.only):
            continue
        out = os.path.join(HERE, name + ".json.gz")
        if os.path.exists(out) and not args.force:
            print("keep", name)
            continue
        t0 = time.time()
        pp = make_fixture(R, name, kw["cnf"], kw["weights"], kw["planner"], kw["seed"], kw["timeout"],
                          kw.get("slicings", []), contract=kw.get("contract", True),
                          tree_check=kw.get("tree_check", True), which_tree=kw.get("which_tree", "best"),
                          einsum=kw.get("einsum", False))
        pp.save(out)
        print("wrote %-28s maxrank=%-3s count=%s  [%0.1fs]" % (name, pp.expected.get("maxrank"),
                                                           pp.expected.get("count"), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
