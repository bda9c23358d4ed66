"""The REAL reference, built into the git-ignored `oracle/_ref/`.  TEST INFRASTRUCTURE ONLY.

vardigroup/TensorOrder is a Python + Cython package, so "the reference compiled here" is: its `src/`
tree with the five Cython extensions built in place by the reference's own `setup.py build_ext
--inplace`, FlowCutter (the planner binary of `line-Flow` / `factor-Flow`) built with one g++ line, and
the one CNF it ships.  `build()` makes that copy from `/root/reference` (read-only mount, build
container only) into `oracle/_ref/`; the directory is listed in `.gitignore` (no reference source enters
the history) but not in `.gpurunignore`, so it travels to the GPU box exactly like `libtob200.so`.
Nothing here reads `/root/reference` at run time: on the GPU box only the prebuilt `oracle/_ref/` is used.

Who may use this module: `tests/`, `bench.py --impl reference` / `cpu_baseline`, `__graft_entry__`
(build + smoke) and the fixture generators under `tests/golden/`.  Nothing under `tensororder_b200/`
imports it.

Two environment shims, no source edits:
  * `numpy.object = object` — `src/tensor_network/tensor_apis/numpy_apis.py:21` evaluates the alias numpy
    >= 1.24 removed on every `add_argument("entry_type", ...)` call;
  * FlowCutter needs `-include string` (`solvers/flow-cutter-pace17/src/list_graph.h:19`)."""
import os
import shutil
import subprocess
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DIR = os.environ.get("TENSORORDER_REF_BUILD", os.path.join(HERE, "_ref"))
_SUBDIRS = ("src", "solvers/flow-cutter-pace17", "benchmarks")
_EXT_MODULES = ("tensor_network/tensor_network", "contraction_methods/contraction_tree",
                "contraction_methods/factor_tree_method", "decompositions/tree_decomposition",
                "decompositions/branch_decomposition")


def _ext_built(mod):
    d, base = os.path.split(os.path.join(REF_DIR, "src", mod))
    return os.path.isdir(d) and any(f.startswith(base + ".") and f.endswith(".so") for f in os.listdir(d))


def available() -> bool:
    """True when a built copy of the reference sits in oracle/_ref/ (here or shipped to the GPU box)."""
    return all(_ext_built(m) for m in _EXT_MODULES)


def flow_cutter_path():
    return os.path.join(REF_DIR, "solvers", "flow-cutter-pace17", "flow_cutter_pace17")


def build(force: bool = False) -> bool:
    """Copy + build the reference into oracle/_ref/ when /root/reference is present; returns available()."""
    if not os.path.isdir(REF_SRC):
        return available()
    if available() and os.path.exists(flow_cutter_path()) and not force:
        return True
    os.makedirs(REF_DIR, exist_ok=True)
    for sub in _SUBDIRS:
        dst = os.path.join(REF_DIR, sub)
        if force and os.path.exists(dst):
            shutil.rmtree(dst)
        if not os.path.exists(dst):
            shutil.copytree(os.path.join(REF_SRC, sub), dst)
    subprocess.check_call(["chmod", "-R", "u+w", REF_DIR])
    src = os.path.join(REF_DIR, "src")
    res = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=src, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout[-4000:] + res.stderr[-4000:])
        raise RuntimeError("building the reference's Cython modules failed")
    shutil.rmtree(os.path.join(src, "build"), ignore_errors=True)  # object files: not needed at run time
    for m in _EXT_MODULES:  # Cython's generated C++ (megabytes): not needed at run time either
        if os.path.exists(os.path.join(src, m + ".cpp")):
            os.remove(os.path.join(src, m + ".cpp"))
    fc = os.path.join(REF_DIR, "solvers", "flow-cutter-pace17")
    subprocess.check_call("g++ -w -include string -std=c++11 -O3 -DNDEBUG src/*.cpp -o flow_cutter_pace17", shell=True, cwd=fc)
    return available()


def import_reference(chdir: bool = False):
    """Imports the reference's top-level packages from oracle/_ref/src.  chdir=True additionally enters
    oracle/_ref (the planners resolve `solvers/...` relative to the cwd, src/util/util.py:280-296)."""
    if not available():
        raise RuntimeError("oracle/_ref is not built (run `python -m oracle.reference` where /root/reference exists)")
    warnings.filterwarnings("ignore")
    import numpy

    if not hasattr(numpy, "object"):
        numpy.object = object
    src = os.path.join(REF_DIR, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    if chdir:
        os.chdir(REF_DIR)
    sys.setrecursionlimit(100000)
    import contraction_methods  # noqa
    import planning  # noqa
    import tensor_network  # noqa
    import util  # noqa
    from tensor_network import sliced_execution_plan
    from util import boolean_formula

    util.set_verbosity(0)
    return dict(tensor_network=tensor_network, contraction_methods=contraction_methods, planning=planning,
                util=util, sliced_execution_plan=sliced_execution_plan,
                WeightFormat=boolean_formula.WeightFormat)


def to_reference_plan(R, pp, as_int: bool = False):
    """Reference objects from a stored portable plan: `TensorNetwork` with `BuiltTensor` leaves (edges
    connected in edge-id order so the ids match), `ContractionTreeContext.leaf/join` along the stored
    post-order, `SlicedExecutionPlan` carrying the stored `groups_to_slice`.  The result is what
    `execution.py:95-96` obtains from a `.con` pickle."""
    import numpy as np
    from contraction_methods.contraction_tree import ContractionTreeContext
    from tensor_network.tensor import BuiltTensor
    from tensor_network.tensor_network import TensorNetwork

    net = TensorNetwork()
    for t in pp.tensors:
        arr = np.array(t["data"], dtype=np.float64).reshape(t["shape"])
        if as_int:  # exact replays: Python ints in an object array, like the reference's own bigint leaves
            assert np.array_equal(arr, np.rint(arr))
            arr = np.array([int(x) for x in arr.reshape(-1)], dtype=object).reshape(t["shape"])
        net.add_node(BuiltTensor(arr))
    for e, (t1, t2) in enumerate(pp.edges):
        got = net.connect(t1, pp.index_lists[t1].index(e), t2, pp.index_lists[t2].index(e))
        assert got == e
    for t, il in enumerate(pp.index_lists):
        assert list(net.index_list(t)) == il
    ctx = ContractionTreeContext()
    ids = []
    for node in pp.postorder:
        ids.append(ctx.leaf(net, node[0]) if len(node) == 1 else ctx.join(ids[node[0]], ids[node[1]]))
    tree = ctx.get_tree(ids[-1])
    plan = R["sliced_execution_plan"].SlicedExecutionPlan(tree, net)
    plan.groups_to_slice = [set(g) for g in pp.groups_to_slice]
    plan.edges_to_slice = set().union(*plan.groups_to_slice) if plan.groups_to_slice else set()
    return plan


def write_con(R, pp, path):
    """A `.con` file as `planning.py --store` writes it (planning.py:108-111): pickle of
    (elapsed, tree, network); `execution.py` reads it from stdin (execution.py:95)."""
    import pickle

    plan = to_reference_plan(R, pp)
    with open(path, "wb") as f:
        pickle.dump((0.0, plan.tree, plan.network), f)
    return path


if __name__ == "__main__":
    print("oracle/_ref built:", build(force="--force" in sys.argv), "at", REF_DIR)
