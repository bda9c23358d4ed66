"""Drop-in checks against the REAL reference, built into the git-ignored oracle/_ref/ (oracle/reference.py),
which travels to the GPU box.  Registers "b200" in the reference's live registry, resolves it through the
reference's own click option type, and drives `execution.run` and both unmodified CLIs
(`src/execution.py`, `src/tensororder.py`) with the backend.

Without a GPU (`-m "not gpu"`) the backend must fail loudly (no CPU fallback) and the reference must
report it the way it reports any backend failure (execution.py:143-152); with one (`-m gpu`) the counts
must equal the reference's own numpy run of the same plan and `Contraction Time:` must be reported."""
import os
import re
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import reference  # noqa: E402

pytestmark = pytest.mark.skipif(not reference.available(), reason="oracle/_ref is not built (needs /root/reference once)")
REF_DIR = reference.REF_DIR


@pytest.fixture(scope="module")
def ref():
    return reference.import_reference()


def _has_gpu():
    from tensororder_b200 import cabi

    return cabi.lib.tob_device_count() > 0


def _cli(script, args, stdin_bytes, library="b200", timeout=600, torchrun=0):
    """Runs an unmodified reference CLI through the launcher, from oracle/_ref (the planners resolve
    `solvers/...` relative to the cwd); returns (stdout, stderr)."""
    env = dict(os.environ, PYTHONPATH=REPO)
    cmd = [sys.executable]
    if torchrun:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(torchrun),
                "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300)]
    cmd += ["-m", "tensororder_b200.launch", os.path.join(REF_DIR, "src", script), "--tensor_library=" + library] + args
    res = subprocess.run(cmd, input=stdin_bytes, capture_output=True, env=env, cwd=REF_DIR, timeout=timeout)
    return res.stdout.decode(), res.stderr.decode(), res.returncode


def _field(out, key):
    m = re.search(r"^%s: (.*)$" % re.escape(key), out, re.M)
    return m.group(1).strip() if m else None


def _con_bytes(ref, name, tmp_path):
    from conftest import load_golden

    path = str(tmp_path / (name.replace(":", "_") + ".con"))
    base, _, variant = name.partition(":")
    pp = load_golden(base)
    reference.write_con(ref, pp.variant(variant) if variant else pp, path)
    return open(path, "rb").read()


# ---------------------------------------------------------------------------------------------------
# host-side (CPU) checks
# ---------------------------------------------------------------------------------------------------
def test_registration_and_option_resolution(ref):
    from tensororder_b200.api import B200API, register

    register()
    tn = ref["tensor_network"]
    assert tn.ALL_APIS["b200"] is B200API
    choice = ref["util"].TaggedChoice(tn.ALL_APIS, case_sensitive=False)
    assert choice.convert("b200", None, None) is B200API  # what --tensor_library=b200 resolves to


def test_flatten_reads_reference_objects_like_stored_plans(ref):
    """The flattener must produce the same flat plan from live reference objects as from the stored
    copy of the same plan (the golden fixture was exported from such objects)."""
    import numpy as np

    from conftest import load_golden
    from tensororder_b200.flatten import flatten_plan, rebuild_leaf_data

    pp = load_golden("vc50_lineflow").variant("min3")
    live = reference.to_reference_plan(ref, pp)
    a = flatten_plan(live)
    b = flatten_plan(pp.as_execution_plan())
    for f in ("node_left", "node_right", "node_leaf", "leaf_rank", "leaf_data_offset", "leaf_axis_start",
              "leaf_axis_edge", "leaf_data"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.n_slice_groups == b.n_slice_groups == 3
    assert np.array_equal(rebuild_leaf_data(live, a), a.leaf_data)  # what a plan-cache hit re-reads


def test_execution_run_reaches_the_cabi(ref, capsys):
    import execution  # the reference's src/execution.py
    import tensor_network

    from conftest import load_golden
    from tensororder_b200.api import B200API

    plan = reference.to_reference_plan(ref, load_golden("vc50_lineflow"))
    api = B200API()
    api.add_argument("entry_type", "float64")
    ref["util"].set_verbosity(0)
    result = execution.run(plan, api, tensor_network.ALL_SLICERS["greedy_mem"], None)
    if _has_gpu():
        assert float(result) == 2802717837.0
    else:
        # no GPU here: the backend raises, the reference prints its generic backend-failure line
        assert result is None
        out = capsys.readouterr()
        assert "Exception during execution" in out.out
        assert "CUDA device" in out.err or "CUDA device" in out.out


def test_launcher_runs_the_unmodified_cli_help():
    out, err, rc = _cli("execution.py", ["--help"], b"")
    assert rc == 0, err
    assert "b200" in out  # listed among the --tensor_library choices
    assert "b200_mem" in out  # and the GPU-aware slicer among --slicer


def test_launcher_reference_numpy_through_con_file(ref, tmp_path):
    """The `.con` files the tests hand to `execution.py` are what `planning.py --store` writes: the
    reference's own numpy backend reads one and reproduces the stored count."""
    out, err, rc = _cli("execution.py", [], _con_bytes(ref, "vc50_lineflow", tmp_path), library="numpy")
    assert rc == 0, err
    assert float(_field(out, "Count")) == 2802717837.0
    assert float(_field(out, "Contraction Time")) > 0


def test_gpu_aware_slicer_slices_less_than_the_reference_model(ref):
    """Same plan, same byte budget: the reference cost model (out + 2*left + 2*right entries) needs more
    slices than the executor's real arena."""
    import tensor_network

    from conftest import load_golden
    from tensororder_b200.slicer import B200MemSlicer, plan_peak_bytes, register

    register()
    assert "b200_mem" in tensor_network.ALL_SLICERS
    pp = load_golden("vc250_lineflow")
    assert pp.expected["maxrank"] == 31
    budget_bytes = 40e9  # between the executor's real need (30.6 GB) and the reference model's estimate (55.8 GB)
    ref_plan = reference.to_reference_plan(ref, pp)
    assert ref_plan.memory * 8 > budget_bytes
    tensor_network.ALL_SLICERS["greedy_mem"].slice_until(ref_plan, memory=budget_bytes / 8)
    ours = reference.to_reference_plan(ref, pp)
    B200MemSlicer().slice_until(ours, memory=budget_bytes / 8)
    assert plan_peak_bytes(ours) <= budget_bytes
    assert len(ours.groups_to_slice) == 0 < len(ref_plan.groups_to_slice)
    # a budget that needs real slicing: still met, with at most as many slices as the reference model asks for
    tight = 8e9
    a = reference.to_reference_plan(ref, pp)
    tensor_network.ALL_SLICERS["greedy_mem"].slice_until(a, memory=tight / 8)
    b = reference.to_reference_plan(ref, pp)
    B200MemSlicer().slice_until(b, memory=tight / 8)
    assert plan_peak_bytes(b) <= tight and 0 < len(b.groups_to_slice) <= len(a.groups_to_slice)
    pp = load_golden("vc100_lineflow")
    with pytest.raises(RuntimeError):
        B200MemSlicer().slice_until(reference.to_reference_plan(ref, pp), memory=16)  # 128 bytes: impossible


def test_launcher_under_torchrun_shares_stdin_and_plan(ref, tmp_path):
    """torchrun's workers inherit ONE stdin and the planners are timing-dependent: the launcher reads stdin
    on rank 0 only and broadcasts it, and plans on rank 0 only.  Two gloo ranks, reference numpy
    backend (no GPU needed): both CLIs must finish and rank 0 must report the right count."""
    out, err, rc = _cli("execution.py", [], _con_bytes(ref, "vc50_lineflow:min3", tmp_path), library="numpy", torchrun=2)
    assert rc == 0, err[-2000:]
    assert out.count("Count:") == 1 and float(_field(out, "Count")) == 2802717837.0
    cnf = open(os.path.join(REF_DIR, "benchmarks", "cubic_vertex_cover", "cubic_vc_50_0.cnf"), "rb").read()
    out, err, rc = _cli("tensororder.py", ["--planner=line-Flow", "--weights=unweighted", "--seed=1", "--timeout=30"], cnf,
                        library="numpy", torchrun=2)
    assert rc == 0, err[-2000:]
    assert out.count("Count:") == 1 and float(_field(out, "Count")) == 2802717837.0


# ---------------------------------------------------------------------------------------------------
# the real reference -> b200 -> count path on the device
# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_execution_run_with_reference_objects(ref):
    """`execution.run(plan, B200API, slicer, cutoff)` (src/execution.py:122-152) on live reference objects:
    unsliced, sliced, slice_cutoff; counts equal the reference's numpy backend on the same plan objects."""
    import execution
    import tensor_network

    from conftest import load_golden
    from tensororder_b200.api import B200API

    numpy_api = tensor_network.ALL_APIS["numpy"]()
    numpy_api.add_argument("entry_type", "float64")
    slicer = tensor_network.ALL_SLICERS["greedy_mem"]
    for name, variant, cutoff in (("vc50_lineflow", None, None), ("vc50_lineflow", "min3", None),
                                  ("vc50_lineflow", "min3", 3), ("vc100_lineflow", "min4", None),
                                  ("vc50_mcc_factorflow", "min3", None)):
        pp = load_golden(name)
        pp = pp.variant(variant) if variant else pp
        plan = reference.to_reference_plan(ref, pp)
        api = B200API()
        api.add_argument("entry_type", "float64")
        got = execution.run(plan, api, slicer, cutoff)
        want = execution.run(plan, numpy_api, slicer, cutoff)
        assert got is not None and want is not None
        assert abs(float(got) - float(want)) <= 1e-9 * abs(float(want)), (name, variant, cutoff)
        if name == "vc50_lineflow" and cutoff is None:
            assert float(got) == float(want) == 2802717837.0  # representable: bit-exact


@pytest.mark.gpu
def test_gpu_oom_retry_loop_slices_until_the_plan_fits(ref):
    """The reference's recovery path (execution.py:133-142): the backend raises the reference's OWN
    OutOfMemoryError class, `run` slices once more and retries; the count is unchanged."""
    import execution
    import tensor_network

    from conftest import load_golden
    from tensororder_b200.api import B200API

    plan = reference.to_reference_plan(ref, load_golden("vc150_lineflow"))
    api = B200API()
    api.add_argument("entry_type", "float64")
    api.add_argument("mem_limit_bytes", 30000000)  # the unsliced arena needs 55 MB: forces re-slicing (4 groups)
    got = execution.run(plan, api, tensor_network.ALL_SLICERS["greedy_mem"], None)
    assert 1 <= len(plan.groups_to_slice) <= 8
    assert api.last_stats["peak_bytes"] <= 30000000
    want = load_golden("vc150_lineflow").expected["count"]
    assert abs(float(got) - want) <= 1e-9 * want
    # a limit no slicing can meet ends the reference's retry loop with its out-of-memory line instead of spinning
    plan = reference.to_reference_plan(ref, load_golden("vc50_lineflow"))
    api = B200API()
    api.add_argument("entry_type", "float64")
    api.add_argument("mem_limit_bytes", 1000)
    assert execution.run(plan, api, tensor_network.ALL_SLICERS["greedy_mem"], None) is None


@pytest.mark.gpu
def test_gpu_cli_execution_py_count_and_contraction_time(ref, tmp_path):
    """`python -m tensororder_b200.launch src/execution.py --tensor_library=b200 < N.con`: the line the
    metric is read from (`Contraction Time:`, src/util/util.py:166-174) and `Count:` next to the numpy run of
    the same `.con`."""
    for name in ("vc50_lineflow", "vc150_lineflow", "vc150_mcc_factorflow"):
        con = _con_bytes(ref, name, tmp_path)
        out_b, err_b, rc_b = _cli("execution.py", [], con, library="b200")
        out_n, err_n, rc_n = _cli("execution.py", [], con, library="numpy")
        assert rc_b == 0 and rc_n == 0, (err_b[-1500:], err_n[-1500:])
        cb, cn = float(_field(out_b, "Count")), float(_field(out_n, "Count"))
        assert abs(cb - cn) <= 1e-9 * abs(cn), (name, cb, cn)
        tb, tn = float(_field(out_b, "Contraction Time")), float(_field(out_n, "Contraction Time"))
        assert tb > 0 and tn > 0
        print("%s: Count %r | Contraction Time b200 %.4f s (first call: CUDA context + plan compile), numpy %.4f s"
              % (name, cb, tb, tn))
    # rank / memory limits and the slicers go through the reference's own option handling
    con = _con_bytes(ref, "vc100_lineflow", tmp_path)
    want = float(_field(_cli("execution.py", [], con, library="numpy")[0], "Count"))
    # (vc100 has max-rank 15: rank 13 = 4 groups / 16 slices, rank 12 = 9 groups / 512 slices; 1.9 MB = 3 groups for b200_mem)
    for args in (["--rank_limit=13"], ["--mem_limit=1900000", "--slicer=b200_mem"], ["--rank_limit=12", "--slice_cutoff=2"]):
        out_b, err_b, rc = _cli("execution.py", args, con, library="b200")
        assert rc == 0, err_b[-1500:]
        out_n, _, _ = _cli("execution.py", [a for a in args if "b200_mem" not in a], con, library="numpy")
        if "--slice_cutoff=2" in args:
            assert abs(float(_field(out_b, "Count")) - float(_field(out_n, "Count"))) <= 1e-9 * abs(float(_field(out_n, "Count")))
        else:
            assert abs(float(_field(out_b, "Count")) - want) <= 1e-9 * want, args
        assert int(_field(out_b, "# Network Slices")) >= 2


@pytest.mark.gpu
def test_gpu_cli_tensororder_py_end_to_end():
    """The README command (README.md:21) with `--tensor_library=b200`: reduction, planning (FlowCutter
    subprocess), slicing, `--early` and execution all inside the unmodified `src/tensororder.py`."""
    cnf = open(os.path.join(REF_DIR, "benchmarks", "cubic_vertex_cover", "cubic_vc_50_0.cnf"), "rb").read()
    base = ["--planner=line-Flow", "--weights=unweighted", "--seed=1", "--timeout=60", "--verbosity=2"]
    for extra in ([], ["--minimum_slice=3"], ["--early=6"], ["--early=6", "--minimum_slice=2"], ["--entry_type=bigint"],
                  ["--entry_type=int"], ["--entry_type=float32"], ["--planner=factor-Flow", "--slicer=b200_mem"]):
        out, err, rc = _cli("tensororder.py", base + extra, cnf, library="b200")
        assert rc == 0, err[-1500:]
        count = _field(out, "Count")
        assert count is not None, (extra, out[-800:], err[-800:])
        assert float(count) == pytest.approx(2802717837.0, rel=(1e-6 if "--entry_type=float32" in extra else 0)), (extra, count)
        assert float(_field(out, "Contraction Time")) > 0


@pytest.mark.gpu
def test_gpu_b200_mem_slicer_contracts_within_the_budget(ref):
    """slice with `B200MemSlicer` until the EXECUTOR's arena fits the byte budget, contract
    on the device, same count as unsliced and `peak_bytes <= budget` (cf. GreedyMemSlicer, slicers.py:27-33)."""
    from conftest import load_golden
    from tensororder_b200.api import B200API, CompiledPlan
    from tensororder_b200.flatten import flatten_plan
    from tensororder_b200.slicer import B200MemSlicer

    for name, budget in (("vc150_lineflow", 2.5e7), ("vc200_lineflow", 5e8)):
        pp = load_golden(name)
        want = pp.expected["count"]
        plan = reference.to_reference_plan(ref, pp)
        unsliced_peak = CompiledPlan(flatten_plan(plan)).peak_bytes  # host-only compile
        assert unsliced_peak > budget
        B200MemSlicer().slice_until(plan, memory=budget / 8)
        assert len(plan.groups_to_slice) >= 1
        api = B200API()
        api.add_argument("entry_type", "float64")
        api.add_argument("mem_limit_bytes", int(budget))
        got = float(api.contract_sliced(plan))
        assert api.last_stats["peak_bytes"] <= budget
        assert abs(got - want) <= 1e-9 * want, (name, got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("entry_type", ["int", "uint", "float32", "float16", "bigint"])
def test_gpu_entry_types_match_the_reference_numpy_backend(ref, entry_type):
    """The reference's dtype table (numpy_apis.py:15-22) on the same live plan objects: int / uint wrap modulo 2^64
    exactly like numpy's integer tensordot (n=150 has 2.3e28 covers: several wraps), float32 / float16 within the
    tolerances stated in B200API's docstring, bigint equal as Python ints."""
    import numpy as np
    import tensor_network

    from conftest import load_golden
    from tensororder_b200.api import B200API

    cases = [("vc50_lineflow", None), ("vc50_lineflow", "min3"), ("vc100_lineflow", None)]
    if entry_type in ("int", "uint", "float32"):
        cases.append(("vc150_lineflow", "min4"))
    if entry_type in ("float32", "float16"):
        cases.append(("vc50_mcc_lineflow", None))  # weighted: the leaves themselves are rounded to the entry type
    for name, variant in cases:
        pp = load_golden(name)
        pp = pp.variant(variant) if variant else pp
        plan = reference.to_reference_plan(ref, pp, as_int=(entry_type == "bigint"))
        theirs = tensor_network.ALL_APIS["numpy"]()
        theirs.add_argument("entry_type", entry_type)
        ours = B200API()
        ours.add_argument("entry_type", entry_type)
        with np.errstate(over="ignore"):
            want = theirs.contract_sliced(plan)
        got = ours.contract_sliced(plan)
        if entry_type in ("int", "uint"):
            assert int(got) == int(want), (name, variant, got, want)
            assert type(got) is type(want)
        elif entry_type == "bigint":
            assert isinstance(got, int) and got == int(want), (name, variant)
        else:
            tol = 5e-6 if entry_type == "float32" else 5e-3
            if np.isfinite(want):
                assert abs(float(got) - float(want)) <= tol * abs(float(want)), (name, variant, got, want)
                assert type(got) is type(want)
            else:  # beyond the entry type's range the reference overflows; so does the rounded float64 result
                assert not np.isfinite(got) or abs(float(got)) > 6e4


@pytest.mark.gpu
def test_gpu_cli_under_torchrun_two_gpus(ref, tmp_path):
    """`torchrun --nproc-per-node 2 -m tensororder_b200.launch src/execution.py --tensor_library=b200 < N.con`: one
    process per GPU, slices r::2, one NCCL all-reduce; rank 0 prints the same count as the reference's numpy run."""
    from tensororder_b200 import cabi

    if cabi.lib.tob_device_count() < 2:
        pytest.skip("needs two GPUs")
    con = _con_bytes(ref, "vc150_lineflow", tmp_path)
    out_n, _, _ = _cli("execution.py", ["--rank_limit=19"], con, library="numpy")
    out_b, err_b, rc = _cli("execution.py", ["--rank_limit=19"], con, library="b200", torchrun=2)
    assert rc == 0, err_b[-2000:]
    assert out_b.count("Count:") == 1
    cb, cn = float(_field(out_b, "Count")), float(_field(out_n, "Count"))
    assert abs(cb - cn) <= 1e-9 * abs(cn)
    assert int(_field(out_b, "# Network Slices")) >= 2 and float(_field(out_b, "Contraction Time")) > 0
    print("2 GPUs: Count %r, Contraction Time %s s (numpy, one process: %s s)" % (cb, _field(out_b, "Contraction Time"), _field(out_n, "Contraction Time")))
