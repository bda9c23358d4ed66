"""Drop-in check against the REAL reference (build container only: needs /root/reference; skipped on the
GPU box).  Registers "b200" in the reference's live registry, resolves it through the reference's own
click option type, and drives `execution.run` with a reference-built plan up to the C ABI.  Without a
GPU the backend must fail loudly (no CPU fallback) and the reference must report it the way it reports
any backend failure (execution.py:143-152)."""
import io
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
BUILD = os.environ.get("TENSORORDER_REF_BUILD", "/tmp/ref_probe")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg

    cwd = os.getcwd()
    mg.ensure_reference_build()
    R = mg.import_reference()
    yield R
    os.chdir(cwd)


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

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import ref_replay
    from conftest import load_golden
    from tensororder_b200.flatten import flatten_plan

    pp = load_golden("vc50_lineflow").variant("min3")
    live = ref_replay.to_reference_plan(ref, pp)
    a = flatten_plan(live)
    b = flatten_plan(pp.as_execution_plan())
    for f in ("node_left", "node_right", "node_leaf", "leaf_rank", "leaf_data_offset", "leaf_axis_start",
              "leaf_axis_edge", "leaf_data"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.n_slice_groups == b.n_slice_groups == 3


def test_execution_run_reaches_the_cabi(ref, capsys):
    import execution  # the reference's src/execution.py
    import tensor_network

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import ref_replay
    from conftest import load_golden
    from tensororder_b200 import cabi
    from tensororder_b200.api import B200API

    plan = ref_replay.to_reference_plan(ref, load_golden("vc50_lineflow"))
    api = B200API()
    api.add_argument("entry_type", "float64")
    ref["util"].set_verbosity(0)
    result = execution.run(plan, api, tensor_network.ALL_SLICERS["greedy_mem"], None)
    if cabi.lib.tob_device_count() > 0:
        assert float(result) == 2802717837.0
    else:
        # no GPU here: the backend raises, the reference prints its generic backend-failure line
        assert result is None
        out = capsys.readouterr()
        assert "Exception during execution" in out.out
        assert "CUDA device" in out.err or "CUDA device" in out.out


def test_launcher_runs_the_unmodified_cli_help():
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    res = subprocess.run([sys.executable, "-m", "tensororder_b200.launch", os.path.join(BUILD, "src", "execution.py"), "--help"],
                         capture_output=True, text=True, env=env, cwd=BUILD)
    assert res.returncode == 0, res.stderr
    assert "b200" in res.stdout  # listed among the --tensor_library choices


def test_gpu_aware_slicer_slices_less_than_the_reference_model(ref):
    """Same plan, same byte budget: the reference cost model (out + 2*left + 2*right entries) needs more
    slices than the executor's real arena; the count is unchanged (interpreted from the compiled program)."""
    import math

    import tensor_network

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import ref_replay
    from conftest import load_golden
    from program_sim import run_program
    from tensororder_b200.api import CompiledPlan
    from tensororder_b200.flatten import flatten_plan
    from tensororder_b200.slicer import B200MemSlicer, plan_peak_bytes, register

    register()
    assert "b200_mem" in tensor_network.ALL_SLICERS
    pp = load_golden("vc250_lineflow")
    assert pp.expected["maxrank"] == 31
    budget_bytes = 40e9  # between the executor's real need (30.6 GB) and the reference model's estimate (55.8 GB)
    ref_plan = ref_replay.to_reference_plan(ref, pp)
    assert ref_plan.memory * 8 > budget_bytes
    tensor_network.ALL_SLICERS["greedy_mem"].slice_until(ref_plan, memory=budget_bytes / 8)
    ours = ref_replay.to_reference_plan(ref, pp)
    B200MemSlicer().slice_until(ours, memory=budget_bytes / 8)
    assert plan_peak_bytes(ours) <= budget_bytes
    assert len(ours.groups_to_slice) == 0 < len(ref_plan.groups_to_slice)
    # a budget that needs real slicing: still met, with at most as many slices as the reference model asks for
    tight = 8e9
    a = ref_replay.to_reference_plan(ref, pp)
    tensor_network.ALL_SLICERS["greedy_mem"].slice_until(a, memory=tight / 8)
    b = ref_replay.to_reference_plan(ref, pp)
    B200MemSlicer().slice_until(b, memory=tight / 8)
    assert plan_peak_bytes(b) <= tight and 0 < len(b.groups_to_slice) <= len(a.groups_to_slice)
    pp = load_golden("vc100_lineflow")
    with pytest.raises(RuntimeError):
        B200MemSlicer().slice_until(ref_replay.to_reference_plan(ref, pp), memory=16)  # 128 bytes: impossible
