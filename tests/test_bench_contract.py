"""bench.py contract on CPU: the reference arm (the real reference from oracle/_ref timed on host cores) prints one JSON line
with the keys the driver reads; the b200 arm refuses to run without its CUDA library/device."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--max-n", "70"], capture_output=True, text=True, cwd=REPO, timeout=300)
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["counts_ok"] is True
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "s/instance" and line["higher_is_better"] is False and line["dtype"] == "f64"
    sys.path.insert(0, REPO)
    from oracle import reference

    assert line["cpu_baseline"]["kind"] == ("reference" if reference.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["truncated"] is False and line["instances"] == 3 and "3 of 3" in line["cpu_baseline"]["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--max-n", "60"], capture_output=True, text=True, cwd=REPO, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""
