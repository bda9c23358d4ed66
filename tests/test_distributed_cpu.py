"""Multi-process host logic of the sliced path on CPU: world_size-2 `gloo`, each rank runs
`B200API.contract_sliced` with the device stage replaced by the numpy program interpreter
(tests/program_sim.py) — the slice partition r, r+W, ..., `num_slice_limit`, and the all-reduce of the
partial counts are the product code under test."""
import math
import os
import sys

import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _worker(rank, world, port, name, variant, limit, out):
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist

    from conftest import load_golden
    from program_sim import run_program
    from tensororder_b200 import api as api_mod

    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = {}

    def fake_upload(self):
        self.uploaded = True

    def fake_run(self, first=0, count=None, stride=1, initial=0.0, skip_invariant=False):
        # the interruptible loop issues chunks; record the whole range this rank was given
        a = calls.setdefault("args", [first, 0, stride])
        a[1] += count
        return initial + (run_program(self.describe(), self.flat, first=first, count=count, stride=stride) if count else 0.0)

    api_mod.CompiledPlan.upload = fake_upload
    api_mod.CompiledPlan.run = fake_run
    api_mod.CompiledPlan.last_ms = 0.0
    api_mod.CompiledPlan.last_launches = 0
    pp = load_golden(name).variant(variant)
    api = api_mod.B200API()
    api.add_argument("entry_type", "float64")
    got = api.contract_sliced(pp.as_execution_plan(), num_slice_limit=limit)
    out[rank] = (float(got), tuple(calls["args"]), api.last_stats["world"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,variant,limit", [("vc50_lineflow", "min3", None), ("vc50_lineflow", "min3", 3),
                                                ("vc50_mcc_lineflow", "min3", None), ("vc50_lineflow", "min3", 1)])
def test_two_ranks_partition_and_allreduce(name, variant, limit):
    sys.path.insert(0, HERE)
    from conftest import load_golden

    exp = load_golden(name).variant(variant).expected
    want = exp["count"] if limit is None else sum(exp["per_slice"][:limit])
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, name, variant, limit, out), nprocs=2, join=True)
    total = exp["num_slices"] if limit is None else min(limit, exp["num_slices"])
    for rank in range(2):
        got, (first, count, stride), world = out[rank]
        assert world == 2
        assert math.isclose(got, want, rel_tol=1e-12), (rank, got, want)  # identical on every rank after the all-reduce
        assert stride == 2
        assert count == (0 if rank >= total else (total - rank + 1) // 2)
        if count:
            assert first == rank
