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
    from tensororder_b200 import api as api_mod

    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = {}
    from program_sim import install_fake_device

    install_fake_device(api_mod.CompiledPlan, record=calls)
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


def _failing_worker(rank, world, port, kind, out):
    """Rank 1's device stage fails (OOM / timeout / other); rank 0's works.  Both must raise the same class
    instead of rank 0 blocking in the all-reduce or summing partials of different slicings."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist

    from conftest import load_golden
    from program_sim import install_fake_device
    from tensororder_b200 import api as api_mod

    dist.init_process_group("gloo", rank=rank, world_size=world)
    install_fake_device(api_mod.CompiledPlan)
    if rank == 1:
        exc = {"oom": api_mod.OutOfMemoryError("arena does not fit"), "timeout": TimeoutError("alarm"),
               "other": RuntimeError("CUDA error")}[kind]

        def failing_update(self, leaf_data=None):
            raise exc

        api_mod.CompiledPlan.update_leaves = failing_update
    pp = load_golden("vc50_lineflow").variant("min3")
    plan = pp.as_execution_plan()
    for entry_type in ("float64", "bigint"):
        api = api_mod.B200API()
        api.add_argument("entry_type", entry_type)
        try:
            api.contract_sliced(plan)
            out[(rank, entry_type)] = "no error"
        except api_mod.OutOfMemoryError:
            out[(rank, entry_type)] = "oom"
        except TimeoutError:
            out[(rank, entry_type)] = "timeout"
        except RuntimeError:
            out[(rank, entry_type)] = "other"
    # the ranks are still in step: a healthy collective call afterwards succeeds on both
    if rank == 1:
        install_fake_device(api_mod.CompiledPlan)
    api = api_mod.B200API()
    api.add_argument("entry_type", "float64")
    out[(rank, "after")] = float(api.contract_sliced(plan))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["oom", "timeout", "other"])
def test_a_failure_on_one_rank_is_raised_on_every_rank(kind):
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29850 + (os.getpid() % 100)
    mp.spawn(_failing_worker, args=(2, port, kind, out), nprocs=2, join=True)
    for rank in range(2):
        for entry_type in ("float64", "bigint"):
            assert out[(rank, entry_type)] == kind, (rank, entry_type, dict(out))
        assert out[(rank, "after")] == 2802717837.0


def _threaded_worker(rank, world, port, out):
    """Three host threads per rank, each contracting a different sliced plan; the threads reach their all-reduce in a
    DIFFERENT order on the two ranks (staggered sleeps).  Tickets keep the collectives paired."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import threading
    import time

    import torch.distributed as dist

    from conftest import load_golden
    from program_sim import install_fake_device
    from tensororder_b200 import api as api_mod

    dist.init_process_group("gloo", rank=rank, world_size=world)
    install_fake_device(api_mod.CompiledPlan)
    jobs = [("vc50_lineflow", "min3"), ("vc50_mcc_lineflow", "min3"), ("vc50_lineflow", "min3")]
    plans = [load_golden(n).variant(v).as_execution_plan() for n, v in jobs]
    order = api_mod.CollectiveOrder()
    results = {}

    def call(i, ticket):
        time.sleep(0.05 * (i if rank == 0 else (2 - i)))  # rank 0 arrives 0,1,2 — rank 1 arrives 2,1,0
        api = api_mod.B200API()
        api.add_argument("entry_type", "float64")
        api.add_argument("collective_ticket", (order, ticket))
        results[ticket] = float(api.contract_sliced(plans[i]))

    for step in range(2):
        threads = [threading.Thread(target=call, args=(i, 3 * step + i)) for i in range(3)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(60)
            assert not t.is_alive(), "deadlock at the collective turnstile"
    # a call that fails before its collective (the same way on every rank) gives its ticket up; the next one still runs
    api = api_mod.B200API()
    api.add_argument("entry_type", "float64")
    api.add_argument("collective_ticket", (order, 6))
    try:
        api.contract_sliced(object())
        results["bad"] = "no error"
    except Exception:  # noqa: BLE001
        results["bad"] = "raised"
    api.add_argument("collective_ticket", (order, 7))
    results[7] = float(api.contract_sliced(plans[0]))
    out[rank] = dict(results)
    dist.barrier()
    dist.destroy_process_group()


def test_threads_with_collective_tickets_stay_paired_across_ranks():
    sys.path.insert(0, HERE)
    from conftest import load_golden

    want = [load_golden(n).variant("min3").expected["count"] for n in ("vc50_lineflow", "vc50_mcc_lineflow", "vc50_lineflow")]
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29950 + (os.getpid() % 40)
    mp.spawn(_threaded_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in range(2):
        res = out[rank]
        for ticket in range(6):
            assert math.isclose(res[ticket], want[ticket % 3], rel_tol=1e-12), (rank, ticket, res)
        assert res["bad"] == "raised" and math.isclose(res[7], want[0], rel_tol=1e-12)
