"""Run an unmodified TensorOrder CLI with the b200 backend registered.

    python -m tensororder_b200.launch /path/to/TensorOrder/src/tensororder.py \
        --planner=line-Flow --weights=unweighted --tensor_library=b200 < benchmark.cnf
    python -m tensororder_b200.launch /path/to/TensorOrder/src/execution.py --tensor_library=b200 < plan.con
    torchrun --nproc-per-node 8 -m tensororder_b200.launch .../src/execution.py --tensor_library=b200 < plan.con

`tensor_network.ALL_APIS` (src/tensor_network/__init__.py:12-16) is the live dict both CLIs hand to
`util.TaggedChoice` (src/tensororder.py:88-94, src/execution.py:34-40), so adding the key before the CLI
module executes is all the integration needs; no reference file is edited.  The shim also restores
`numpy.object`, which src/tensor_network/tensor_apis/numpy_apis.py:21 touches on every `entry_type`
call and numpy >= 1.24 removed.

Under torchrun (WORLD_SIZE > 1, one process per GPU) every rank must contract the SAME plan:
  * torchrun's workers inherit one stdin file description, so only rank 0 reads it; the bytes are
    broadcast and every rank gets its own in-memory stdin (text and binary views, as click opens them);
  * the planners are anytime and stop on wall-clock heuristics (src/planning.py:149-172), so two ranks
    would pick different trees and slicings: `planning.run` executes on rank 0 only and its result — the
    `SlicedExecutionPlan`, picklable like the `.con` files of `planning.py --store` — is broadcast;
  * rank 0 alone reports (`Count:` is identical on every rank after the all-reduce)."""
import io
import os
import pickle
import runpy
import sys


def _broadcast_bytes(dist, payload, src=0):
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def _share_stdin(dist):
    data = None
    if dist.get_rank() == 0:
        data = sys.stdin.buffer.read()
    data = _broadcast_bytes(dist, data)
    sys.stdin = io.TextIOWrapper(io.BytesIO(data), encoding="utf-8", errors="replace")  # .buffer is the BytesIO


def _plan_on_rank0(dist):
    """Wraps the reference's `planning.run` (src/planning.py:119-197): rank 0 plans, everyone gets its plan."""
    import planning  # the reference module (its src/ is on sys.path)

    original = planning.run

    def run_on_rank0(*args, **kwargs):
        payload = None
        if dist.get_rank() == 0:
            try:
                payload = pickle.dumps(("ok", original(*args, **kwargs)))
            except BaseException as exc:  # every rank must leave the broadcast, then fail the same way
                payload = pickle.dumps(("error", repr(exc)))
        status, value = pickle.loads(_broadcast_bytes(dist, payload))
        if status != "ok":
            raise RuntimeError("planning failed on rank 0: " + value)
        return value

    planning.run = run_on_rank0


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.stderr.write(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    src_dir = os.path.dirname(script)
    if src_dir not in sys.path:
        sys.path.insert(0, src_dir)
    import numpy

    if not hasattr(numpy, "object"):
        numpy.object = object
    from tensororder_b200.api import register
    from tensororder_b200.slicer import register as register_slicer

    register()
    register_slicer()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # launched by torchrun: one process per GPU; B200API shards the slices r::W and all-reduces the count
        import torch
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group("nccl")
        else:  # host-logic tests: the backend itself still fails loudly without a device
            dist.init_process_group("gloo")
        _share_stdin(dist)
        _plan_on_rank0(dist)
        if dist.get_rank() != 0:
            sys.stdout = open(os.devnull, "w")  # every rank computes the same Count; rank 0 reports it
    sys.argv = [script] + argv[1:]
    try:
        runpy.run_path(script, run_name="__main__")
    except SystemExit as exc:  # click exits through SystemExit: tear the process group down first
        _shutdown()
        raise exc
    _shutdown()
    return 0


def _shutdown():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
