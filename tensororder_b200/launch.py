"""Run an unmodified TensorOrder CLI with the b200 backend registered.

    python -m tensororder_b200.launch /path/to/TensorOrder/src/tensororder.py \
        --planner=line-Flow --weights=unweighted --tensor_library=b200 < benchmark.cnf
    python -m tensororder_b200.launch /path/to/TensorOrder/src/execution.py --tensor_library=b200 < plan.con

`tensor_network.ALL_APIS` (src/tensor_network/__init__.py:12-16) is the live dict both CLIs hand to
`util.TaggedChoice` (src/tensororder.py:88-94, src/execution.py:34-40), so adding the key before the CLI
module executes is all the integration needs; no reference file is edited.  The shim also restores
`numpy.object`, which src/tensor_network/tensor_apis/numpy_apis.py:21 touches on every `entry_type`
call and numpy >= 1.24 removed."""
import os
import runpy
import sys


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
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        if dist.get_rank() != 0:
            sys.stdout = open(os.devnull, "w")  # every rank computes the same Count; rank 0 reports it
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
