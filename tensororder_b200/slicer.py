"""GPU-aware slicer: slices a plan until the EXECUTOR's own arena fits.

The reference's `GreedyMemSlicer` (src/tensor_network/slicers.py:27-33) slices until the cost model's
estimate `max(..., out + 2*left + 2*right)` (src/contraction_methods/contraction_tree.pyx:423-446) drops
below `--mem_limit`; that estimate charges two transposed copies per join which this backend never makes,
so it slices more than necessary (each extra slice repeats the slice-dependent part of the tree).
`B200MemSlicer` asks the plan compiler (`tob_plan_peak_bytes`, host only) instead and otherwise behaves
like the reference slicer: same edge choice (`plan.next_edge_to_slice`), same `slice_until` contract
(memory in ENTRIES, as `tensororder.py:221-222` converts it)."""
from .api import CompiledPlan
from .flatten import flatten_plan


def plan_peak_bytes(plan, mem_limit_bytes: int = 0) -> int:
    """Device bytes the executor needs for `plan` (host-only compile).  With a budget the compiler drops its
    optional second slice lane before it would exceed it, exactly as `B200API` compiles under `mem_limit_bytes`."""
    cp = CompiledPlan(flatten_plan(plan), mem_limit_bytes=int(mem_limit_bytes))
    try:
        return cp.peak_bytes
    finally:
        cp.close()


class B200MemSlicer:
    def slice_once(self, plan):
        plan.slice_at(plan.next_edge_to_slice)  # GreedyMemSlicer.slice_once, slicers.py:32-33

    def slice_until(self, plan, memory=None, rank=None, slices=None):
        """Same signature and units as BaseSlicer.slice_until (slicers.py:10-24)."""
        while memory is not None and memory * 8 < plan_peak_bytes(plan, int(memory * 8)):
            if plan.next_edge_to_slice is None or plan.next_edge_to_slice < 0 or len(plan.groups_to_slice) >= 60:
                raise RuntimeError("b200_mem slicer: the plan cannot be sliced below %d bytes "
                                   "(leaf tensors and tables alone need more)" % int(memory * 8))
            self.slice_once(plan)
        while rank is not None and rank < plan.maxrank:
            self.slice_once(plan)
        while slices is not None and len(plan.groups_to_slice) < slices:
            self.slice_once(plan)


def register(all_slicers=None):
    """Adds "b200_mem" to the reference's `tensor_network.ALL_SLICERS` (slicers.py:119-124)."""
    if all_slicers is None:
        import tensor_network  # type: ignore

        all_slicers = tensor_network.ALL_SLICERS
    all_slicers["b200_mem"] = B200MemSlicer()
    return all_slicers
