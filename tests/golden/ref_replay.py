#!/usr/bin/env python3
"""Replay stored plans through the REFERENCE's own numpy executor and record its count.

Rebuilds reference objects from a portable plan — `TensorNetwork` (+`BuiltTensor` leaves, edges
connected in edge-id order), `ContractionTreeContext.leaf/join`, `SlicedExecutionPlan` with the stored
`groups_to_slice` — and calls `NumpyAPI.contract_sliced`.  Build container only (needs the scratch
build of /root/reference made by make_golden.py).

Usage: python tests/golden/ref_replay.py NAME[:VARIANT] [...]      (writes expected.count into the fixture)
       python tests/golden/ref_replay.py --bigint NAME[:VARIANT] [...]   (reference `--entry_type=bigint`:
                                                   exact Python-int count -> expected.count_exact, a decimal string)
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def to_reference_plan(R, pp, as_int=False):
    import numpy as np
    from contraction_methods.contraction_tree import ContractionTreeContext
    from tensor_network.tensor import BuiltTensor
    from tensor_network.tensor_network import TensorNetwork

    net = TensorNetwork()
    slots = []
    for t in pp.tensors:
        arr = np.array(t["data"], dtype=np.float64).reshape(t["shape"])
        if as_int:  # exact replays: Python ints in an object array, like the reference's own bigint leaves
            assert np.array_equal(arr, np.rint(arr))
            arr = np.array([int(x) for x in arr.reshape(-1)], dtype=object).reshape(t["shape"])
        slots.append(net.add_node(BuiltTensor(arr)))
    for e, (t1, t2) in enumerate(pp.edges):
        a1 = pp.index_lists[t1].index(e)
        a2 = pp.index_lists[t2].index(e)
        got = net.connect(t1, a1, t2, a2)
        assert got == e
    for t, il in enumerate(pp.index_lists):
        assert list(net.index_list(t)) == il
    ctx = ContractionTreeContext()
    ids = []
    for node in pp.postorder:
        if len(node) == 1:
            ids.append(ctx.leaf(net, node[0]))
        else:
            ids.append(ctx.join(ids[node[0]], ids[node[1]]))
    tree = ctx.get_tree(ids[-1])
    plan = R["sliced_execution_plan"].SlicedExecutionPlan(tree, net)
    plan.groups_to_slice = [set(g) for g in pp.groups_to_slice]
    plan.edges_to_slice = set().union(*plan.groups_to_slice) if plan.groups_to_slice else set()
    return plan


def main():
    from tensororder_b200.plan_format import PortablePlan

    mg.ensure_reference_build()
    R = mg.import_reference()
    bigint = "--bigint" in sys.argv
    for spec in [a for a in sys.argv[1:] if a != "--bigint"]:
        name, _, variant = spec.partition(":")
        path = os.path.join(HERE, name + ".json.gz")
        pp = PortablePlan.load(path)
        target = pp.variant(variant) if variant else pp
        plan = to_reference_plan(R, target, as_int=bigint)
        t0 = time.time()
        if bigint:
            api = R["tensor_network"].ALL_APIS["numpy"]()
            api.add_argument("entry_type", "bigint")
            exact = api.contract_sliced(plan)
            assert isinstance(exact, int), type(exact)
            pp = PortablePlan.load(path)
            rec = {"count_exact": str(exact), "count_exact_source": "reference numpy backend, --entry_type=bigint"}
            if variant:
                [v for v in pp.variants if v["name"] == variant][0]["expected"].update(rec)
            else:
                pp.expected.update(rec)
            pp.save(path)
            print("%s: exact count=%d  [%.1fs]" % (spec, exact, time.time() - t0), flush=True)
            continue
        count, dt, _ = mg.reference_contract(R, plan)
        rec = {"count": count, "count_hex": float(count).hex(), "numpy_seconds_buildbox_8c": dt,
               "count_source": "tests/golden/ref_replay.py (reference numpy executor on the stored plan)"}
        pp = PortablePlan.load(path)  # reload: another replay may have updated the file meanwhile
        if variant:
            [v for v in pp.variants if v["name"] == variant][0]["expected"].update(rec)
        else:
            pp.expected.update(rec)
        pp.save(path)
        print("%s: count=%r  [%.1fs]" % (spec, count, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
