#!/usr/bin/env python3
"""Replay stored plans through the REFERENCE's own numpy executor and record its count.

Rebuilds reference objects from a portable plan — `TensorNetwork` (+`BuiltTensor` leaves, edges
connected in edge-id order), `ContractionTreeContext.leaf/join`, `SlicedExecutionPlan` with the stored
`groups_to_slice` (`oracle.reference.to_reference_plan`) — and calls `NumpyAPI.contract_sliced`.  Needs the reference
built into oracle/_ref/ (`python -m oracle.reference`).

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


from oracle.reference import to_reference_plan  # noqa: E402,F401  (kept importable from here)


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
