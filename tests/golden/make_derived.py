"""Fingerprints of what ``jt_plan_create`` derives from a plan blob when it loads it: the dense
contractions (``jt_plan_dense_get``) and the int32 table behind them, which also holds the walk
tables of the belief kernel (``jt_plan_dense_table``).  The kernels read these tables as they
are, so a change of the plan-load code that is meant to be a pure speed-up must leave every
fingerprint unchanged -- checked on the CPU by ``tests/test_schedule_abi.py::
test_plan_load_derivation_is_pinned``.  After an intended change (another walk order, another
selection rule) rerun ``python tests/golden/make_derived.py`` and commit the new file together
with the GPU parity run that covers it.

(Written for the round-2 change that moved the classification to direct tables and the walk
tables onto all host cores: old and new library gave identical fingerprints on these plans.)
"""

import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for _p in (ROOT, os.path.join(ROOT, "junction-tree_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

OUT = os.path.join(HERE, "derived_tables.json")


def nets():
    import jt_bench_lib as bl
    import jt_workloads as wl
    return [("sprinkler", wl.sprinkler()), ("dag37", wl.dag37()), ("ising16", bl.make_net("ising16")),
            ("large_state_tree", bl.make_net("large_state_tree")), ("dag500", wl.dag500()),
            ("large_state_small", wl.large_state_tree((8, 12, 16, 8, 12, 16))),
            ("large_state_odd", wl.large_state_tree((5, 7, 9, 3, 11, 6))),
            ("dag60", wl.random_dag(60, 3, 2, 5, 8, 3)), ("dag24_wide", wl.random_dag(24, 4, 3, 6, 6, 11))]


def fingerprints():
    from helpers import compile_net
    from junctiontree import _native
    from junctiontree import schedule as sch
    out = {}
    for name, net in nets():
        tree, seps, mc, f2c, eff, evars = compile_net(net)
        plan = sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"])
        dp = _native.DevicePlan(plan.to_blob())
        tasks, table = dp.dense_tasks()
        digest = hashlib.sha256()
        digest.update(json.dumps(tasks, sort_keys=True).encode())
        digest.update(table.tobytes())
        out[name] = {"dense_tasks": len(tasks), "table_entries": int(table.size), "sha256": digest.hexdigest()}
        dp.close()
    return out


if __name__ == "__main__":
    with open(OUT, "w") as fh:
        json.dump(fingerprints(), fh, indent=1, sort_keys=True)
    print("wrote", OUT)
