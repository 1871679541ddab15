"""Single-propagate latency of the drop-in API (BASELINE.json configs[0]: README network).

    python tests/tools/latency.py

Prints one JSON line per network: microseconds per `tree.propagate(values)` call (host arrays in,
host arrays out; one library call for small trees) and per `compute_beliefs` call, next to the
NumPy oracle on one host core.  Test infrastructure: it times the oracle, so it lives under tests/.
"""

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "junction-tree_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import jt_workloads as wl  # noqa: E402
import junctiontree as jt  # noqa: E402
from oracle import ref_fixed  # noqa: E402


def main():
    for net in (wl.sprinkler(), wl.huang_darwiche(), wl.wisconsin()):
        tree = jt.create_junction_tree(net["factors"], net["sizes"])
        vals = net["values"]
        for _ in range(20):
            out = tree.propagate(vals)
        n = 2000
        t0 = time.perf_counter()
        for _ in range(n):
            out = tree.propagate(vals)
        gpu_us = (time.perf_counter() - t0) / n * 1e6
        ct = tree.clique_tree
        args = (tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique, net["factors"], net["sizes"], vals)
        for _ in range(20):
            want, _ = ref_fixed.propagate(*args)
        t0 = time.perf_counter()
        for _ in range(200):
            want, _ = ref_fixed.propagate(*args)
        cpu_us = (time.perf_counter() - t0) / 200 * 1e6
        err = max(float(np.max(np.abs(o - w) / np.maximum(np.abs(w).max(), 1e-300))) for o, w in zip(out, want))
        # the operator-level call of the reference: compute_beliefs(tree, potentials, clique_vars)
        from junctiontree import computation as comp
        psi = ref_fixed.evaluate(net["factors"], vals, ct.maxcliques, ct.factor_to_maxclique, net["sizes"])
        pots = psi + [np.ones([net["sizes"][v] for v in s]) for s in tree.separators]
        node_vars = ct.maxcliques + tree.separators
        for _ in range(20):
            comp.compute_beliefs(tree.tree, pots, node_vars)
        t0 = time.perf_counter()
        for _ in range(n):
            comp.compute_beliefs(tree.tree, pots, node_vars)
        cb_us = (time.perf_counter() - t0) / n * 1e6
        print(json.dumps({"network": net["name"], "propagate_us": round(gpu_us, 1), "compute_beliefs_us": round(cb_us, 1),
                          "numpy_oracle_us_1core": round(cpu_us, 1), "max_rel_err": err}))


if __name__ == "__main__":
    main()
