"""Single propagations of the large configs through the drop-in call `tree.propagate(values)`
(host to host): compile time, first call (plan upload, CUDA-graph capture) and steady state.

    python junction-tree_b200/tools/single_instance.py
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))):
    if p not in sys.path:
        sys.path.insert(0, p)

import junctiontree as jt  # noqa: E402
import jt_workloads as wl  # noqa: E402


def main():
    for net in (wl.dag37(), wl.ising(16), wl.large_state_tree(), wl.dag500()):
        t0 = time.perf_counter()
        tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
        t1 = time.perf_counter()
        out = tree.propagate(net["values"])
        t2 = time.perf_counter()
        n = 5
        for _ in range(n):
            out = tree.propagate(net["values"])
        t3 = time.perf_counter()
        print(net["name"], "compile %.3f s, first call %.3f s, steady %.3f ms" % (t1 - t0, t2 - t1, (t3 - t2) / n * 1e3),
              "Z", float(out[0].sum()))


if __name__ == "__main__":
    main()
