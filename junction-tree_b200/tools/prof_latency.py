"""Where the time of one `tree.propagate(values)` call goes (host to host, small networks).

    python junction-tree_b200/tools/prof_latency.py [sprinkler|huang_darwiche|wisconsin]
"""
import cProfile
import os
import pstats
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import jt_workloads as wl  # noqa: E402
import junctiontree as jt  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "sprinkler"
    net = getattr(wl, name)()
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    vals = net["values"]
    for _ in range(50):
        tree.propagate(vals)
    fg = tree.clique_tree.factor_graph
    engine = tree._engine(dict(fg.sizes))
    runner = engine.host_runner(1, np.dtype(np.float64))
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n):
        runner.run()
    print("%s: runner.run() (copies + kernel + sync, one library call): %.1f us"
          % (name, (time.perf_counter() - t0) / n * 1e6))
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(runner.stream):
        start.record()
        for _ in range(200):
            engine.dev.propagate(runner.factors.data_ptr(), False, None, 1, np.dtype(np.float64), runner.ws.data_ptr(),
                                 runner.fout.data_ptr(), runner.flags, runner.stream.cuda_stream)
        end.record()
    torch.cuda.synchronize()
    print("kernel alone, back to back: %.1f us" % (start.elapsed_time(end) / 200 * 1e3))
    t0 = time.perf_counter()
    for _ in range(n):
        runner.set_factors(vals)
    print("set_factors: %.1f us" % ((time.perf_counter() - t0) / n * 1e6))
    t0 = time.perf_counter()
    for _ in range(n):
        tree.propagate(vals)
    print("propagate total: %.1f us" % ((time.perf_counter() - t0) / n * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2000):
        tree.propagate(vals)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(8)


if __name__ == "__main__":
    main()
