"""Experiment: the same batch as N independent sub-batches on N CUDA streams, so that the tail
of one sub-batch's level launch overlaps the body of another's (instances are independent).

    python junction-tree_b200/tools/split_streams.py --config dag37 --batch 65536 --splits 1 2 4
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "junction-tree_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import jt_workloads as wl  # noqa: E402


def main():
    import torch
    import junctiontree as jt
    from junctiontree import _native
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="dag37")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--splits", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-uniform", action="store_true")
    args = ap.parse_args()
    net = {"dag37": wl.dag37, "dag500": wl.dag500, "ising16": lambda: wl.ising(16),
           "large_state_tree": wl.large_state_tree}[args.config]()
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = list(net.get("evidence_vars", []))
    plan = tree.plan(evars)
    engine = tree._engine(plan.sizes, evars, plan.full_sizes)
    dtype = np.dtype(np.float64)
    fdev, _ = engine.factors_to_device(net["values"], dtype)
    B = args.batch
    ev = torch.from_numpy(wl.draw_evidence(net, B)).cuda()
    flags = _native.JT_SEP_BELIEFS | (0 if args.no_uniform else _native.JT_UNIFORM)
    engine.dev.upload()
    main_stream = torch.cuda.current_stream()
    for n in args.splits:
        sub = B // n
        streams = [torch.cuda.Stream() for _ in range(n)]
        wss = [engine.new_workspace(sub, dtype) for _ in range(n)]

        def step():
            fork = torch.cuda.Event()
            fork.record(main_stream)
            for k, (st, ws) in enumerate(zip(streams, wss)):
                st.wait_event(fork)
                evp = ev[k * sub:(k + 1) * sub].data_ptr()
                engine.dev.init(fdev.data_ptr(), False, evp, sub, dtype, ws.data_ptr(), flags, st.cuda_stream)
                engine.dev.collect(sub, dtype, ws.data_ptr(), flags, st.cuda_stream)
                engine.dev.distribute(sub, dtype, ws.data_ptr(), flags, st.cuda_stream)
                main_stream.wait_stream(st)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print("splits %d: %.3f ms per step, %.0f props/s" % (n, ms, B / ms * 1e3), flush=True)
        del wss


if __name__ == "__main__":
    main()
