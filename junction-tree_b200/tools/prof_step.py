"""The hot path alone (evidence slicing + init + collect + distribute [+ marginal]) for one config:
what the ncu launch lists under profiles/ are taken over, and a quick A-B timer.

    python junction-tree_b200/tools/prof_step.py --config large_state_tree --batch 512 [--dtype f32]
        [--no-uniform] [--no-evidence] [--no-beliefs] [--steps 5] [--warmup 3]

Prints one JSON line: ms per step (CUDA events), bytes the schedule moves, fraction of the HBM peak.
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import jt_workloads as wl  # noqa: E402
import junctiontree as jt  # noqa: E402
from junctiontree import _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="dag37")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-uniform", action="store_true")
    ap.add_argument("--no-evidence", action="store_true")
    ap.add_argument("--no-beliefs", action="store_true", help="outputs only (what the pipelines run)")
    args = ap.parse_args()
    net = {"dag37": wl.dag37, "dag500": wl.dag500, "ising16": lambda: wl.ising(16),
           "large_state_tree": wl.large_state_tree, "sprinkler": wl.sprinkler}[args.config]()
    dtype = np.dtype(np.float64 if args.dtype == "f64" else np.float32)
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = [] if args.no_evidence else list(net.get("evidence_vars", []))
    plan = tree.plan(evars)
    engine = tree._engine(plan.sizes, evars, plan.full_sizes)
    B = args.batch
    fdev, batched = engine.factors_to_device(net["values"], dtype)
    ev = torch.from_numpy(wl.draw_evidence(net, B)).cuda() if evars else None
    engine.dev.upload()
    if args.no_beliefs and not args.no_uniform:
        ws = engine.new_pipeline_workspace(B, dtype)      # sparse when that saves memory
    else:
        ws = engine.workspace(B, dtype)
    ws_ptr = ws.data_ptr()
    fout = torch.empty((plan.fout_entries, B), dtype=torch.float64 if dtype.itemsize == 8 else torch.float32,
                       device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    flags = (0 if args.no_uniform else _native.JT_UNIFORM)
    flags |= _native.JT_NO_BELIEFS if args.no_beliefs else _native.JT_SEP_BELIEFS
    ev_ptr = ev.data_ptr() if evars else None

    def step():
        engine.dev.init(fdev.data_ptr(), batched, ev_ptr, B, dtype, ws_ptr, flags, stream)
        engine.dev.collect(B, dtype, ws_ptr, flags, stream)
        engine.dev.distribute(B, dtype, ws_ptr, flags, stream)
        if args.no_beliefs:
            engine.dev.marginal(B, dtype, ws_ptr, fout.data_ptr(), stream, flags)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _native.launch_count()
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    uniform = not args.no_uniform and plan.uni_entries > 0
    S = dtype.itemsize * plan.scheduled_entries(uniform=uniform)
    A = dtype.itemsize * plan.algorithmic_entries()
    print(json.dumps({"config": args.config, "batch": B, "dtype": args.dtype, "uniform": uniform,
                      "beliefs": not args.no_beliefs, "evidence": bool(evars), "ms_per_step": ms,
                      "props_per_s": B / ms * 1e3, "launches_per_step": (_native.launch_count() - l0) / args.steps,
                      "scheduled_gb": S * B / 1e9, "scheduled_frac": S * B / ms / 1e6 / 6451.2,
                      "algorithmic_gb": A * B / 1e9, "algorithmic_frac": A * B / ms / 1e6 / 6451.2}))


if __name__ == "__main__":
    main()
