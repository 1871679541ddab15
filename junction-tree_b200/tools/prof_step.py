"""The hot path alone (evidence slicing + init + collect + distribute [+ marginal]) for one config:
the command the ncu launch lists under profiles/ are taken over, and a quick A-B timer.  Runs
exactly the step `bench.py` times for its `configs` block (`jt_bench_lib.HotPath`).

    python junction-tree_b200/tools/prof_step.py --config large_state_tree --batch 512 [--dtype f32]
        [--no-uniform] [--no-evidence] [--no-beliefs] [--no-dense] [--steps 5] [--warmup 3]

Prints one JSON line: ms per step (CUDA events), bytes the schedule moves, fraction of the HBM peak.
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import jt_bench_lib as bl  # noqa: E402


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
    ap.add_argument("--no-dense", action="store_true", help="keep dense contractions on the projection kernels")
    ap.add_argument("--peak", type=float, default=6451.2, help="HBM GB/s the fractions are quoted against")
    ap.add_argument("--uniform-valid", action="store_true",
                    help="analysis only: time the steps with JT_UNIFORM_VALID (the uniform phases skipped) to see "
                         "how much of a step they cost; never a bench number")
    ap.add_argument("--compare", action="store_true",
                    help="also time the same step with JT_NO_DENSE in this process (same box, same clocks)")
    args = ap.parse_args()
    hp = bl.HotPath(args.config, args.batch, args.dtype, uniform=not args.no_uniform, evidence=not args.no_evidence,
                    beliefs=not args.no_beliefs, dense=not args.no_dense)
    t = hp.time(args.steps, args.warmup)
    if args.uniform_valid:
        from junctiontree import _native
        hp.flags |= _native.JT_UNIFORM_VALID
        t["ms_uniform_valid"] = hp.time(args.steps, args.warmup)["ms_per_step"]
        hp.flags &= ~_native.JT_UNIFORM_VALID
    base_ms = None
    if args.compare and hp.dense:
        from junctiontree import _native
        hp.flags |= _native.JT_NO_DENSE
        base_ms = hp.time(args.steps, args.warmup)["ms_per_step"]
        hp.flags &= ~_native.JT_NO_DENSE
        t2 = hp.time(args.steps, args.warmup)              # and the default once more (drift check)
        t["ms_per_step_again"] = t2["ms_per_step"]
    A, A_msg, S, S_msg = hp.bytes_per_propagation()
    ms, B = t["ms_per_step"], hp.B
    print(json.dumps({"config": args.config, "batch": B, "dtype": args.dtype, "uniform": hp.uniform,
                      "beliefs": hp.beliefs, "evidence": bool(hp.evars), "dense": hp.dense, "sparse_workspace": hp.sparse,
                      "ms_per_step": ms, "ms_per_step_no_dense": base_ms, "ms_per_step_again": t.get("ms_per_step_again"), "ms_uniform_valid": t.get("ms_uniform_valid"),
                      "init_ms": t["init_ms"], "message_passing_ms": t["msg_ms"],
                      "props_per_s": B / ms * 1e3, "launches_per_step": t["launches_per_step"],
                      "scheduled_gb": S * B / 1e9, "scheduled_frac": S * B / ms / 1e6 / args.peak,
                      "algorithmic_gb": A * B / 1e9, "algorithmic_frac": A * B / ms / 1e6 / args.peak}))


if __name__ == "__main__":
    main()
