"""BASELINE.json configs[4] at full size: the 500-node DAG with a 1,048,576-instance evidence
batch sharded over the GPUs of one box, through the public API (`tree.marginals_batch`: host
evidence in, normalised single-variable posteriors + log Z out).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        junction-tree_b200/tools/config5_full.py [--total 1048576]

Prints one JSON line (rank 0): wall time of the slowest rank, propagations/s, chunk size.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import jt_workloads as wl  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    import junctiontree as jt
    from junctiontree import distributed as jdist
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1 << 20)
    args = ap.parse_args()
    rank, world = jdist.init_from_env("nccl")
    net = wl.dag500()
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    lo, hi = jdist.shard_bounds(args.total, world, rank)
    rng = np.random.default_rng(net["seed"] + 2000 + rank)
    card = np.array([net["sizes"][v] for v in evars])
    ev = rng.integers(0, card, size=(hi - lo, len(card))).astype(np.int32)
    free_vars = [v for v in sorted(net["sizes"]) if v not in evars]
    tree.marginals_batch(net["values"], free_vars, evars, ev[:4096])          # warm-up: plan, kernels, pinned pools
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    marg, log_z = tree.marginals_batch(net["values"], free_vars, evars, ev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert np.all(np.isfinite(log_z)) and all(np.allclose(m.sum(axis=1), 1.0) for m in list(marg.values())[:5])
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        plan = tree.plan(evars)
        engine = tree._engine(plan.sizes, evars, plan.full_sizes, outputs=[[v] for v in free_vars])
        print(json.dumps({"workload": "configs[4]: DAG-500, 100 observed variables, %d instances" % args.total,
                          "n_gpus": world, "seconds": float(t.item()), "propagations_per_s": args.total / float(t.item()),
                          "instances_per_gpu": hi - lo, "chunk": tree._chunk_for(engine, hi - lo, np.dtype(np.float64)),
                          "outputs": "%d normalised single-variable posteriors + log Z per instance" % len(free_vars)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
