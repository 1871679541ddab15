"""Batch sharding over two GPUs with the NCCL all-gather of requested marginals (skipped on a
single-GPU box; the host logic is covered on CPU by tests/test_distributed_cpu.py)."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import jt_workloads as wl
    import junctiontree as jt
    from junctiontree import distributed as jdist
    from oracle import ref_fixed
    jdist.init_from_env("nccl")
    net = wl.random_dag(12, 3, 2, 3, 8, 5)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    B = 301                                             # ragged split: 151 + 150
    ev = wl.draw_evidence(net, B)
    requested = [0, 3, 7]
    local, gathered = jdist.propagate_sharded(tree, net["values"], net["evidence_vars"], ev, requested=requested)
    lo, hi = jdist.shard_bounds(B, world, rank)
    assert local[0].shape[0] == hi - lo and gathered.shape[0] == B
    ct = tree.clique_tree
    pick = [0, 150, 151, 300]
    want, _ = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                        net["factors"], net["sizes"], net["values"], net["evidence_vars"],
                                        ev[pick], n=len(pick))
    expect = np.concatenate([want[f].reshape(len(pick), -1) for f in requested], axis=1)
    np.testing.assert_allclose(gathered[pick].cpu().numpy(), expect, rtol=1e-12)
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_propagation_and_nccl_all_gather(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d.npy" % r)) for r in range(world))
