"""Multi-GPU host logic on CPU: batch sharding and the all-gather of requested marginals,
world_size 2 over gloo (the data path itself has no collective)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from junctiontree import distributed as jdist


def test_shard_bounds_partition_the_batch():
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            bounds = [jdist.shard_bounds(total, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == total
            for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
                assert a1 == b0
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1
            assert sizes == jdist.shard_sizes(total, world)
    with pytest.raises(ValueError):
        jdist.shard_bounds(10, 2, 2)


def test_pack_marginals_concatenates_factor_outputs():
    outs = [torch.arange(6.).reshape(3, 2), torch.arange(12.).reshape(3, 2, 2)]
    packed = jdist.pack_marginals(outs)
    assert packed.shape == (3, 6)
    assert torch.equal(packed[:, :2], outs[0]) and torch.equal(packed[:, 2:], outs[1].reshape(3, 4))
    assert torch.equal(jdist.pack_marginals(outs, requested=[1]), outs[1].reshape(3, 4))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, width, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w = jdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    full = torch.arange(total * width, dtype=torch.float64).reshape(total, width)
    lo, hi = jdist.shard_bounds(total, world, rank)
    gathered = jdist.all_gather_rows(full[lo:hi].clone(), total)
    ok = torch.equal(gathered, full)
    # a wrong shard size is refused
    refused = False
    try:
        jdist.all_gather_rows(full[lo:hi + 1 if hi < total else hi - 1].clone(), total)
    except ValueError:
        refused = True
    np.save(os.path.join(result_dir, "rank%d.npy" % rank), np.array([ok, refused]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7])
def test_all_gather_rows_world_size_2_gloo(tmp_path, total):
    """Even and ragged shards come back in global instance order on every rank."""
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, total, 5, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        ok, refused = np.load(os.path.join(str(tmp_path), "rank%d.npy" % rank))
        assert ok and refused


class _FakeTree:
    """Stands in for a JunctionTree on a machine without a GPU: ``propagate_batch`` returns CPU
    tensors whose values encode the global instance index, ``plan`` gives the output shapes."""

    class _Plan:
        fout_shape = [[2], [2, 3]]

    def plan(self, evidence_vars=()):
        return self._Plan()

    def propagate_batch(self, xs, evidence_vars, evidence, dtype=None, device_output=True):
        rows = torch.as_tensor(np.asarray(evidence)[:, 0], dtype=torch.float64)
        return [rows[:, None].repeat(1, 2), rows[:, None, None].repeat(1, 2, 3)]


def _sharded_worker(rank, world, port, total, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    jdist.init_from_env("gloo")
    evidence = np.arange(total, dtype=np.int32).reshape(total, 1)      # instance b carries the value b
    local, gathered = jdist.propagate_sharded(_FakeTree(), [np.ones(2), np.ones((2, 3))], ["v"], evidence)
    lo, hi = jdist.shard_bounds(total, world, rank)
    ok = local[0].shape[0] == hi - lo and tuple(gathered.shape) == (total, 8)
    ok = ok and torch.equal(gathered[:, 0], torch.arange(total, dtype=torch.float64))
    np.save(os.path.join(result_dir, "sharded%d.npy" % rank), np.array([ok, hi - lo]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [1, 5])
def test_propagate_sharded_with_fewer_instances_than_ranks(tmp_path, total):
    """a rank whose shard is empty skips the compute but still joins the all-gather
    (its peers would block otherwise); world size 2 over gloo."""
    world, port = 2, _free_port()
    mp.spawn(_sharded_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    sizes = []
    for rank in range(world):
        ok, n = np.load(os.path.join(str(tmp_path), "sharded%d.npy" % rank))
        assert ok
        sizes.append(int(n))
    assert sum(sizes) == total and (total > 1 or 0 in sizes)
