"""Multi-GPU execution: shard the evidence batch, one process per GPU.

Instances are independent (same tree, same factor tables, different evidence), so the batch is
split contiguously over the ranks and there is no exchange during propagation.  The only
collective is the optional all-gather of requested marginals (NCCL over NVLink on GPUs, gloo in
the CPU tests of the host logic).  The reference has no distributed code at all.
"""

import os

import numpy as np


def shard_bounds(total, world_size, rank):
    """Contiguous, balanced split of ``range(total)``: the first ``total % world_size`` ranks get
    one extra instance."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of size %d" % (rank, world_size))
    base, extra = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total, world_size):
    return [shard_bounds(total, world_size, r)[1] - shard_bounds(total, world_size, r)[0]
            for r in range(world_size)]


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host
    buffers (first touch) and the copy threads are local to the GPU's PCIe root.  With eight
    ranks streaming results to the host at once, remote-socket buffers halve the aggregate
    device->host rate.  Best effort: silently does nothing when the topology is not exposed."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def init_from_env(backend=None):
    """Initialise ``torch.distributed`` from RANK / WORLD_SIZE / MASTER_* (torchrun); binds the
    process to ``cuda:LOCAL_RANK`` when the backend is NCCL.  Returns ``(rank, world_size)``."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        torch.cuda.set_device(local)
        if world > 1:
            bind_to_gpu_numa_node(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def all_gather_rows(local, total, group=None):
    """All-gather of per-instance rows.

    :param local: tensor ``[n_local, M]`` holding this rank's shard (``shard_bounds`` split)
    :param total: global number of instances
    :return: tensor ``[total, M]`` on every rank, rows in global instance order
    """
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError("local shard has %d rows, expected %d" % (local.shape[0], sizes[dist.get_rank(group)]))
    width = max(sizes)
    padded = local
    if local.shape[0] != width:       # equal-size buffers for the collective
        padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    gathered = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if all(s == width for s in sizes):
        return gathered
    return torch.cat([gathered[r * width: r * width + sizes[r]] for r in range(world)], dim=0)


def pack_marginals(outputs, requested=None):
    """Concatenate per-factor outputs ``[n_local, *shape]`` into one ``[n_local, M]`` tensor."""
    import torch
    if requested is None:
        requested = range(len(outputs))
    def width(t):               # explicit: reshape(-1) is ambiguous for the empty shard of a small batch
        n = 1
        for d in t.shape[1:]:
            n *= int(d)
        return n

    cols = [outputs[f].reshape(outputs[f].shape[0], width(outputs[f])) for f in requested]
    return torch.cat(cols, dim=1)


def _empty_outputs(tree, xs, evidence_vars, dtype):
    """Per-factor outputs of an empty shard: ``[0, *factor_shape]`` tensors (shapes from the plan)."""
    import torch
    plan = tree.plan(list(evidence_vars))
    if dtype is None:
        dtype = np.float32 if xs and all(np.asarray(x).dtype == np.float32 for x in xs) else np.float64
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    device = "cuda" if torch.cuda.is_available() else "cpu"
    return [torch.empty((0,) + tuple(shape), dtype=tdt, device=device) for shape in plan.fout_shape]


def propagate_sharded(tree, xs, evidence_vars, evidence, requested=None, gather=True, dtype=None,
                      group=None):
    """Propagate this rank's shard of the batch; optionally all-gather requested marginals.

    :param evidence: the *global* ``[B, |E|]`` evidence array (every rank passes the same) --
                     each rank slices its own contiguous shard
    :param requested: factor indices whose outputs are gathered (default: all)
    :return: ``(local_outputs, gathered)`` -- per-factor CUDA tensors for the local shard and the
             ``[B, M]`` gathered marginals (``None`` when ``gather`` is False)
    """
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    B = int(np.shape(evidence)[0])
    lo, hi = shard_bounds(B, world, rank)
    if hi > lo:
        local = tree.propagate_batch(xs, evidence_vars, evidence[lo:hi], dtype=dtype, device_output=True)
    else:
        # fewer instances than ranks: nothing to compute here, but the collective below still
        # needs this rank (its peers would block in all_gather otherwise)
        local = _empty_outputs(tree, xs, evidence_vars, dtype)
    gathered = None
    if gather:
        gathered = all_gather_rows(pack_marginals(local, requested), B, group)
    return local, gathered
