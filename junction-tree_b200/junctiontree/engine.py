"""Device execution of a compiled schedule: memory, streams and stage calls.

PyTorch is used for device memory, streams and (optionally) ``torch.distributed``; all compute
goes through the C ABI of ``libjt_b200.so`` (``_native.py``).  Buffers use the batch-innermost
layout ``[entry, B]`` (see ``include/jt_b200.h``).
"""

import os

import numpy as np

from . import _native
from . import schedule as sch

_torch = None
_have_cuda = False


def torch():
    """Import torch lazily (the host-side compile must work without it being initialised)."""
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise _native.NativeError(
            "no CUDA device: junctiontree (B200 build) has no CPU execution path")
    return t


def current_device():
    """Index of the current CUDA device, -1 without one (part of the engine cache keys: plan
    descriptors and workspaces belong to one device)."""
    global _have_cuda
    t = torch()
    if not _have_cuda:
        _have_cuda = t.cuda.is_available()         # a device does not go away once seen
    return t.cuda.current_device() if _have_cuda else -1


def torch_dtype(dtype):
    t = torch()
    return {np.dtype(np.float64): t.float64, np.dtype(np.float32): t.float32}[np.dtype(dtype)]


class Engine:
    """Runs one :class:`schedule.Plan` on the current CUDA device."""

    def __init__(self, plan):
        self.plan = plan
        self.dev = _native.DevicePlan(plan.to_blob())
        self._workspaces = {}
        #: which binding of the C ABI the stage calls on tensor workspaces go through: "ctypes"
        #: (``_native.py``) or "torch" (the PyTorch extension, ``torch_ops.py``); same library,
        #: same plan handle either way
        self.binding = os.environ.get("JT_BINDING", "ctypes")
        if self.binding not in ("ctypes", "torch"):
            raise ValueError("JT_BINDING must be 'ctypes' or 'torch', got %r" % self.binding)

    # -----------------------------------------------------------------------------------------
    # memory

    def workspace(self, B, dtype):
        """uint8 tensor holding ``[work_entries, B]`` values plus the evidence scratch.

        One workspace per (B, dtype) is cached; callers running several streams allocate their
        own with :meth:`new_workspace`."""
        key = (int(B), np.dtype(dtype).str)
        ws = self._workspaces.pop(key, None)
        if ws is None:
            # at most MAX_CACHED_WORKSPACES batch sizes stay cached (least recently used first out):
            # a caller with ever-changing batch sizes -- propagate_evidence groups instances by
            # evidence pattern -- must not pin one workspace per size it ever used
            plain = [k for k in self._workspaces if isinstance(k[0], int)]
            for stale in plain[:max(0, len(plain) - self.MAX_CACHED_WORKSPACES + 1)]:
                del self._workspaces[stale]
            ws = self.new_workspace(B, dtype)
        self._workspaces[key] = ws              # (re)inserted last: dicts keep insertion order
        return ws

    #: plain (B, dtype) workspaces kept by :meth:`workspace`
    MAX_CACHED_WORKSPACES = 4

    def new_workspace(self, B, dtype):
        t = require_cuda()
        self.dev.upload()
        nbytes = self.dev.workspace_bytes(B, dtype)
        ws = t.empty(nbytes, dtype=t.uint8, device="cuda")
        self.clear_evidence_errors(ws, B, dtype)
        return ws

    #: flags the streaming pipelines run the stages with (which rows a sparse workspace backs)
    PIPELINE_FLAGS = _native.JT_UNIFORM | _native.JT_NO_BELIEFS

    def pipeline_bytes_per_instance(self, dtype, probe=1024):
        """Device bytes a pipeline chunk needs per instance: workspace (sparse when that saves
        memory) plus the output block."""
        self.dev.upload()
        mapped, dense = self.dev.sparse_bytes(probe, dtype, self.PIPELINE_FLAGS)
        return min(mapped, dense) // probe + self.plan.fout_entries * np.dtype(dtype).itemsize

    def new_pipeline_workspace(self, B, dtype):
        """Workspace for a pipeline chunk of ``B`` instances: a sparse one (only the rows the
        uniform, no-beliefs stages touch are backed by memory -- ``jt_workspace_sparse_*``) when
        that saves at least 30 %, else a dense tensor.  Batches of at most 16 instances may run
        as one general-mode launch that touches every row, so they stay dense."""
        self.dev.upload()
        if B > 16:
            mapped, dense = self.dev.sparse_bytes(B, dtype, self.PIPELINE_FLAGS)
            if mapped < 0.7 * dense:
                require_cuda()
                return self.dev.sparse_workspace(B, dtype, self.PIPELINE_FLAGS)
        return self.new_workspace(B, dtype)

    def clear_evidence_errors(self, ws, B, dtype):
        off = self.dev.workspace_layout(B, dtype)["errors"]
        ws[off:off + 256].zero_()

    def release(self):
        self._workspaces.clear()

    def work_view(self, ws, B, dtype):
        """``[work_entries, B]`` typed view of a workspace."""
        n = self.plan.work_entries
        itemsize = np.dtype(dtype).itemsize
        return ws[: n * B * itemsize].view(torch_dtype(dtype)).view(n, B)

    def evidence_offsets_view(self, ws, B, dtype):
        """int32 ``[n_factors, B]`` per-instance factor base offsets written by ``jt_init``
        (workspace layout: include/jt_b200.h)."""
        t = torch()
        start = self.dev.workspace_layout(B, dtype)["fbase"]
        F = len(self.plan.factors)
        return ws[start:start + F * B * 4].view(t.int32).view(F, B)

    def factors_to_device(self, values, dtype, B=None):
        """Concatenate factor tables in plan order.

        ``values[f]`` has the factor's stored shape (shared by the batch) or ``[B, *shape]``
        (per-instance tables; all factors must then be per-instance).
        Returns ``(tensor, batched)``.
        """
        t = require_cuda()
        plan = self.plan
        tdt = torch_dtype(dtype)
        arrays = [np.asarray(v) if not t.is_tensor(v) else v for v in values]
        if len(arrays) != len(plan.factors):
            raise ValueError("expected %d factor arrays, got %d" % (len(plan.factors), len(arrays)))
        batched = [a.ndim == len(shape) + 1 for a, shape in zip(arrays, plan.fin_shape)]
        if any(batched) and not all(batched):
            raise ValueError("either all factor arrays carry a leading batch axis or none does")
        if all(batched) and arrays:
            if B is None:
                B = arrays[0].shape[0]
            flat = t.empty((plan.fin_entries, B), dtype=tdt, device="cuda")
            for f, a in enumerate(arrays):
                if list(a.shape) != [B] + plan.fin_shape[f]:
                    raise ValueError("factor %d: expected shape %s, got %s"
                                     % (f, [B] + plan.fin_shape[f], list(a.shape)))
                dev = t.as_tensor(a).to(device="cuda", dtype=tdt, non_blocking=True)
                flat[plan.fin_off[f]:plan.fin_off[f] + plan.fin_size[f]] = dev.reshape(B, -1).t()
            return flat, True
        host = np.empty(plan.fin_entries, dtype=np.dtype(dtype))
        for f, a in enumerate(arrays):
            if t.is_tensor(a):
                a = a.detach().cpu().numpy()
            if list(a.shape) != plan.fin_shape[f]:
                raise ValueError("factor %d: expected shape %s, got %s"
                                 % (f, plan.fin_shape[f], list(a.shape)))
            host[plan.fin_off[f]:plan.fin_off[f] + plan.fin_size[f]] = a.reshape(-1)
        return t.from_numpy(host).to("cuda"), False

    def likelihoods_host(self, likelihoods, B, dtype, pin=False):
        """Soft evidence ``{variable: [B, size]}`` as one ``[lik_entries, B]`` host tensor in the
        layout of the workspace's likelihood region (batch-innermost)."""
        t = torch()
        plan = self.plan
        given = dict(likelihoods or {})
        if set(given) != set(plan.likelihood_vars):
            raise ValueError("likelihoods must be given for exactly %r" % (plan.likelihood_vars,))
        host = np.empty((plan.lik_entries, B), np.dtype(dtype))
        for k, v in enumerate(plan.likelihood_vars):
            lam = np.asarray(given[v])
            if lam.shape != (B, plan.lik_size[k]):
                raise ValueError("likelihood of %r must have shape [%d, %d], got %s"
                                 % (v, B, plan.lik_size[k], lam.shape))
            host[plan.lik_off[k]:plan.lik_off[k] + plan.lik_size[k]] = lam.T
        out = t.from_numpy(host)
        return out.pin_memory() if pin else out

    def load_likelihoods(self, ws, B, dtype, likelihoods):
        """Fill the likelihood region of a workspace (before ``propagate`` / ``jt_init``)."""
        plan = self.plan
        if not plan.likelihood_vars:
            if likelihoods:
                raise ValueError("this plan was compiled without soft-evidence variables")
            return
        host = self.likelihoods_host(likelihoods, B, dtype)
        self.work_view(ws, B, dtype)[plan.lik_base:plan.lik_base + plan.lik_entries].copy_(host, non_blocking=True)

    def evidence_to_device(self, evidence, B):
        t = require_cuda()
        n_ev = len(self.plan.evidence_vars)
        if n_ev == 0:
            return None
        if evidence is None:
            raise ValueError("the plan has evidence variables %r but no evidence was given"
                             % (self.plan.evidence_vars,))
        if t.is_tensor(evidence):
            ev = evidence.to(device="cuda", dtype=t.int32).contiguous()
        else:
            ev = t.from_numpy(np.ascontiguousarray(evidence, dtype=np.int32)).to("cuda")
        if tuple(ev.shape) != (B, n_ev):
            raise ValueError("evidence must have shape [%d, %d], got %s" % (B, n_ev, tuple(ev.shape)))
        return ev

    # -----------------------------------------------------------------------------------------
    # stages

    @staticmethod
    def _stream():
        return torch().cuda.current_stream().cuda_stream

    def propagate(self, factor_dev, batched, evidence_dev, B, dtype, ws=None, sep_beliefs=False,
                  marginal=True, uniform=True, beliefs=True, semiring=0, dense=True):
        """init + collect + distribute (+ marginal).  Returns ``(ws, factor_out)``; ``factor_out``
        is a ``[fout_entries, B]`` tensor (``None`` when ``marginal`` is False).  ``uniform``:
        compute potentials and messages no evidence reaches once per batch (shared tables only;
        results are identical).  ``semiring``: a ``JT_SR_*`` flag (``semirings.py``).  ``dense``:
        in uniform mode, run tasks with a shared potential and one per-instance message as dense
        contractions on the FP64 tensor pipe (``jt_dense.cu``; same results to rounding)."""
        t = require_cuda()
        self.dev.upload()
        if ws is None:
            ws = self.workspace(B, dtype)
        fout = None
        flags = (_native.JT_SEP_BELIEFS if sep_beliefs else 0) | (0 if uniform else _native.JT_NO_UNIFORM)
        flags |= semiring | (0 if dense else _native.JT_NO_DENSE)
        if not beliefs:       # only the outputs are wanted: no clique belief is written
            flags |= _native.JT_NO_BELIEFS
        if marginal:
            fout = t.empty((self.plan.fout_entries, B), dtype=torch_dtype(dtype), device="cuda")
        else:
            flags |= _native.JT_SKIP_MARGINAL
        if self.binding == "torch" and t.is_tensor(ws):
            from . import torch_ops
            torch_ops.ops().propagate(torch_ops.handle_of(self.dev), factor_dev, bool(batched), evidence_dev, ws,
                                      fout, B, flags)
            return ws, fout
        self.dev.propagate(factor_dev.data_ptr(), batched,
                           evidence_dev.data_ptr() if evidence_dev is not None else None,
                           B, dtype, ws.data_ptr(), fout.data_ptr() if fout is not None else None,
                           flags, self._stream())
        return ws, fout

    def beliefs_from_potentials(self, work, B, dtype, ws, sep_beliefs=True, semiring=0):
        """collect + distribute on clique potentials already stored in the workspace."""
        self.dev.upload()
        sep = _native.JT_SEP_BELIEFS if sep_beliefs else 0
        if self.binding == "torch" and torch().is_tensor(ws):
            from . import torch_ops
            handle, tdt = torch_ops.handle_of(self.dev), torch_dtype(dtype)
            torch_ops.ops().collect(handle, ws, B, tdt, semiring)
            torch_ops.ops().distribute(handle, ws, B, tdt, sep | semiring)
            return
        stream = self._stream()
        self.dev.collect(B, dtype, ws.data_ptr(), semiring, stream)
        self.dev.distribute(B, dtype, ws.data_ptr(), sep | semiring, stream)

    # -----------------------------------------------------------------------------------------
    # views of results

    def node_tensor(self, ws, node, B, dtype):
        """Node ``node`` (clique or separator belief) as a ``[B, *shape]`` view."""
        work = self.work_view(ws, B, dtype)
        off, n = self.plan.node_off[node], self.plan.node_size[node]
        shape = tuple(self.plan.node_shape[node])
        return work[off:off + n].view(shape + (B,)).movedim(-1, 0)

    def factor_tensor(self, fout, f, B):
        off, n = self.plan.fout_off[f], self.plan.fout_size[f]
        shape = tuple(self.plan.fout_shape[f])
        return fout[off:off + n].view(shape + (B,)).movedim(-1, 0)


class BatchPipeline:
    """End-to-end streaming of a large batch in chunks over several CUDA streams.

    Per chunk: pinned int32 evidence -> device, ``jt_propagate`` (init, collect, distribute,
    marginal), per-factor beliefs -> pinned host.  Chunks alternate between ``n_streams``
    streams, each with its own workspace, so the PCIe copies of one chunk overlap the kernels of
    another.  The host result keeps the batch-innermost layout ``[fout_entries, B]``; per-factor
    ``[B, *shape]`` arrays are strided views of it (``factor_views``), so no transpose is done on
    either side.
    """

    def __init__(self, engine, B, dtype, chunk=8192, n_streams=2, normalize=False, log_z=False, semiring=0):
        t = require_cuda()
        self.engine, self.B, self.dtype = engine, int(B), np.dtype(dtype)
        self.normalize, self.log_z, self.semiring = bool(normalize), bool(log_z), int(semiring)
        self.host_logz = t.empty(self.B, dtype=torch_dtype(dtype)).pin_memory() if log_z else None
        plan = engine.plan
        self.chunk = int(min(chunk, B))
        self._primed = set()             # slots whose uniform workspace holds the current factor tables
        self.bounds = [(lo, min(lo + self.chunk, self.B)) for lo in range(0, self.B, self.chunk)]
        n_streams = max(1, min(n_streams, len(self.bounds)))
        engine.dev.upload()
        self.n_ev = len(plan.evidence_vars)
        self.slots = []
        for _ in range(n_streams):
            slot = {
                "stream": t.cuda.Stream(),
                "ws": engine.new_pipeline_workspace(self.chunk, self.dtype),
                "fout": t.empty((plan.fout_entries, self.chunk), dtype=torch_dtype(self.dtype), device="cuda"),
                "ev": t.empty((self.chunk, max(self.n_ev, 1)), dtype=t.int32, device="cuda"),
                "logz": t.empty(self.chunk, dtype=torch_dtype(self.dtype), device="cuda"),
            }
            self.slots.append(slot)
        # a ragged last chunk needs buffers of its own pitch
        self.tail = None
        last = self.bounds[-1][1] - self.bounds[-1][0]
        if last != self.chunk:
            self.tail = {
                "stream": self.slots[(len(self.bounds) - 1) % n_streams]["stream"],
                "ws": engine.new_pipeline_workspace(last, self.dtype),
                "fout": t.empty((plan.fout_entries, last), dtype=torch_dtype(self.dtype), device="cuda"),
                "ev": t.empty((last, max(self.n_ev, 1)), dtype=t.int32, device="cuda"),
                "logz": t.empty(last, dtype=torch_dtype(self.dtype), device="cuda"),
            }

    def host_output(self):
        """Pinned ``[fout_entries, B]`` host buffer for :meth:`run`."""
        t = torch()
        return t.empty((self.engine.plan.fout_entries, self.B), dtype=torch_dtype(self.dtype)).pin_memory()

    def factor_views(self, out_host):
        """Per-factor ``[B, *shape]`` NumPy views of the host buffer."""
        plan = self.engine.plan
        arr = out_host.numpy()
        return [
            np.moveaxis(arr[plan.fout_off[f]:plan.fout_off[f] + plan.fout_size[f]]
                        .reshape(tuple(plan.fout_shape[f]) + (self.B,)), -1, 0)
            for f in range(len(plan.fout_off))
        ]

    def run(self, factor_dev, batched, ev_host, out_host, sync=False, lik_host=None, same_tables=False):
        """Enqueue the whole batch.  ``ev_host``: pinned int32 ``[B, |E|]`` (or None);
        ``out_host``: tensor from :meth:`host_output`; ``lik_host``: pinned ``[lik_entries, B]``
        soft evidence (``Engine.likelihoods_host``) when the plan has likelihood variables.
        ``same_tables``: the factor tables are those of the previous ``run`` of this pipeline, so
        the uniform workspaces are still valid and are not recomputed.  The calling stream waits
        for all chunks."""
        t = torch()
        plan, dev = self.engine.plan, self.engine.dev
        if batched:
            raise ValueError("per-instance factor tables are not streamed; use Engine.propagate")
        if (lik_host is None) != (plan.lik_entries == 0):
            raise ValueError("soft evidence: the plan has %d likelihood entries" % plan.lik_entries)
        cur = t.cuda.current_stream()
        item = self.dtype.itemsize
        for slot in self.slots:
            slot["stream"].wait_stream(cur)
        # the factor tables are the same for every chunk of this call: the uniform workspace of
        # a slot is computed by the first chunk that uses the slot and reused by the later ones
        primed = self._primed if same_tables else set()
        self._primed = primed
        for i, (lo, hi) in enumerate(self.bounds):
            n = hi - lo
            slot = self.tail if (self.tail is not None and n != self.chunk) else self.slots[i % len(self.slots)]
            stream = slot["stream"]
            with t.cuda.stream(stream):
                ev_ptr = None
                if self.n_ev:
                    slot["ev"].copy_(ev_host[lo:hi], non_blocking=True)
                    ev_ptr = slot["ev"].data_ptr()
                if lik_host is not None:          # chunk columns of the likelihood rows -> workspace region
                    _native.copy_rows(slot["ws"].data_ptr() + plan.lik_base * n * item, n * item,
                                      lik_host.data_ptr() + lo * item, self.B * item, n * item, plan.lik_entries,
                                      False, stream.cuda_stream)
                # host output only: clique beliefs are never read, so they are not written
                flags = _native.JT_NO_BELIEFS | (_native.JT_UNIFORM_VALID if id(slot) in primed else 0)
                flags |= self.semiring
                primed.add(id(slot))
                dev.propagate(factor_dev.data_ptr(), False, ev_ptr, n, self.dtype, slot["ws"].data_ptr(),
                              slot["fout"].data_ptr(), flags, stream.cuda_stream)
                if self.normalize or self.log_z:
                    # output stage: per-scope normalisation and log Z = log P(evidence)
                    # (log Z alone: the same kernel with JT_LOGZ_ONLY leaves the outputs unnormalised)
                    dev.normalize(n, self.dtype, slot["fout"].data_ptr(),
                                  slot["logz"].data_ptr() if self.log_z else None, stream.cuda_stream,
                                  self.semiring | (0 if self.normalize else _native.JT_LOGZ_ONLY))
                    if self.log_z:
                        self.host_logz[lo:hi].copy_(slot["logz"], non_blocking=True)
                _native.copy_rows(out_host.data_ptr() + lo * item, self.B * item, slot["fout"].data_ptr(),
                                  n * item, n * item, plan.fout_entries, True, stream.cuda_stream)
        for slot in self.slots:
            cur.wait_stream(slot["stream"])
        if sync:
            cur.synchronize()

    def evidence_errors(self):
        total = 0
        for slot in self.slots + ([self.tail] if self.tail else []):
            n = slot["fout"].shape[1]
            total += self.engine.dev.evidence_errors(n, self.dtype, slot["ws"].data_ptr(),
                                                     slot["stream"].cuda_stream)
        return total


def _pipeline(self, B, dtype, chunk=8192, n_streams=2, normalize=False, log_z=False, semiring=0):
    return BatchPipeline(self, B, dtype, chunk, n_streams, normalize, log_z, semiring)


Engine.pipeline = _pipeline


class GraphedPropagation:
    """One ``jt_propagate`` over static buffers, captured in a CUDA graph.

    For small batches the propagation is launch-bound (config 1: ~10 launches of a few
    microseconds each); replaying a graph that also contains the host->device copy of the
    factor tables and the device->host copy of the result removes the per-launch CPU cost.
    The library never synchronises or allocates in the stage calls, so they capture as is.
    """

    def __init__(self, engine, B, dtype, sep_beliefs=False, uniform=True, semiring=0):
        t = require_cuda()
        plan = engine.plan
        engine.dev.upload()
        self.engine, self.B, self.dtype = engine, int(B), np.dtype(dtype)
        tdt = torch_dtype(dtype)
        n_ev = len(plan.evidence_vars)
        self.host_factors = t.zeros(max(plan.fin_entries, 1), dtype=tdt).pin_memory()
        self.host_evidence = t.zeros((self.B, max(n_ev, 1)), dtype=t.int32).pin_memory()
        self.host_out = t.zeros((max(plan.fout_entries, 1), self.B), dtype=tdt).pin_memory()
        self.factors = t.zeros_like(self.host_factors, device="cuda")
        self.evidence = t.zeros_like(self.host_evidence, device="cuda")
        self.fout = t.zeros_like(self.host_out, device="cuda")
        self.ws = engine.new_workspace(self.B, dtype)
        flags = (_native.JT_SEP_BELIEFS if sep_beliefs else 0) | (0 if uniform else _native.JT_NO_UNIFORM)
        flags |= _native.JT_NO_BELIEFS | int(semiring)
        ev_ptr = self.evidence.data_ptr() if n_ev else None

        def enqueue(stream):
            self.factors.copy_(self.host_factors, non_blocking=True)
            if n_ev:
                self.evidence.copy_(self.host_evidence, non_blocking=True)
            engine.dev.propagate(self.factors.data_ptr(), False, ev_ptr, self.B, self.dtype,
                                 self.ws.data_ptr(), self.fout.data_ptr(), flags, stream.cuda_stream)
            self.host_out.copy_(self.fout, non_blocking=True)

        self.stream = t.cuda.Stream()
        self.stream.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(self.stream):
            enqueue(self.stream)                      # warm-up: module load, kernel attributes
        self.stream.synchronize()
        self.graph = t.cuda.CUDAGraph()
        with t.cuda.graph(self.graph, stream=self.stream):
            enqueue(t.cuda.current_stream())

    def run(self):
        """Replay and wait; the result is in ``host_out`` (``[fout_entries, B]``)."""
        t = torch()
        with t.cuda.stream(self.stream):              # replay runs on the current stream
            self.graph.replay()
        self.stream.synchronize()
        return self.host_out

    def set_factors(self, values):
        """Copy the factor tables (plan order, stored shapes) into the pinned staging buffer."""
        plan = self.engine.plan
        host = self.host_factors.numpy()
        for f, v in enumerate(values):
            a = np.asarray(v)
            if list(a.shape) != plan.fin_shape[f]:
                raise ValueError("factor %d: expected shape %s, got %s" % (f, plan.fin_shape[f], list(a.shape)))
            host[plan.fin_off[f]:plan.fin_off[f] + plan.fin_size[f]] = a.reshape(-1)


class HostPropagation:
    """Host-to-host propagation of a few instances of a small tree in ONE library call
    (``jt_propagate_host``): tables host -> device, the whole-propagation kernel, outputs device
    -> host, stream synchronised.  For such trees a propagation is launch latency end to end
    (config 1: one kernel of a few microseconds), so the call path matters more than the
    kernel: no CUDA graph, no per-call torch stream handling, one ``ctypes`` call.  Same
    staging buffers and ``set_factors`` interface as :class:`GraphedPropagation`."""

    def __init__(self, engine, B, dtype, semiring=0):
        t = require_cuda()
        plan = engine.plan
        engine.dev.upload()
        self.engine, self.B, self.dtype = engine, int(B), np.dtype(dtype)
        tdt = torch_dtype(dtype)
        n_ev = len(plan.evidence_vars)
        self.host_factors = t.zeros(max(plan.fin_entries, 1), dtype=tdt).pin_memory()
        self.host_evidence = t.zeros((self.B, max(n_ev, 1)), dtype=t.int32).pin_memory()
        self.host_out = t.zeros((max(plan.fout_entries, 1), self.B), dtype=tdt).pin_memory()
        self.factors = t.zeros_like(self.host_factors, device="cuda")
        self.evidence = t.zeros_like(self.host_evidence, device="cuda")
        self.fout = t.zeros_like(self.host_out, device="cuda")
        self.ws = engine.new_workspace(self.B, dtype)
        self.stream = t.cuda.Stream()
        self.flags = _native.JT_NO_BELIEFS | int(semiring)
        t.cuda.synchronize()                            # staging buffers are zero-filled on the default stream
        item = self.dtype.itemsize
        self._call = (self.host_factors.data_ptr(), plan.fin_entries * item,
                      self.host_evidence.data_ptr() if n_ev else None, self.B, self.dtype,
                      self.factors.data_ptr(), self.evidence.data_ptr() if n_ev else None, self.ws.data_ptr(),
                      self.fout.data_ptr(), self.host_out.data_ptr(), plan.fout_entries * self.B * item, self.flags,
                      self.stream.cuda_stream)
        self._host_factors_np = self.host_factors.numpy()
        self._host_out_np = self.host_out.numpy()
        self._slots = [(plan.fin_off[f], plan.fin_off[f] + plan.fin_size[f], tuple(plan.fin_shape[f]))
                       for f in range(len(plan.fin_off))]

    def set_factors(self, values):
        """Copy the factor tables (plan order, stored shapes) into the pinned staging buffer."""
        host = self._host_factors_np
        for (lo, hi, shape), v in zip(self._slots, values):
            a = v if isinstance(v, np.ndarray) else np.asarray(v)
            if a.shape != shape:
                raise ValueError("factor table: expected shape %s, got %s" % (list(shape), list(a.shape)))
            host[lo:hi] = a.reshape(-1)

    def run(self):
        """Propagate and wait; the result is in ``host_out`` (``[fout_entries, B]``)."""
        self.engine.dev.propagate_host(*self._call)
        return self.host_out


def _host_runner(self, B, dtype, semiring=0):
    """Runner for host-in / host-out propagation of ``B`` instances: one library call when the
    plan runs as a single launch, a CUDA-graph replay otherwise."""
    flags = _native.JT_NO_BELIEFS | int(semiring)
    self.dev.upload()
    if self.dev.single_launch(int(B), flags):
        key = ("host", int(B), np.dtype(dtype).str, int(semiring))
        hit = self._workspaces.get(key)
        if hit is None:
            hit = HostPropagation(self, B, dtype, semiring)
            self._workspaces[key] = hit
        return hit
    return self.graphed(B, dtype, semiring=semiring)


Engine.host_runner = _host_runner


def _graphed(self, B, dtype, sep_beliefs=False, uniform=True, semiring=0):
    key = ("graph", int(B), np.dtype(dtype).str, bool(sep_beliefs), bool(uniform), int(semiring))
    hit = self._workspaces.get(key)
    if hit is None:
        hit = GraphedPropagation(self, B, dtype, sep_beliefs, uniform, semiring)
        self._workspaces[key] = hit
    return hit


Engine.graphed = _graphed
