"""The timed hot path of one BASELINE.json config, shared by ``bench.py`` (all five configs in the
``configs`` block of its JSON line) and ``tools/prof_step.py`` (the command the ncu launch lists
under ``profiles/`` are taken over), so both measure the same steps.

A step = evidence slicing + clique initialisation + collect + distribute (clique and separator
beliefs written per instance), i.e. what SURVEY.md 8d's algorithmic bytes A count; with
``beliefs=False`` it is the pipelines' mode instead: messages + per-factor outputs, no clique
belief stored (sparse workspace when that saves memory).
"""

import os

import numpy as np

import jt_workloads as wl


def make_net(name):
    nets = {"dag37": wl.dag37, "dag500": wl.dag500, "ising16": lambda: wl.ising(16),
            "large_state_tree": wl.large_state_tree, "sprinkler": wl.sprinkler}
    if name not in nets:
        raise SystemExit("unknown config %s" % name)
    return nets[name]()


def largest_batch(config, dtype, cap, step=256, fraction=0.88):
    """Largest multiple of ``step`` <= ``cap`` whose dense workspace (all beliefs stored) fits in
    ``fraction`` of the free device memory -- config 5 takes 80 MB per instance, so the chunk size
    is whatever the GPU holds."""
    import torch
    import junctiontree as jt
    net = make_net(config)
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = list(net.get("evidence_vars", []))
    plan = tree.plan(evars)
    engine = tree._engine(plan.sizes, evars, plan.full_sizes)
    per = engine.dev.workspace_bytes(step, np.dtype(np.float64 if dtype == "f64" else np.float32)) / step
    free, _ = torch.cuda.mem_get_info()
    tree.clique_tree._engines.clear()
    return int(max(step, min(cap, int(fraction * free / per) // step * step)))


SR_KERNEL = {"sum_product": "SrSumProduct", "max_product": "SrMaxProduct", "log_sum_exp": "SrLogSumExp",
             "max_sum": "SrMaxSum"}


class HotPath:
    """Plan, engine and device buffers of one (config, batch, dtype, mode); ``step()`` enqueues
    one pass on the current stream."""

    def __init__(self, config, batch, dtype="f64", uniform=True, evidence=True, beliefs=True, dense=True,
                 semiring="sum_product", net=None, ev_offset=0, ev_total=None):
        import torch
        import junctiontree as jt
        from junctiontree import _native
        self.torch, self._native = torch, _native
        self.config, self.B = config, int(batch)
        self.dtype = np.dtype(np.float64 if dtype == "f64" else np.float32)
        self.dtype_name = dtype
        self.semiring = semiring
        self.beliefs, self.dense = bool(beliefs), bool(dense)
        net = dict(net if net is not None else make_net(config))
        if semiring in ("log_sum_exp", "max_sum"):          # log-domain laws take log potentials
            net["values"] = [np.log(v) for v in net["values"]]
        self.net = net
        self.tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
        self.evars = list(net.get("evidence_vars", [])) if evidence else []
        self.plan = self.tree.plan(self.evars)
        self.engine = self.tree._engine(self.plan.sizes, self.evars, self.plan.full_sizes)
        self.uniform = bool(uniform) and self.plan.uni_entries > 0
        self.fdev, self.batched = self.engine.factors_to_device(net["values"], self.dtype)
        self.ev_host = None
        if self.evars:
            total = ev_total if ev_total is not None else self.B
            ev_all = wl.draw_evidence(net, total)
            self.ev_host = torch.from_numpy(ev_all[ev_offset:ev_offset + self.B].copy()).pin_memory()
        self.ev_dev = self.ev_host.to("cuda") if self.evars else None
        self.engine.dev.upload()
        if not self.beliefs and self.uniform:
            self.ws = self.engine.new_pipeline_workspace(self.B, self.dtype)     # sparse when that saves memory
        else:
            self.ws = self.engine.new_workspace(self.B, self.dtype)
        self.sparse = not hasattr(self.ws, "numel")
        self.fout = None
        if not self.beliefs:
            self.fout = torch.empty((self.plan.fout_entries, self.B), dtype=torch.float64 if self.dtype.itemsize == 8
                                    else torch.float32, device="cuda")
        sr_flag = {"sum_product": _native.JT_SR_SUM_PRODUCT, "max_product": _native.JT_SR_MAX_PRODUCT,
                   "log_sum_exp": _native.JT_SR_LOG_SUM_EXP, "max_sum": _native.JT_SR_MAX_SUM}[semiring]
        self.flags = (_native.JT_UNIFORM if self.uniform else 0) | sr_flag
        self.flags |= _native.JT_SEP_BELIEFS if self.beliefs else _native.JT_NO_BELIEFS
        self.flags |= 0 if self.dense else _native.JT_NO_DENSE

    def step(self, events=None):
        dev, B, dtype = self.engine.dev, self.B, self.dtype
        stream = self.torch.cuda.current_stream().cuda_stream
        ws = self.ws.data_ptr()
        if events is not None:
            events[0].record()
        dev.init(self.fdev.data_ptr(), self.batched, self.ev_dev.data_ptr() if self.evars else None, B, dtype, ws,
                 self.flags, stream)
        if events is not None:
            events[1].record()
        dev.collect(B, dtype, ws, self.flags, stream)
        dev.distribute(B, dtype, ws, self.flags, stream)
        if not self.beliefs:
            dev.marginal(B, dtype, ws, self.fout.data_ptr(), stream, self.flags)
        if events is not None:
            events[2].record()

    def time(self, steps, warmup, barrier=None):
        """CUDA-event timing of exactly ``steps`` steps after ``warmup`` untimed ones.  Returns ms
        per step in total and split into init / message passing, and the launches per step."""
        torch = self.torch
        sync = barrier or torch.cuda.synchronize
        for _ in range(warmup if os.environ.get("JT_BENCH_SHORT_WARMUP") else max(warmup, 3)):   # (ncu launch lists)
            self.step()
        sync()
        marks = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self._native.launch_count()
        sync()
        t0.record()
        for k in range(steps):
            self.step(marks[k])
        t1.record()
        sync()
        return {"ms_per_step": t0.elapsed_time(t1) / steps,
                "init_ms": sum(m[0].elapsed_time(m[1]) for m in marks) / steps,
                "msg_ms": sum(m[1].elapsed_time(m[2]) for m in marks) / steps,
                "launches_per_step": (self._native.launch_count() - l0) / steps}

    # ---- byte accounting ----

    def bytes_per_propagation(self):
        """(A, A_msg, S, S_msg): SURVEY.md 8d's algorithmic bytes with and without the init write,
        and the bytes this schedule has to move through HBM in the mode that runs (uniform
        operands are read once per batch, not per instance), total and message passing only."""
        w, plan = self.dtype.itemsize, self.plan
        A = w * plan.algorithmic_entries(with_init=True)
        A_msg = w * plan.algorithmic_entries(with_init=False)
        S = w * plan.scheduled_entries(uniform=self.uniform, beliefs=self.beliefs)
        n_init = sum(plan.node_size[c] for c in range(plan.n_cliques) if not (self.uniform and plan.uniform[c]))
        return A, A_msg, S, S - w * n_init

    def msg_launches(self):
        """Batch launches of collect + distribute (+ marginal) of the plan in the mode that runs
        (mixed launches run a dense part and a projection part: counted once here)."""
        from junctiontree import schedule as sch
        main = sch.PHASE_DIST_MAIN if self.beliefs else sch.PHASE_DIST_MAIN_MESSAGES
        phases = (sch.PHASE_COLLECT_INSTANCE, sch.PHASE_DIST_PRE_INSTANCE, main) if self.uniform else \
            (sch.PHASE_COLLECT, sch.PHASE_DIST_PRE, main)
        if not self.beliefs:
            phases += (sch.PHASE_MARGINAL_DIRECT,)
        return sum(1 for L in self.plan.launches_arr if L[0] in phases)

    def working_set_gb(self):
        return self.plan.work_entries * self.B * self.dtype.itemsize / 1e9

    def release(self):
        self.ws = self.fout = self.fdev = self.ev_dev = self.ev_host = None
        self.engine.release()
        self.tree.clique_tree._engines.clear()
        self.torch.cuda.empty_cache()
