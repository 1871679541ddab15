"""Benchmark: batched junction-tree propagations/sec on B200 (contract: see DESIGN.md section 6).

    python bench.py --gpus 1 --steps 10 --warmup 3              # our arm (sm_100a kernels)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1   # CPU reference arm
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W        # N > 1, weak scaling

A step is one pass of the hot path (evidence slicing + clique initialisation + collect +
distribute, all clique and separator beliefs written) over one batch of synthetic evidence for
BASELINE.json configs[1]: the 37-node random DAG, 65,536 instances per GPU, float64.  One JSON
line is printed by rank 0.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "junction-tree_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import jt_workloads as wl  # noqa: E402

METRIC = "batched propagations/sec (collect+distribute)"
UNIT = "propagations/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(args, uniform):
    """DRAM bytes per launch of the dominant kernel from the committed ncu launch list
    (profiles/r01_traffic.json), when one exists for this exact configuration."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            table = json.load(fh)
        key = "%s:%d:%s:%s" % (args.config, args.batch, args.dtype, "uniform" if uniform else "per_instance")
        return table[key]["traffic_per_launch"] if args.semiring == "sum_product" else None
    except Exception:
        return None


def make_net(name):
    if name == "dag37":
        return wl.dag37()
    if name == "dag500":
        return wl.dag500()
    if name == "ising16":
        return wl.ising(16)
    if name == "large_state_tree":
        return wl.large_state_tree()
    if name == "sprinkler":
        return wl.sprinkler()
    raise SystemExit("unknown --config %s" % name)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        while self.ok and not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.005)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# CPU arm (the oracle port of the reference's NumPy path; the reference itself cannot build the
# valid trees these configs need -- SURVEY.md 0.4 -- and does not travel to the GPU box)

_CPU = {}


def _cpu_setup(name, semiring="sum_product"):
    from junctiontree import construction as cons
    net = make_net(name)
    if semiring in ("log_sum_exp", "max_sum"):
        net["values"] = [np.log(v) for v in net["values"]]
    _CPU["semiring"] = semiring
    _, mc, f2c = cons.find_triangulation(net["factors"], net["sizes"], net.get("order"))
    tree, seps = cons.construct_junction_tree(mc, net["sizes"])
    _CPU.update(net=net, mc=mc, f2c=f2c, tree=tree, seps=seps)


def _cpu_work(ev_rows):
    from oracle import ref_fixed
    net = _CPU["net"]
    outs, _ = ref_fixed.propagate_batch(_CPU["tree"], _CPU["seps"], _CPU["mc"], _CPU["f2c"], net["factors"],
                                        net["sizes"], net["values"], net.get("evidence_vars", []), ev_rows,
                                        n=len(ev_rows), semiring=_CPU["semiring"])
    return float(sum(o.sum() for o in outs))


def cpu_throughput(name, per_core, repeats=1, pool=None, semiring="sum_product"):
    """props/s of oracle/ref_fixed.py over all host cores on a bounded sample."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    net = make_net(name)
    own_pool = pool is None
    if own_pool:
        pool = mp.get_context("fork").Pool(cores, initializer=_cpu_setup, initargs=(name, semiring))
    try:
        n = per_core * cores
        ev = wl.draw_evidence(net, n) if net.get("evidence_vars") else np.zeros((n, 0), np.int32)
        chunks = [ev[i::cores] for i in range(cores)]
        pool.map(_cpu_work, [c[:1] for c in chunks])          # warm-up: imports, first einsum paths
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_work, chunks)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    finally:
        if own_pool:
            pool.close()
            pool.join()
    return n / best, cores, n, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_setup, initargs=(args.config, args.semiring))
    try:
        for _ in range(args.warmup):
            cpu_throughput(args.config, 1, pool=pool)
        times, n = [], 0
        for _ in range(args.steps):
            value, cores, n, dt = cpu_throughput(args.config, args.cpu_per_core, pool=pool)
            times.append(dt)
    finally:
        pool.close()
        pool.join()
    ms = 1e3 * float(np.mean(times))
    value = n / (ms / 1e3)
    sample = "%d instances per step (%d per core) of %s through oracle/ref_fixed.py, %d processes" % (
        n, args.cpu_per_core, args.config, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_line(json.dumps(line))


def workload_name(args):
    return "configs[1]: random 37-node DAG (<=4 parents, 2-4 states, window 8, seed 0), evidence on the " \
           "8 highest-numbered variables, %d instances per GPU" % args.batch if args.config == "dag37" \
        else "%s, %d instances per GPU" % (args.config, args.batch)


# ---------------------------------------------------------------------------------------------
# GPU arm


def run_gpu(args):
    import torch
    import junctiontree as jt
    from junctiontree import _native, distributed as jdist

    rank, world = jdist.init_from_env("nccl")
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    import torch.distributed as dist
    dev = torch.cuda.current_device()
    dtype = np.dtype(np.float64 if args.dtype == "f64" else np.float32)
    w = dtype.itemsize

    net = make_net(args.config)
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = [] if args.no_evidence else list(net.get("evidence_vars", []))
    plan = tree.plan(evars)
    engine = tree._engine(plan.sizes, evars, plan.full_sizes)
    B = args.batch                                   # per GPU (weak scaling)
    ev_all = wl.draw_evidence(net, B * world) if evars else None
    lo, hi = jdist.shard_bounds(B * world, world, rank)
    ev_host = torch.from_numpy(ev_all[lo:hi].copy()).pin_memory() if evars else None

    sr_flag = {"sum_product": _native.JT_SR_SUM_PRODUCT, "max_product": _native.JT_SR_MAX_PRODUCT,
               "log_sum_exp": _native.JT_SR_LOG_SUM_EXP, "max_sum": _native.JT_SR_MAX_SUM}[args.semiring]
    if args.semiring in ("log_sum_exp", "max_sum"):          # log-domain laws take log potentials
        net["values"] = [np.log(v) for v in net["values"]]
    fdev, batched = engine.factors_to_device(net["values"], dtype)
    ev_dev = ev_host.to("cuda") if evars else None
    ws = engine.workspace(B, dtype)
    engine.dev.upload()
    fout = torch.empty((plan.fout_entries, B), dtype=engine_dtype(dtype), device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ev_ptr = ev_dev.data_ptr() if evars else None
    flags = _native.JT_SEP_BELIEFS | (0 if args.no_uniform else _native.JT_UNIFORM) | sr_flag

    def hot_path(events=None):
        if events is not None:
            events[0].record()
        engine.dev.init(fdev.data_ptr(), batched, ev_ptr, B, dtype, ws.data_ptr(), flags, stream)
        if events is not None:
            events[1].record()
        engine.dev.collect(B, dtype, ws.data_ptr(), flags, stream)
        engine.dev.distribute(B, dtype, ws.data_ptr(), flags, stream)
        if events is not None:
            events[2].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        hot_path()
    barrier()

    # ---- timed region: exactly K steps, device-timed, max over ranks ----
    sampler = ClockSampler(dev)
    sampler.start()
    launches0 = _native.launch_count()
    ev_pairs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for k in range(args.steps):
        hot_path(ev_pairs[k])
    t_end.record()
    barrier()
    launches = _native.launch_count() - launches0
    clocks = sampler.finish()
    total_ms = t_start.elapsed_time(t_end)
    init_ms = sum(e[0].elapsed_time(e[1]) for e in ev_pairs)
    msg_ms = sum(e[1].elapsed_time(e[2]) for e in ev_pairs)

    # ---- end to end through the public API: pinned host evidence in, per-factor beliefs out ----
    # (--e2e-batch: a larger batch for the streaming pipelines than fits the dense, all-beliefs
    # workspace of the timed region above; their chunks use sparse workspaces)
    B_value = B
    if args.e2e_batch and args.e2e_batch != B:
        del ws, fout
        engine.release()
        torch.cuda.empty_cache()
        B = args.e2e_batch
        ev_e2e = wl.draw_evidence(net, B * world) if evars else None
        lo, hi = jdist.shard_bounds(B * world, world, rank)
        ev_host = torch.from_numpy(ev_e2e[lo:hi].copy()).pin_memory() if evars else None
    # the public serving call: tree.propagate_session(...).run(evidence) -- host evidence in,
    # per-factor beliefs in host memory out (views of the session's pinned buffer)
    law = {"sum_product": None, "max_product": jt.semirings.max_product, "log_sum_exp": jt.semirings.log_sum_exp,
           "max_sum": jt.semirings.max_sum}[args.semiring]
    ev_np = ev_host.numpy() if evars else None
    session = tree.propagate_session(net["values"], B, evars, dtype=dtype, dl=law, chunk=args.chunk)
    for _ in range(2):
        session.run(ev_np, copy=False)
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    l_e2e0 = _native.launch_count()
    t_e2e = time.perf_counter()
    for _ in range(e2e_steps):
        beliefs = session.run(ev_np, copy=False)          # synchronous: the results are on the host on return
    e2e_ms = (time.perf_counter() - t_e2e) * 1e3
    assert len(beliefs) == len(net["factors"]) and beliefs[0].shape[0] == B
    e2e_launches = _native.launch_count() - l_e2e0
    e2e_chunk = session.pipe.chunk
    session.close()
    del session, beliefs

    # ---- the same end to end with the device output stage: normalised single-variable
    # posteriors of the unobserved variables + log P(evidence) instead of raw factor beliefs ----
    free_vars = [v for v in sorted(net["sizes"]) if v not in evars]
    m_session = tree.marginals_session(net["values"], B, free_vars, evars, dtype=dtype, dl=law, chunk=args.chunk)
    for _ in range(2):
        m_session.run(ev_np, copy=False)
    barrier()
    t_marg = time.perf_counter()
    for _ in range(e2e_steps):
        m_session.run(ev_np, copy=False)
    marg_ms = (time.perf_counter() - t_marg) * 1e3
    marg_d2h = int((m_session.engine.plan.fout_entries + 1) * B * w)
    m_session.close()
    del m_session

    def max_over_ranks(x):
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    total_ms, init_ms, msg_ms, e2e_ms, marg_ms = (max_over_ranks(x) for x in
                                                  (total_ms, init_ms, msg_ms, e2e_ms, marg_ms))
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    B_e2e, B = B, B_value
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step / 1e3)
    peak, peak_src = load_peaks()
    # algorithmic bytes: A = w(4 sum n_C - n_root + 6 sum n_S) per instance;
    # the message-passing kernel moves everything but the init write (w * sum n_C)
    A = w * plan.algorithmic_entries(with_init=True)
    A_msg = w * plan.algorithmic_entries(with_init=False)
    uniform = not args.no_uniform and plan.uni_entries > 0
    from junctiontree import schedule as sch
    # batch launches of collect + distribute in the mode that ran
    msg_phases = (sch.PHASE_COLLECT_INSTANCE, sch.PHASE_DIST_PRE_INSTANCE, sch.PHASE_DIST_MAIN) if uniform \
        else (sch.PHASE_COLLECT, sch.PHASE_DIST_PRE, sch.PHASE_DIST_MAIN)
    msg_launches = sum(1 for L in plan.launches_arr if L[0] in msg_phases)
    msg_ms_per_launch = msg_ms / args.steps / max(msg_launches, 1)
    achieved = A_msg * B / (msg_ms / args.steps / 1e3) / 1e9
    step_gbs = A * B / (ms_per_step / 1e3) / 1e9
    # bytes this schedule has to move through HBM (uniform operands are read once, not per instance)
    S_all = w * plan.scheduled_entries(uniform=uniform)
    n_init = sum(plan.node_size[c] for c in range(plan.n_cliques) if not (uniform and plan.uniform[c]))
    S_msg = S_all - w * n_init
    scheduled = S_msg * B / (msg_ms / args.steps / 1e3) / 1e9

    cpu = None
    if not args.skip_cpu and world == 1:          # the CPU baseline is reported by the single-GPU run only
        v, cores, n, dt = cpu_throughput(args.config, args.cpu_per_core, semiring=args.semiring)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d instances of the same workload (%d per core, %.1f s) through oracle/ref_fixed.py "
                         "(NumPy restatement of the reference), %d processes" % (n, args.cpu_per_core, dt, cores)}

    e2e_value = B_e2e * world / (e2e_ms / e2e_steps / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {
            "workload": workload_name(args), "batch_per_gpu": B, "global_batch": B * world,
            "parallelism": "batch-sharded x%d, no data-path collective" % world,
            "cliques": plan.n_cliques, "clique_entries": plan.clique_entries, "sep_entries": plan.sep_entries,
            "levels": plan.max_depth, "algorithmic_bytes_per_propagation": A,
            "scheduled_bytes_per_propagation": S_all, "uniform_mode": bool(uniform),
            "uniform_clique_entries": plan.uni_entries if uniform else 0, "semiring": args.semiring,
            "step": "evidence slicing + clique init + collect + distribute (clique and separator beliefs)",
            "l2": "inputs larger than L2 (working set %.1f GB per GPU)" % (plan.work_entries * B * w / 1e9),
            "step_gbs_algorithmic": step_gbs, "init_ms_per_step": init_ms / args.steps,
            "message_passing_ms_per_step": msg_ms / args.steps,
        },
        "roofline": {
            "bound": "hbm", "kernel": "jt_project_tma_kernel<%s, %s>" % (
                {"sum_product": "SrSumProduct", "max_product": "SrMaxProduct", "log_sum_exp": "SrLogSumExp",
                 "max_sum": "SrMaxSum"}[args.semiring], "double" if w == 8 else "float"),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "traffic": load_traffic(args, uniform),
            "launches_per_step": msg_launches, "avg_launch_ms": msg_ms_per_launch,
            "algorithmic_bytes_per_launch": A_msg * B / max(msg_launches, 1),
            "scheduled_bytes_per_launch": S_msg * B / max(msg_launches, 1),
            "scheduled_gbs": scheduled, "scheduled_frac": scheduled / peak,
            "note": ("achieved/frac use SURVEY.md 8d's algorithmic bytes A (every potential per instance in HBM); "
                     "uniform mode keeps potentials and up-messages that no evidence reaches once per batch, so "
                     "the schedule moves fewer bytes than A and frac can exceed 1 -- scheduled_* count the bytes "
                     "this schedule must move and are the figure to compare with dram traffic")
                    if uniform else "per-instance potentials: scheduled bytes = algorithmic bytes + re-reads of "
                                    "psi_C by cliques with several children",
        },
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(ev_host.numel() * 4) if evars else 0,
                "d2h_bytes_per_step": int(plan.fout_entries * B_e2e * w),
                "ms_per_step": e2e_ms / e2e_steps, "chunk": e2e_chunk, "batch_per_gpu": B_e2e,
                "what": "tree.propagate_session(values, batch, evidence_vars).run(evidence): host int32 "
                        "evidence -> device, propagate incl. marginalisation to factor scopes, per-factor "
                        "beliefs -> host memory; wall clock around the synchronous calls"},
        "gpu_launches": int(launches),
        "gpu_launches_e2e": int(e2e_launches),
        "e2e_marginals": {"value": B_e2e * world / (marg_ms / e2e_steps / 1e3), "unit": UNIT,
                          "ms_per_step": marg_ms / e2e_steps,
                          "h2d_bytes_per_step": int(ev_host.numel() * 4) if evars else 0,
                          "d2h_bytes_per_step": marg_d2h,
                          "what": "tree.marginals_session(...).run(evidence): same propagation, device output "
                                  "stage (normalised single-variable posteriors + log Z) -> host memory"},
        "clocks": clocks,
    }
    print_line(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def engine_dtype(dtype):
    from junctiontree import engine as eng
    return eng.torch_dtype(dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="dag37")
    ap.add_argument("--batch", type=int, default=65536, help="instances per GPU")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--chunk", type=int, default=8192, help="instances per pipeline chunk (e2e)")
    ap.add_argument("--e2e-batch", type=int, default=0,
                    help="instances per GPU for the end-to-end pipelines (default: --batch)")
    ap.add_argument("--no-evidence", action="store_true")
    ap.add_argument("--no-uniform", action="store_true",
                    help="materialise every potential and message per instance (general path)")
    ap.add_argument("--semiring", default="sum_product",
                    choices=["sum_product", "max_product", "log_sum_exp", "max_sum"],
                    help="distributive law of the kernels (default: the reference's sum-product)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-per-core", type=int, default=512, help="CPU baseline: instances per core")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything libraries print there meanwhile (e.g. NCCL's
    # version banner) goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(saved, "w")
    global print_line
    print_line = lambda text: (out.write(text + "\n"), out.flush())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


def print_line(text):
    print(text)


if __name__ == "__main__":
    main()
