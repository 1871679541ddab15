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
    """DRAM bytes per launch of the message-passing kernels as ncu measured them for this exact
    configuration (profiles/r02_traffic.json: derived from the committed launch lists by
    tools/summarize_launches.py --update).  A recorded figure, labelled with its source."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            table = json.load(fh)
        key = "%s:%d:%s:%s" % (args.config, args.batch, args.dtype, "uniform" if uniform else "per_instance")
        if args.semiring != "sum_product" or args.no_evidence or args.no_dense:
            return None
        return table[key]
    except Exception:
        return None


def make_net(name):
    import jt_bench_lib as bl
    return bl.make_net(name)


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


#: BASELINE.json's configs at their stated sizes (per GPU; weak scaling), beside the main line's
#: configs[1] with evidence in uniform mode.  (name, batch, dtype, uniform, evidence, beliefs,
#: sub_batches): config 3's 256 instances run in one 142 GB workspace (as 2 x 128 when that does not
#: fit beside whatever else holds memory on the GPU), config 5 once with all beliefs stored (dense
#: workspace, 1,024 instances) and once in the pipelines' mode on a sparse workspace (4,096
#: instances: outputs only).  Config 5 with all beliefs stored takes 80 MB per instance: its chunk is
#: the largest multiple of 256 up to 2,048 that fits the free memory.
EXTRA_CONFIGS = [
    ("dag37", 65536, "f64", True, False, True, 1),
    ("dag37", 65536, "f64", False, True, True, 1),
    ("ising16", 256, "f64", True, True, True, 1),
    ("ising16", 256, "f64", False, True, True, 1),
    ("large_state_tree", 512, "f64", True, True, True, 1),
    ("large_state_tree", 512, "f64", False, True, True, 1),
    ("large_state_tree", 512, "f32", True, True, True, 1),
    ("large_state_tree", 512, "f32", False, True, True, 1),
    ("dag500", 2048, "f64", True, True, True, 1),       # the largest chunk (<= 2,048) that fits: 80 MB per instance
    ("dag500", 2048, "f64", False, True, True, 1),
    ("dag500", 4096, "f64", True, True, False, 1),
]


def config_entry(hp, t, peak, sub_batches=1):
    A, A_msg, S, S_msg = hp.bytes_per_propagation()
    ms = t["ms_per_step"] * sub_batches
    B = hp.B * sub_batches
    return {
        "config": hp.config, "batch_per_gpu": B, "sub_batches": sub_batches, "dtype": hp.dtype_name,
        "mode": "uniform" if hp.uniform else "per_instance", "evidence": bool(hp.evars),
        "beliefs_stored": hp.beliefs, "sparse_workspace": hp.sparse, "dense_contractions": hp.dense and hp.uniform,
        "ms_per_step": ms, "value": B / ms * 1e3, "unit": UNIT, "launches_per_step": t["launches_per_step"] * sub_batches,
        "scheduled_bytes_per_propagation": S, "algorithmic_bytes_per_propagation": A,
        # bytes the schedule moves / step time / measured peak; the A-based figure can exceed 1 in
        # uniform mode (potentials no evidence reaches are not streamed per instance)
        "frac": S * B / ms / 1e6 / peak, "frac_algorithmic": A * B / ms / 1e6 / peak,
        "cliques": hp.plan.n_cliques, "levels": hp.plan.max_depth,
    }


def single_propagate_latency(jt, net, n=300):
    """Config 1 (BASELINE configs[0]): one `tree.propagate(values)` call, host arrays in and out."""
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    for _ in range(30):
        tree.propagate(net["values"])
    times = []
    for _ in range(n):
        t0 = time.perf_counter()
        tree.propagate(net["values"])
        times.append(time.perf_counter() - t0)
    return {"config": "sprinkler", "batch_per_gpu": 1, "dtype": "f64", "mode": "single propagate() call, host to host",
            "us_per_propagate": 1e6 * float(np.median(times)), "us_per_propagate_p90": 1e6 * float(np.percentile(times, 90)),
            "value": 1.0 / float(np.median(times)), "unit": UNIT, "calls": n,
            "frac": None, "note": "latency-bound: one kernel launch between two small copies (SURVEY.md 8d)"}


def run_gpu(args):
    import torch
    import junctiontree as jt
    from junctiontree import _native, distributed as jdist
    import jt_bench_lib as bl

    rank, world = jdist.init_from_env("nccl")
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    import torch.distributed as dist
    dev = torch.cuda.current_device()
    dtype = np.dtype(np.float64 if args.dtype == "f64" else np.float32)
    w = dtype.itemsize
    peak, peak_src = load_peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(xs):
        if world == 1:
            return list(xs)
        tns = torch.tensor(list(xs), dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return [float(v) for v in tns.tolist()]

    B = args.batch                                   # per GPU (weak scaling)
    hp = bl.HotPath(args.config, B, args.dtype, uniform=not args.no_uniform, evidence=not args.no_evidence,
                    dense=not args.no_dense, semiring=args.semiring, ev_offset=rank * B, ev_total=B * world)
    net, tree, plan, evars = hp.net, hp.tree, hp.plan, hp.evars

    # ---- timed region: exactly K steps, device-timed, max over ranks ----
    sampler = ClockSampler(dev)
    sampler.start()
    t_main = hp.time(args.steps, args.warmup, barrier)
    clocks = sampler.finish()
    launches = int(round(t_main["launches_per_step"] * args.steps))
    working_set_gb = hp.working_set_gb()
    main_entry = config_entry(hp, t_main, peak)
    A, A_msg, S_all, S_msg = hp.bytes_per_propagation()
    uniform = hp.uniform
    msg_launches = hp.msg_launches()
    ev_host = hp.ev_host
    net_values = net["values"]
    hp.release()
    del hp

    # ---- end to end through the public API: pinned host evidence in, per-factor beliefs out ----
    B_e2e = args.e2e_batch or B
    if B_e2e != B:
        ev_e2e = wl.draw_evidence(net, B_e2e * world) if evars else None
        lo, hi = jdist.shard_bounds(B_e2e * world, world, rank)
        ev_host = torch.from_numpy(ev_e2e[lo:hi].copy()).pin_memory() if evars else None
    law = {"sum_product": None, "max_product": jt.semirings.max_product, "log_sum_exp": jt.semirings.log_sum_exp,
           "max_sum": jt.semirings.max_sum}[args.semiring]
    ev_np = ev_host.numpy() if evars else None
    e2e_steps = max(2, min(args.steps, 5))
    e2e_ms = marg_ms = 0.0
    e2e_launches = e2e_chunk = marg_d2h = 0
    d2h = {}
    if not args.skip_e2e:
        # the public serving call: tree.propagate_session(...).run(evidence) -- host evidence in,
        # per-factor beliefs in host memory out (views of the session's pinned buffer)
        session = tree.propagate_session(net_values, B_e2e, evars, dtype=dtype, dl=law, chunk=args.chunk)
        for _ in range(2):
            session.run(ev_np, copy=False)
        barrier()
        l_e2e0 = _native.launch_count()
        t_e2e = time.perf_counter()
        for _ in range(e2e_steps):
            beliefs = session.run(ev_np, copy=False)          # synchronous: the results are on the host on return
        e2e_ms = (time.perf_counter() - t_e2e) * 1e3
        assert len(beliefs) == len(net["factors"]) and beliefs[0].shape[0] == B_e2e
        e2e_launches = _native.launch_count() - l_e2e0
        e2e_chunk = session.pipe.chunk
        # the host ceiling of that call: the same bytes as one plain device -> pinned host copy,
        # all ranks at once (the copy engines share the host's memory system)
        out_host = session.out_host
        probe = torch.empty(out_host.shape, dtype=out_host.dtype, device="cuda")
        barrier()
        t_copy = time.perf_counter()
        for _ in range(3):
            out_host.copy_(probe, non_blocking=True)
        torch.cuda.synchronize()
        d2h_ms = (time.perf_counter() - t_copy) * 1e3 / 3
        d2h = {"plain_copy_ms": d2h_ms, "plain_copy_gbs_per_gpu": out_host.numel() * w / d2h_ms / 1e6}
        del probe, out_host
        session.close()
        del session, beliefs

        # ---- the same end to end with the device output stage: normalised single-variable
        # posteriors of the unobserved variables + log P(evidence) instead of raw factor beliefs ----
        free_vars = [v for v in sorted(net["sizes"]) if v not in evars]
        m_session = tree.marginals_session(net_values, B_e2e, free_vars, evars, dtype=dtype, dl=law, chunk=args.chunk)
        for _ in range(2):
            m_session.run(ev_np, copy=False)
        barrier()
        t_marg = time.perf_counter()
        for _ in range(e2e_steps):
            m_session.run(ev_np, copy=False)
        marg_ms = (time.perf_counter() - t_marg) * 1e3
        marg_d2h = int((m_session.engine.plan.fout_entries + 1) * B_e2e * w)
        m_session.close()
        del m_session

    # ---- N > 1: the one collective of the path, the all-gather of requested marginals ----
    gather = None
    if world > 1 and evars and not args.skip_e2e:
        gather = gather_leg(jt, jdist, tree, net, net_values, evars, ev_host, B_e2e, dtype, rank, world, barrier)

    # ---- the other BASELINE configs at their stated sizes (same step, same timing) ----
    extras = []
    if args.configs == "all" and args.semiring == "sum_product":
        if rank == 0:
            extras.append(single_propagate_latency(jt, wl.sprinkler()))
        for name, batch, dt, uni, evid, bel, sub in EXTRA_CONFIGS:
            if name == "ising16" and sub == 1:
                free, _ = torch.cuda.mem_get_info()
                sub = 1 if free > 150e9 else 2
            if name == "dag500" and bel:
                batch = bl.largest_batch(name, dt, batch)
            ehp = bl.HotPath(name, batch // sub, dt, uniform=uni, evidence=evid, beliefs=bel, dense=not args.no_dense,
                             ev_offset=rank * (batch // sub), ev_total=(batch // sub) * world)
            t = ehp.time(max(3, min(args.steps, 5)), 3, barrier)
            extras.append((ehp, t, sub))
            entry = config_entry(ehp, t, peak, sub)
            ehp.release()
            extras[-1] = entry
    times = [t_main["ms_per_step"], t_main["init_ms"], t_main["msg_ms"], e2e_ms, marg_ms] + \
        [e["ms_per_step"] for e in extras if "ms_per_step" in e]
    times = max_over_ranks(times)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    ms_per_step, init_ms, msg_ms, e2e_ms, marg_ms = times[:5]
    k = 5
    for e in extras:
        if "ms_per_step" in e:                       # max over ranks; every rank ran the same shape
            scale = times[k] / e["ms_per_step"]
            e["ms_per_step"] = times[k]
            e["value"] = e["value"] / scale * world
            e["frac"] /= scale
            e["frac_algorithmic"] /= scale
            k += 1

    value = B * world / (ms_per_step / 1e3)
    # the main line's own entry of the `configs` block: max-over-ranks time, whole-job value like the others
    scale = ms_per_step / main_entry["ms_per_step"]
    main_entry.update(ms_per_step=ms_per_step, value=value, frac=main_entry["frac"] / scale,
                      frac_algorithmic=main_entry["frac_algorithmic"] / scale)
    msg_ms_per_launch = msg_ms / max(msg_launches, 1)
    scheduled = S_msg * B / (msg_ms / 1e3) / 1e9            # GB/s this schedule moves in the message-passing launches
    algorithmic = A_msg * B / (msg_ms / 1e3) / 1e9
    step_gbs = S_all * B / (ms_per_step / 1e3) / 1e9

    cpu = None
    if not args.skip_cpu and world == 1:          # the CPU baseline is reported by the single-GPU run only
        v, cores, n, dt = cpu_throughput(args.config, args.cpu_per_core, semiring=args.semiring)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d instances of the same workload (%d per core, %.1f s) through oracle/ref_fixed.py "
                         "(NumPy restatement of the reference; a Python loop over instances, as the reference "
                         "itself is), %d processes" % (n, args.cpu_per_core, dt, cores)}

    traffic = load_traffic(args, uniform)
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
            "dense_contractions": bool(uniform and not args.no_dense),
            "step": "evidence slicing + clique init + collect + distribute (clique and separator beliefs)",
            "l2": "inputs larger than L2 (working set %.1f GB per GPU)" % working_set_gb,
            "step_gbs_scheduled": step_gbs, "init_ms_per_step": init_ms,
            "message_passing_ms_per_step": msg_ms,
        },
        "roofline": {
            "bound": "hbm", "kernel": "jt_project_tma_kernel<%s, %s> (+ jt_dense_kernel<double> for the dense "
                                      "contractions of uniform mode)" % (bl.SR_KERNEL[args.semiring],
                                                                         "double" if w == 8 else "float"),
            # bytes the schedule moves in the collect + distribute launches / their measured time
            "achieved": scheduled, "peak": peak, "unit": "GB/s", "frac": scheduled / peak,
            "achieved_algorithmic": algorithmic, "frac_algorithmic": algorithmic / peak,
            "peak_source": peak_src, "traffic": traffic["traffic_per_launch"] if traffic else None,
            "traffic_source": traffic,
            "launches_per_step": msg_launches, "avg_launch_ms": msg_ms_per_launch,
            "scheduled_bytes_per_launch": S_msg * B / max(msg_launches, 1),
            "algorithmic_bytes_per_launch": A_msg * B / max(msg_launches, 1),
            "note": ("frac = bytes this schedule has to move through HBM in the collect + distribute launches "
                     "(every buffer once per consuming task; uniform operands once per batch) / their CUDA-event "
                     "time / peak.  frac_algorithmic divides SURVEY.md 8d's A (every potential per instance in "
                     "HBM) by the same time: uniform mode does not stream the potentials no evidence reaches, so "
                     "it can exceed 1.  traffic is the ncu dram__bytes of the same command, recorded under "
                     "profiles/ (traffic_source), not measured in this run"),
        },
        "configs": [main_entry] + extras,
        "cpu_baseline": cpu,
        "e2e": {"value": B_e2e * world / (e2e_ms / e2e_steps / 1e3) if e2e_ms else None, "unit": UNIT,
                "h2d_bytes_per_step": int(ev_host.numel() * 4) if evars else 0,
                "d2h_bytes_per_step": int(plan.fout_entries * B_e2e * w),
                "ms_per_step": e2e_ms / e2e_steps, "chunk": e2e_chunk, "batch_per_gpu": B_e2e,
                "d2h_gbs_per_gpu": plan.fout_entries * B_e2e * w / (e2e_ms / e2e_steps) / 1e6 if e2e_ms else None,
                "host_ceiling": d2h,
                "what": "tree.propagate_session(values, batch, evidence_vars).run(evidence): host int32 "
                        "evidence -> device, propagate incl. marginalisation to factor scopes, per-factor "
                        "beliefs -> host memory; wall clock around the synchronous calls.  host_ceiling: the "
                        "same bytes as one plain device -> pinned-host copy on every rank at once"},
        "gpu_launches": int(launches),
        "gpu_launches_e2e": int(e2e_launches),
        "e2e_marginals": {"value": B_e2e * world / (marg_ms / e2e_steps / 1e3) if marg_ms else None, "unit": UNIT,
                          "ms_per_step": marg_ms / e2e_steps,
                          "h2d_bytes_per_step": int(ev_host.numel() * 4) if evars else 0,
                          "d2h_bytes_per_step": marg_d2h,
                          "what": "tree.marginals_session(...).run(evidence): same propagation, device output "
                                  "stage (normalised single-variable posteriors + log Z) -> host memory"},
        "all_gather": gather,
        "clocks": clocks,
    }
    print_line(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def gather_leg(jt, jdist, tree, net, values, evars, ev_host, B, dtype, rank, world, barrier):
    """The path's only collective: NCCL all-gather of the requested marginals.
    Every rank propagates its shard, normalises the single-variable posteriors on the device and
    gathers [B/G, M] -> [B, M]; rank 0 checks a sample of gathered rows (one of every rank's shard)
    against the CPU oracle."""
    import torch
    import torch.distributed as dist
    from junctiontree import engine as eng, _native
    free_vars = [v for v in sorted(net["sizes"]) if v not in evars]
    full = dict(net["sizes"])
    eff = dict(full)
    for v in evars:
        eff[v] = 1
    engine = tree._engine(eff, evars, full, outputs=[[v] for v in free_vars])
    plan = engine.plan
    fdev, _ = engine.factors_to_device(values, dtype)
    ev_dev = ev_host.to("cuda")
    ws = engine.new_pipeline_workspace(B, dtype)
    fout = torch.empty((plan.fout_entries, B), dtype=eng.torch_dtype(dtype), device="cuda")
    logz = torch.empty(B, dtype=eng.torch_dtype(dtype), device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    flags = _native.JT_NO_BELIEFS
    engine.dev.propagate(fdev.data_ptr(), False, ev_dev.data_ptr(), B, dtype, ws.data_ptr(), fout.data_ptr(), flags, stream)
    engine.dev.normalize(B, dtype, fout.data_ptr(), logz.data_ptr(), stream)
    local = fout.t().contiguous()                                   # [B, M] rows per instance
    for _ in range(2):
        jdist.all_gather_rows(local, B * world)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    reps = 3
    for _ in range(reps):
        gathered = jdist.all_gather_rows(local, B * world)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    tns = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(tns, op=dist.ReduceOp.MAX)
    ms = float(tns.item())
    checked, max_err = 0, 0.0
    if rank == 0:
        from oracle import ref_fixed                                # the checker, never the thing measured
        ct = tree.clique_tree
        ev_all = wl.draw_evidence(net, B * world)
        pick = [r * B + j for r in range(world) for j in (0, B - 1)]
        outs, _ = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                            net["factors"], net["sizes"], values, evars, ev_all[pick], n=len(pick))
        got = gathered[pick].cpu().numpy()
        col = 0
        for v, size in zip(free_vars, plan.fout_size):
            f = next(i for i, fv in enumerate(net["factors"]) if v in fv)
            axes = tuple(1 + a for a, u in enumerate(net["factors"][f]) if u != v)
            want = outs[f].sum(axis=axes)
            want = want / want.sum(axis=1, keepdims=True)
            max_err = max(max_err, float(np.max(np.abs(got[:, col:col + size] - want) / want)))
            col += size
        checked = len(pick)
        assert max_err < 1e-10, "all-gathered marginals differ from the oracle: %g" % max_err
    M = plan.fout_entries
    return {"collective": "nccl all_gather_into_tensor of normalised single-variable posteriors",
            "rows_per_rank": B, "row_entries": M, "bytes_received_per_rank": int((world - 1) * B * M * dtype.itemsize),
            "ms": ms, "gbs_per_rank": (world - 1) * B * M * dtype.itemsize / ms / 1e6,
            "rows_checked_against_oracle": checked, "max_rel_err": max_err}


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
    ap.add_argument("--no-dense", action="store_true",
                    help="uniform mode: keep the dense contractions on the projection kernels (A-B timing)")
    ap.add_argument("--configs", default="all", choices=["all", "none"],
                    help="all: also time the other BASELINE configs at their stated sizes (the `configs` block)")
    ap.add_argument("--skip-e2e", action="store_true")
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
