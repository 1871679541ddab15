"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.

NumPy interpreter for a compiled schedule (``junctiontree.schedule.Plan``).  It executes the
projection tasks exactly as documented in ``schedule.py`` / ``include/jt_b200.h`` on
``[entry, B]`` arrays, so the index tables, task wiring and launch order can be checked against
the oracles on a machine without a GPU.  The CUDA kernels implement the same task semantics.
"""

import numpy as np

from junctiontree import schedule as sch


def _map(tables, hi, lo, n, n_lo):
    x = np.arange(n, dtype=np.int64)
    return tables[hi + x // n_lo].astype(np.int64) + tables[lo + x % n_lo].astype(np.int64)


def evidence_offsets(plan, evidence, B):
    """fbase[f, b] = sum over observed axes of factor f of state * stride (stage V1)."""
    F = len(plan.factors)
    fbase = np.zeros((F, B), np.int64)
    for f in range(F):
        for k in range(plan.evf_ptr[f], plan.evf_ptr[f + 1]):
            fbase[f] += np.asarray(evidence)[:, plan.evf_var[k]].astype(np.int64) * plan.evf_stride[k]
    return fbase


def run(plan, B, work=None, factor_in=None, evidence=None, phases=None, dtype=np.float64):
    """Execute the plan.  ``work``: [work_entries, B] (clique potentials preloaded when the init
    phase is skipped); ``factor_in``: flat shared factor tables (fin_entries) or per-instance
    [fin_entries, B].  Returns (work, factor_out)."""
    tab = plan.tables
    if work is None:
        work = np.zeros((plan.work_entries, B), dtype)
    fout = np.zeros((plan.fout_entries, B), dtype)
    fbase = None
    if plan.factors is not None and plan.evidence_vars:
        fbase = evidence_offsets(plan, evidence, B)
    for phase, begin, end, _level in plan.launches_arr:
        if phases is not None and phase not in phases:
            continue
        # tasks of one launch are independent: evaluate all against the pre-launch state for
        # reads of other nodes, but in-place beta writes only touch the task's own clique
        for t in plan.tasks_arr[begin:end]:
            n_s, n_r, n_slo, n_rlo = (int(t[sch.T_NS]), int(t[sch.T_NR]),
                                      int(t[sch.T_NSLO]), int(t[sch.T_NRLO]))
            if t[sch.T_KIND] == sch.KIND_INIT:
                val = np.ones((n_s, B), dtype)
                for m in plan.msgs_arr[t[sch.T_SMSG_BEGIN]:t[sch.T_SMSG_END]]:
                    a = _map(tab, m[sch.M_AHI], m[sch.M_ALO], n_s, n_slo)
                    f = int(m[sch.M_FID])
                    if factor_in.ndim == 2:
                        idx = m[sch.M_OFF] + a
                        val = val * factor_in[idx, :]
                    else:
                        idx = m[sch.M_OFF] + a[:, None]
                        if fbase is not None:
                            idx = idx + fbase[f][None, :]
                        val = val * factor_in[idx]
                work[t[sch.T_OUT]:t[sch.T_OUT] + n_s] = val
                continue
            S = _map(tab, t[sch.T_SRC_SHI], t[sch.T_SRC_SLO], n_s, n_slo)
            R = _map(tab, t[sch.T_SRC_RHI], t[sch.T_SRC_RLO], n_r, n_rlo)
            e = t[sch.T_SRC] + S[:, None] + R[None, :]                       # [n_s, n_r]
            term = work[e]                                                   # [n_s, n_r, B]
            for m in plan.msgs_arr[t[sch.T_RMSG_BEGIN]:t[sch.T_RMSG_END]]:
                a = _map(tab, m[sch.M_AHI], m[sch.M_ALO], n_s, n_slo)
                b = _map(tab, m[sch.M_BHI], m[sch.M_BLO], n_r, n_rlo)
                term = term * work[m[sch.M_OFF] + a[:, None] + b[None, :]]
            sm = np.ones((n_s, B), dtype)
            for m in plan.msgs_arr[t[sch.T_SMSG_BEGIN]:t[sch.T_SMSG_END]]:
                a = _map(tab, m[sch.M_AHI], m[sch.M_ALO], n_s, n_slo)
                sm = sm * work[m[sch.M_OFF] + a]
            out = term.sum(axis=1) * sm
            own = work[t[sch.T_OWN]:t[sch.T_OWN] + n_s].copy() if t[sch.T_OWN] >= 0 else None
            if t[sch.T_OUT] >= 0:
                if t[sch.T_OUT_SPACE] == sch.SPACE_FOUT:
                    fout[t[sch.T_OUT]:t[sch.T_OUT] + n_s] = out
                else:
                    work[t[sch.T_OUT]:t[sch.T_OUT] + n_s] = out
            if t[sch.T_BEL] >= 0:
                work[t[sch.T_BEL]:t[sch.T_BEL] + n_s] = out * own
            if t[sch.T_BETA] >= 0:
                beta = term * sm[:, None, :]
                if own is not None:
                    beta = beta * own[:, None, :]
                work[t[sch.T_BETA] + S[:, None] + R[None, :]] = beta
    return work, fout


def node_array(plan, work, node, B):
    """Node ``node`` of the workspace as ``[B, *shape]``."""
    off, n = plan.node_off[node], plan.node_size[node]
    return np.moveaxis(work[off:off + n].reshape(tuple(plan.node_shape[node]) + (B,)), -1, 0)


def factor_array(plan, fout, f, B):
    off, n = plan.fout_off[f], plan.fout_size[f]
    return np.moveaxis(fout[off:off + n].reshape(tuple(plan.fout_shape[f]) + (B,)), -1, 0)


def flatten_factors(plan, values, dtype=np.float64):
    """Concatenate the factor tables in plan order (shared across the batch)."""
    flat = np.zeros(plan.fin_entries, dtype)
    for f, v in enumerate(values):
        v = np.asarray(v, dtype)
        assert list(v.shape) == plan.fin_shape[f], (v.shape, plan.fin_shape[f])
        flat[plan.fin_off[f]:plan.fin_off[f] + plan.fin_size[f]] = v.reshape(-1)
    return flat
