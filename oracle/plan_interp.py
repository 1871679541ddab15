"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.

NumPy interpreter for a compiled schedule (``junctiontree.schedule.Plan``).  It executes the
projection tasks exactly as documented in ``schedule.py`` / ``include/jt_b200.h`` on
``[entry, B]`` arrays, so the index tables, task wiring and launch order can be checked against
the oracles on a machine without a GPU.  The CUDA kernels implement the same task semantics.
"""

import numpy as np

from junctiontree import schedule as sch

from .ref_fixed import semiring_ops


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


def run(plan, B, work=None, factor_in=None, evidence=None, phases=None, dtype=np.float64, uniform=False,
        beliefs=True, semiring="sum_product", sep_beliefs=True):
    """Execute the plan.  ``work``: [work_entries, B] (clique potentials preloaded when the init
    phase is skipped); ``factor_in``: flat shared factor tables (fin_entries) or per-instance
    [fin_entries, B].  ``uniform``: keep the potentials of evidence-free cliques once, in the
    uniform region, as the device does for shared factor tables.  ``semiring``: the (+, x) pair
    of the task (``ref_fixed.SEMIRINGS``).  ``sep_beliefs``: the JT_SEP_BELIEFS flag (separator
    beliefs up * down are stored).  Returns (work, factor_out)."""
    mul, reduce_, one = semiring_ops(semiring)
    tab = plan.tables
    uni = np.zeros((plan.work_entries, 1), dtype)       # the uniform workspace: same offsets, B = 1
    general = {sch.PHASE_INIT, sch.PHASE_COLLECT}
    split = {sch.PHASE_INIT_UNIFORM, sch.PHASE_INIT_INSTANCE, sch.PHASE_COLLECT_UNIFORM, sch.PHASE_COLLECT_INSTANCE}
    in_uni_ws = {sch.PHASE_INIT_UNIFORM, sch.PHASE_COLLECT_UNIFORM, sch.PHASE_DIST_UNIFORM}
    general |= {sch.PHASE_DIST_PRE}
    split |= {sch.PHASE_DIST_UNIFORM, sch.PHASE_DIST_PRE_INSTANCE}
    skip = set(general if uniform else split)
    # beliefs=False: messages and outputs only -- no clique belief is written, outputs come
    # straight from psi_C and the incoming messages
    skip |= ({sch.PHASE_DIST_MAIN, sch.PHASE_MARGINAL} if not beliefs else
             {sch.PHASE_DIST_MAIN_MESSAGES, sch.PHASE_MARGINAL_DIRECT})
    launches = list(plan.launches_arr)
    if uniform:   # evidence-free subtrees are collected first, once
        launches = [L for L in launches if L[0] in in_uni_ws] + [L for L in launches if L[0] not in in_uni_ws]
    real_work = work
    fout = np.zeros((plan.fout_entries, B), dtype)
    fbase = None
    if plan.factors is not None and plan.evidence_vars:
        fbase = evidence_offsets(plan, evidence, B)
    work = real_work if real_work is not None else np.zeros((plan.work_entries, B), dtype)
    real_work = work
    for phase, begin, end, _level in launches:
        if (phases is not None and phase not in phases) or phase in skip:
            continue
        # a launch in the uniform workspace is an ordinary B = 1 launch on that buffer
        work = uni if phase in in_uni_ws else real_work
        flagged = uniform and phase not in in_uni_ws
        # tasks of one launch are independent: evaluate all against the pre-launch state for
        # reads of other nodes, but in-place beta writes only touch the task's own clique
        for t in plan.tasks_arr[begin:end]:
            n_s, n_r, n_slo, n_rlo = (int(t[sch.T_NS]), int(t[sch.T_NR]),
                                      int(t[sch.T_NSLO]), int(t[sch.T_NRLO]))
            if t[sch.T_KIND] == sch.KIND_INIT:
                to_uni = phase in in_uni_ws
                val = np.full((n_s, 1 if to_uni else B), one, dtype)
                for m in plan.msgs_arr[t[sch.T_SMSG_BEGIN]:t[sch.T_SMSG_END]]:
                    a = _map(tab, m[sch.M_AHI], m[sch.M_ALO], n_s, n_slo)
                    f = int(m[sch.M_FID])
                    if f == -2:                       # soft evidence: likelihood table in the workspace
                        val = mul(val, work[m[sch.M_OFF] + a, :])
                        continue
                    if factor_in.ndim == 2:
                        idx = m[sch.M_OFF] + a
                        val = mul(val, factor_in[idx, :])
                    else:
                        idx = m[sch.M_OFF] + a[:, None]
                        if fbase is not None and not to_uni:
                            idx = idx + fbase[f][None, :]
                        val = mul(val, factor_in[idx])
                work[t[sch.T_OUT]:t[sch.T_OUT] + n_s] = val
                continue
            S = _map(tab, t[sch.T_SRC_SHI], t[sch.T_SRC_SLO], n_s, n_slo)
            R = _map(tab, t[sch.T_SRC_RHI], t[sch.T_SRC_RLO], n_r, n_rlo)
            e = t[sch.T_SRC] + S[:, None] + R[None, :]                       # [n_s, n_r]
            def buf(is_uniform):
                return uni if (flagged and is_uniform) else work
            term = buf(t[sch.T_FLAGS] & sch.TF_SRC_UNIFORM)[e]               # [n_s, n_r, B or 1]
            for m in plan.msgs_arr[t[sch.T_RMSG_BEGIN]:t[sch.T_RMSG_END]]:
                a = _map(tab, m[sch.M_AHI], m[sch.M_ALO], n_s, n_slo)
                b = _map(tab, m[sch.M_BHI], m[sch.M_BLO], n_r, n_rlo)
                term = mul(term, buf(m[sch.M_UNI])[m[sch.M_OFF] + a[:, None] + b[None, :]])
            sm = np.full((n_s, work.shape[1]), one, dtype)
            for m in plan.msgs_arr[t[sch.T_SMSG_BEGIN]:t[sch.T_SMSG_END]]:
                a = _map(tab, m[sch.M_AHI], m[sch.M_ALO], n_s, n_slo)
                sm = mul(sm, buf(m[sch.M_UNI])[m[sch.M_OFF] + a])
            out = mul(reduce_(term, 1), sm)
            own = buf(t[sch.T_FLAGS] & sch.TF_OWN_UNIFORM)[t[sch.T_OWN]:t[sch.T_OWN] + n_s].copy() \
                if t[sch.T_OWN] >= 0 else None
            if t[sch.T_OUT] >= 0:
                if t[sch.T_OUT_SPACE] == sch.SPACE_FOUT:
                    fout[t[sch.T_OUT]:t[sch.T_OUT] + n_s] = out
                else:
                    work[t[sch.T_OUT]:t[sch.T_OUT] + n_s] = out
            if t[sch.T_BEL] >= 0 and sep_beliefs:
                work[t[sch.T_BEL]:t[sch.T_BEL] + n_s] = mul(out, own)
            if t[sch.T_BETA] >= 0 and beliefs:
                beta = mul(term, sm[:, None, :])
                if own is not None:
                    beta = mul(beta, own[:, None, :])
                work[t[sch.T_BETA] + S[:, None] + R[None, :]] = beta
    return real_work, fout


def node_array(plan, work, node, B):
    """Node ``node`` of the workspace as ``[B, *shape]``."""
    off, n = plan.node_off[node], plan.node_size[node]
    return np.moveaxis(work[off:off + n].reshape(tuple(plan.node_shape[node]) + (B,)), -1, 0)


def factor_array(plan, fout, f, B):
    off, n = plan.fout_off[f], plan.fout_size[f]
    return np.moveaxis(fout[off:off + n].reshape(tuple(plan.fout_shape[f]) + (B,)), -1, 0)


def load_likelihoods(plan, work, likelihoods):
    """Write ``{variable: [B, size]}`` likelihood vectors into the workspace's likelihood region."""
    for k, v in enumerate(plan.likelihood_vars):
        off = plan.lik_base + plan.lik_off[k]
        work[off:off + plan.lik_size[k]] = np.asarray(likelihoods[v], work.dtype).T
    return work


def flatten_factors(plan, values, dtype=np.float64):
    """Concatenate the factor tables in plan order (shared across the batch)."""
    flat = np.zeros(plan.fin_entries, dtype)
    for f, v in enumerate(values):
        v = np.asarray(v, dtype)
        assert list(v.shape) == plan.fin_shape[f], (v.shape, plan.fin_shape[f])
        flat[plan.fin_off[f]:plan.fin_off[f] + plan.fin_size[f]] = v.reshape(-1)
    return flat
