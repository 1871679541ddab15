"""Schedule emitter and the C ABI on a machine without a GPU: the library loads, exports every
symbol the header declares, parses and validates plan blobs; no compute call is made."""

import os
import re

import numpy as np
import pytest

import jt_workloads as wl
from helpers import compile_net
from junctiontree import _native
from junctiontree import schedule as sch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jt_b200.h")


def _plan(net, with_evidence=True):
    tree, seps, mc, f2c, eff, evars = compile_net(net, with_evidence)
    return sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"])


def test_library_exports_every_declared_symbol():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(jt_[a-z_0-9]+)\s*\(", text))
    assert len(declared) >= 18
    lib = _native.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "libjt_b200.so does not export %s" % name
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert lib.jt_abi_version() == _native.ABI_VERSION
    assert ("#define JT_ABI_VERSION %d" % _native.ABI_VERSION) in open(HEADER).read()


def test_header_constants_match_the_emitter():
    text = open(HEADER).read()

    def enum_after(marker):
        for body in re.findall(r"enum\s*\{(.*?)\}", text, flags=re.S):
            names = [n.strip().split("=")[0].strip() for n in body.split(",") if n.strip()]
            if marker in names:
                return names
        raise AssertionError("no enum with %s" % marker)

    hdr = enum_after("JT_H_MAGIC")
    assert hdr.index("JT_H_WORDS") == sch.H_WORDS
    assert hdr.index("JT_H_UNI_ENTRIES") == sch.H_UNI_ENTRIES
    task = enum_after("JT_T_KIND")
    assert task.index("JT_T_FLAGS") == sch.T_FLAGS and task.index("JT_T_NRLO") == sch.T_NRLO
    msg = enum_after("JT_M_OFF")
    assert msg.index("JT_M_UNI") == sch.M_UNI
    phases = enum_after("JT_PHASE_INIT")
    assert phases.index("JT_PHASE_COLLECT_INSTANCE") == sch.PHASE_COLLECT_INSTANCE
    assert "#define JT_TASK_WORDS %d" % sch.TASK_WORDS in text
    assert "#define JT_MSG_WORDS %d" % sch.MSG_WORDS in text
    assert "0x%X" % sch.MAGIC in text.upper().replace("0X", "0x")


@pytest.mark.parametrize("net", [wl.sprinkler(), wl.dag37(), wl.ising(4)], ids=lambda n: n["name"])
def test_plan_blob_round_trip_through_the_c_parser(net):
    plan = _plan(net)
    dp = _native.DevicePlan(plan.to_blob())
    assert dp.query(sch.H_NCLIQUES) == plan.n_cliques
    assert dp.query(sch.H_NSEPS) == plan.n_seps
    assert dp.query(sch.H_CLIQUE_ENTRIES) == plan.clique_entries
    assert dp.query(sch.H_SEP_ENTRIES) == plan.sep_entries
    assert dp.query(sch.H_NTASKS) == len(plan.tasks_arr)
    assert dp.query(sch.H_UNI_ENTRIES) == plan.uni_entries
    for k in range(len(plan.node_vars)):
        assert dp.node_range(k) == (plan.node_off[k], plan.node_size[k])
    for k in range(plan.n_cliques, len(plan.node_vars)):
        assert dp.message_offsets(k) == (plan.up_off(k), plan.down_off(k))
    B = 96
    lay = dp.workspace_layout(B, np.float64)
    assert lay["fbase"] >= plan.work_entries * B * 8
    assert lay["total"] == dp.workspace_bytes(B, np.float64)
    assert lay["total"] >= lay["uniform"] + plan.work_entries * 8
    assert dp.workspace_bytes(B, np.float32) < dp.workspace_bytes(B, np.float64)
    dp.close()


def test_malformed_blobs_are_rejected():
    plan = _plan(wl.huang_darwiche())
    blob = plan.to_blob()
    words = np.frombuffer(blob, np.int64).copy()

    def rejected(w, nbytes=None):
        data = w.tobytes()
        if nbytes is not None:
            data = data[:nbytes]
        with pytest.raises(_native.NativeError):
            _native.DevicePlan(data)

    bad = words.copy(); bad[sch.H_MAGIC] += 1; rejected(bad)
    bad = words.copy(); bad[sch.H_VERSION] = 1; rejected(bad)
    bad = words.copy(); bad[sch.H_NTASKS] += 1; rejected(bad)
    rejected(words, len(blob) - 8)
    rejected(words, 64)
    # a task whose table range leaves the table arena
    n_nodes = plan.n_cliques + plan.n_seps
    F = len(plan.factors)
    task0 = sch.H_WORDS + 2 * n_nodes + 4 * F + len(plan.ev_card) + (F + 1) + 2 * len(plan.evf_var)
    proj = next(i for i, t in enumerate(plan.tasks_arr) if t[sch.T_KIND] == sch.KIND_PROJECT)
    bad = words.copy(); bad[task0 + proj * sch.TASK_WORDS + sch.T_SRC_SLO] = plan.tables.size + 5; rejected(bad)
    bad = words.copy(); bad[task0 + proj * sch.TASK_WORDS + sch.T_NS] = 0; rejected(bad)
    bad = words.copy(); bad[task0 + proj * sch.TASK_WORDS + sch.T_SRC] = plan.work_entries + 1; rejected(bad)
    _native.DevicePlan(words.tobytes()).close()     # the untouched blob is fine


def test_compute_entry_points_fail_without_an_uploaded_plan():
    plan = _plan(wl.sprinkler())
    dp = _native.DevicePlan(plan.to_blob())
    with pytest.raises(_native.NativeError):
        dp.collect(1, np.float64, 1 << 20, 0, None)      # not uploaded: refused before any CUDA call
    with pytest.raises(TypeError):
        dp.workspace_bytes(4, np.int32)


@pytest.mark.parametrize("net", [wl.huang_darwiche(), wl.dag37(), wl.random_dag(30, 3, 2, 3, 30, 4)],
                         ids=lambda n: n["name"])
def test_schedule_structure(net):
    """One collect task per non-root clique, one distribute task per child separator, exactly
    one belief writer per clique, launches respect the data dependencies."""
    plan = _plan(net)
    T = plan.tasks_arr
    phase_of = {}
    for ph, b, e, lvl in plan.launches_arr:
        for t in range(b, e):
            phase_of.setdefault(t, []).append(int(ph))
    collect = [t for t, p in phase_of.items() if sch.PHASE_COLLECT in p]
    assert len(collect) == plan.n_cliques - 1
    assert sorted(T[t][sch.T_OUT] for t in collect) == sorted(plan.up_off(s) for s in range(plan.n_cliques, plan.n_cliques + plan.n_seps))
    dist = [t for t, p in phase_of.items() if sch.PHASE_DIST_PRE in p or sch.PHASE_DIST_MAIN in p]
    writers = [t for t in dist if T[t][sch.T_BETA] >= 0]
    assert sorted(T[t][sch.T_BETA] for t in writers) == sorted(plan.node_off[c] for c in range(plan.n_cliques))
    assert sorted(T[t][sch.T_OUT] for t in dist if T[t][sch.T_OUT] >= 0) == \
        sorted(plan.down_off(s) for s in range(plan.n_cliques, plan.n_cliques + plan.n_seps))
    # dependency order: a buffer is read only after the launch that writes it (general mode)
    written = set()
    general = [L for L in plan.launches_arr if L[0] in (sch.PHASE_INIT, sch.PHASE_COLLECT, sch.PHASE_DIST_PRE,
                                                         sch.PHASE_DIST_MAIN, sch.PHASE_MARGINAL)]
    for ph, b, e, lvl in general:
        outs = set()
        for t in T[b:e]:
            if t[sch.T_KIND] == sch.KIND_PROJECT:
                assert t[sch.T_SRC] in written, "clique potential read before it is written"
                for m in plan.msgs_arr[t[sch.T_RMSG_BEGIN]:t[sch.T_SMSG_END]]:
                    assert m[sch.M_OFF] in written, "message read before it is written"
                if t[sch.T_OWN] >= 0:
                    assert t[sch.T_OWN] in written
            if t[sch.T_OUT] >= 0 and t[sch.T_OUT_SPACE] == sch.SPACE_WORK:
                outs.add(int(t[sch.T_OUT]))
        written |= outs
    # the split launches cover the general ones exactly
    for all_ph, uni_ph, ins_ph in ((sch.PHASE_INIT, sch.PHASE_INIT_UNIFORM, sch.PHASE_INIT_INSTANCE),
                                   (sch.PHASE_COLLECT, sch.PHASE_COLLECT_UNIFORM, sch.PHASE_COLLECT_INSTANCE)):
        cover = lambda ph: sorted(t for L in plan.launches_arr if L[0] == ph for t in range(L[1], L[2]))
        assert cover(all_ph) == sorted(cover(uni_ph) + cover(ins_ph))
        for t in cover(uni_ph):
            assert T[t][sch.T_FLAGS] & sch.TF_TASK_UNIFORM


@pytest.mark.parametrize("net", [wl.dag37(), wl.random_dag(30, 3, 2, 3, 30, 4), wl.ising(5)],
                         ids=lambda n: n["name"])
def test_uniform_down_messages(net):
    """Uniform mode's distribute pass: a down-message is uniform iff its whole source side is;
    it is then computed by a PHASE_DIST_UNIFORM task from uniform operands only, every consumer
    reads the uniform copy, and the launches uniform mode runs (DIST_PRE_INSTANCE + DIST_MAIN)
    still write every down-message, separator belief and clique belief exactly once."""
    plan = _plan(net)
    T, M = plan.tasks_arr, plan.msgs_arr
    seps = range(plan.n_cliques, plan.n_cliques + plan.n_seps)
    child_of = {s: k for c in plan.order for s, k in plan.children[c]}
    for c in plan.order:
        above = plan.parent[c] < 0 or plan.uniform_down[plan.parent_sep[c]]
        for s, _ in plan.children[c]:
            others = all(plan.uniform_up[k2] for s2, k2 in plan.children[c] if s2 != s)
            assert plan.uniform_down[s] == (plan.uniform[c] and above and others)
    if net["name"] in ("dag37", "ising5x5"):
        assert any(plan.uniform_down.values()) and not all(plan.uniform_down.values())
    down_of = {plan.down_off(s): s for s in seps}
    tasks_of = lambda ph: [t for L in plan.launches_arr if L[0] == ph for t in range(L[1], L[2])]
    uni_tasks = tasks_of(sch.PHASE_DIST_UNIFORM)
    assert sorted(down_of[int(T[t][sch.T_OUT])] for t in uni_tasks) == sorted(s for s in seps if plan.uniform_down[s])
    for t in uni_tasks:
        assert T[t][sch.T_FLAGS] & sch.TF_TASK_UNIFORM and T[t][sch.T_FLAGS] & sch.TF_SRC_UNIFORM
        assert T[t][sch.T_BETA] < 0 and T[t][sch.T_BEL] < 0
        assert all(m[sch.M_UNI] for m in M[T[t][sch.T_RMSG_BEGIN]:T[t][sch.T_SMSG_END]])
    # every reader of a down-message is flagged according to its uniformity
    for t in range(len(T)):
        if T[t][sch.T_KIND] != sch.KIND_PROJECT:
            continue
        for m in M[T[t][sch.T_RMSG_BEGIN]:T[t][sch.T_SMSG_END]]:
            if int(m[sch.M_OFF]) in down_of:
                assert bool(m[sch.M_UNI]) == plan.uniform_down[down_of[int(m[sch.M_OFF])]]
    # what uniform mode launches per level writes each buffer once
    inst = tasks_of(sch.PHASE_DIST_PRE_INSTANCE) + tasks_of(sch.PHASE_DIST_MAIN)
    assert sorted(int(T[t][sch.T_OUT]) for t in inst if T[t][sch.T_OUT] >= 0) == sorted(down_of)
    assert sorted(int(T[t][sch.T_BEL]) for t in inst if T[t][sch.T_BEL] >= 0) == sorted(plan.bel_off(s) for s in seps)
    assert sorted(int(T[t][sch.T_BETA]) for t in inst if T[t][sch.T_BETA] >= 0) == \
        sorted(plan.node_off[c] for c in range(plan.n_cliques))
    # the elementwise forms: uniform source = the uniform down-message, own = the child's up-message
    general = set(tasks_of(sch.PHASE_DIST_PRE))
    for t in tasks_of(sch.PHASE_DIST_PRE_INSTANCE):
        if t in general:
            continue
        s = down_of[int(T[t][sch.T_OUT])]
        assert plan.uniform_down[s] and T[t][sch.T_NR] == 1 and T[t][sch.T_SRC] == T[t][sch.T_OUT]
        assert T[t][sch.T_FLAGS] & sch.TF_SRC_UNIFORM and T[t][sch.T_OWN] == plan.up_off(s)
        assert bool(T[t][sch.T_FLAGS] & sch.TF_OWN_UNIFORM) == plan.uniform_up[child_of[s]]


def test_uniform_flags_follow_the_evidence():
    net = wl.dag37()
    plan = _plan(net)
    observed = set(net["evidence_vars"])
    for c in range(plan.n_cliques):
        touched = any(observed & set(net["factors"][f]) for f in plan.clique_factors[c])
        assert plan.uniform[c] == (not touched)
    for c in plan.order:
        below = [c] + [k for k in plan.order if _is_below(plan, k, c)]
        assert plan.uniform_up[c] == all(plan.uniform[k] for k in below)
    assert plan.scheduled_entries(uniform=True) < plan.scheduled_entries(uniform=False)
    assert plan.scheduled_entries(uniform=False) >= plan.algorithmic_entries()


def _is_below(plan, k, c):
    while plan.parent[k] >= 0:
        k = plan.parent[k]
        if k == c:
            return True
    return False


def test_algorithmic_bytes_of_the_readme_network():
    """A = 8 (4*16 - 8 + 6*4) = 640 bytes for config 1."""
    plan = _plan(wl.sprinkler())
    assert 8 * plan.algorithmic_entries() == 640


def test_identical_tables_are_stored_once():
    plan = _plan(wl.ising(6))
    # the row sweep repeats the same clique structure: far fewer table entries than task maps
    refs = sum(int(t[sch.T_NS]) // int(t[sch.T_NSLO]) + int(t[sch.T_NSLO]) for t in plan.tasks_arr)
    assert plan.tables.size < refs / 2


@pytest.mark.parametrize("beliefs", [False, True], ids=["no_beliefs", "beliefs"])
@pytest.mark.parametrize("uniform", [True, False], ids=["uniform", "per_instance"])
@pytest.mark.parametrize("net", [wl.dag37(), wl.random_dag(16, 3, 2, 4, 6, 11), wl.ising(4),
                                 wl.large_state_tree((4, 6, 8, 4, 6, 8))], ids=lambda n: n["name"])
def test_sparse_workspace_rows_cover_exactly_what_the_schedule_touches(net, uniform, beliefs):
    """jt_workspace_sparse_rows (which rows of the workspace get memory) against the NumPy
    interpreter of the same plan: with every other row poisoned the outputs are unchanged (no
    row outside the intervals is read), no such row is written, and in uniform mode without
    clique beliefs a large part of the workspace needs no memory."""
    from oracle import plan_interp
    tree, seps, mc, f2c, eff, evars = compile_net(net)
    soft = [v for v in sorted(net["sizes"]) if v not in evars][:1]
    plan = sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"], likelihood_vars=soft)
    flags = (_native.JT_UNIFORM if uniform else 0) | (_native.JT_SEP_BELIEFS if beliefs else _native.JT_NO_BELIEFS)
    rows = _native.DevicePlan(plan.to_blob()).sparse_rows(flags)
    assert all(0 <= lo < hi <= plan.work_entries for lo, hi in rows)
    assert all(a[1] < b[0] for a, b in zip(rows, rows[1:]))          # merged and sorted
    mask = np.zeros(plan.work_entries, bool)
    for lo, hi in rows:
        mask[lo:hi] = True
    B = 2
    ev = wl.draw_evidence(net, B) if evars else None
    rng = np.random.default_rng(0)
    lik = {v: rng.random((B, net["sizes"][v])) + 0.1 for v in soft}
    fin = plan_interp.flatten_factors(plan, net["values"])

    def run(poison):
        work = np.full((plan.work_entries, B), np.nan if poison else 0.0)
        plan_interp.load_likelihoods(plan, work, lik)
        return plan_interp.run(plan, B, work=work, factor_in=fin, evidence=ev, uniform=uniform, beliefs=beliefs,
                               sep_beliefs=beliefs)

    work0, fout0 = run(False)
    work1, fout1 = run(True)
    assert np.all(np.isfinite(fout1)) and np.array_equal(fout0, fout1)
    assert np.all(np.isnan(work1[~mask]))                            # nothing outside the intervals was written
    written = ~np.isnan(work1).any(axis=1)
    assert written[mask].mean() > 0.9                                # and the intervals are (nearly) tight
    if uniform and not beliefs and net["name"] in ("dag37", "large_state_tree"):
        assert mask.mean() < 0.6


def test_host_entry_points_validate_their_arguments_without_a_device():
    """The host-buffer entry points and the sparse-workspace queries reject bad arguments before
    touching CUDA (status code + message, no exception, no crash)."""
    lib = _native.lib()
    plan = _plan(wl.huang_darwiche())
    dp = _native.DevicePlan(plan.to_blob())
    assert lib.jt_propagate_host(dp.handle, None, 0, None, 1, _native.JT_F64, None, None, None, None, None, 0, 0, None) \
        == 1 and b"null argument" in lib.jt_last_error_string()
    assert lib.jt_beliefs_host(None, None, _native.JT_F64, None, None, 0, None) == 1
    assert b"plan is null" in lib.jt_last_error_string()
    assert lib.jt_plan_single_launch(dp.handle, 1, 0) == 1                 # small tree, one instance
    assert lib.jt_plan_single_launch(dp.handle, 17, 0) == 0                # beyond 16 instances: per-level launches
    assert lib.jt_plan_single_launch(None, 1, 0) == 0
    big = _native.DevicePlan(_plan(wl.dag37()).to_blob())
    assert lib.jt_plan_single_launch(big.handle, 1, 0) == 0                # too many (s, r) items for one CTA
    mapped, dense = dp.sparse_bytes(64, np.float64, _native.JT_UNIFORM | _native.JT_NO_BELIEFS)
    assert 0 < mapped and dense > 0
    with pytest.raises(_native.NativeError):
        dp.sparse_bytes(0, np.float64, 0)
    # creating a sparse workspace needs a device: a clean error, not a crash
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(_native.NativeError):
            _native.SparseWorkspace(dp, 64, np.float64, 0)


def _dense_check(net, with_evidence=True):
    """Every dense contraction the library derives from a plan equals the projection task it
    replaces: random uniform workspace and message rows, W built as jt_dense_prep_kernel builds
    it, grouped products against the direct (s, r) sum."""
    plan = _plan(net, with_evidence)
    dp = _native.DevicePlan(plan.to_blob())
    tasks, dtab = dp.dense_tasks()
    rng = np.random.default_rng(7)
    tab = plan.tables.astype(np.int64)
    B = 3
    uni = rng.random(plan.work_entries) + 0.1
    work = rng.random((plan.work_entries, B)) + 0.1

    def amap(hi, lo, n, n_lo):
        x = np.arange(n)
        return tab[hi + x // n_lo] + tab[lo + x % n_lo]

    for d in tasks:
        t = plan.tasks_arr[d["task"]]
        n_s, n_r, n_slo, n_rlo = (int(t[k]) for k in (sch.T_NS, sch.T_NR, sch.T_NSLO, sch.T_NRLO))
        assert t[sch.T_FLAGS] & sch.TF_SRC_UNIFORM and t[sch.T_OUT] >= 0
        m = plan.msgs_arr[d["msg"]]
        assert not m[sch.M_UNI] and t[sch.T_RMSG_BEGIN] <= d["msg"] < t[sch.T_RMSG_END]
        S = amap(t[sch.T_SRC_SHI], t[sch.T_SRC_SLO], n_s, n_slo)
        R = amap(t[sch.T_SRC_RHI], t[sch.T_SRC_RLO], n_r, n_rlo)
        U = uni[t[sch.T_SRC] + S[:, None] + R[None, :]]                       # [n_s, n_r]
        for j in range(t[sch.T_RMSG_BEGIN], t[sch.T_RMSG_END]):
            mj = plan.msgs_arr[j]
            if mj[sch.M_UNI]:
                U = U * uni[mj[sch.M_OFF] + amap(mj[sch.M_AHI], mj[sch.M_ALO], n_s, n_slo)[:, None] +
                            amap(mj[sch.M_BHI], mj[sch.M_BLO], n_r, n_rlo)[None, :]]
        for j in range(t[sch.T_SMSG_BEGIN], t[sch.T_SMSG_END]):
            mj = plan.msgs_arr[j]
            if mj[sch.M_UNI]:
                U = U * uni[mj[sch.M_OFF] + amap(mj[sch.M_AHI], mj[sch.M_ALO], n_s, n_slo)][:, None]
        A = amap(m[sch.M_AHI], m[sch.M_ALO], n_s, n_slo)
        Bm = amap(m[sch.M_BHI], m[sch.M_BLO], n_r, n_rlo)
        direct = np.einsum("sr,srb->sb", U, work[m[sch.M_OFF] + A[:, None] + Bm[None, :]])
        n_g, n_i, K, n_q = d["n_g"], d["n_i"], d["K"], d["n_q"]
        assert n_g * n_i == n_s and K * n_q == n_r
        s_of = dtab[d["s_of"]:d["s_of"] + n_s].reshape(n_g, n_i)
        assert sorted(s_of.reshape(-1).tolist()) == list(range(n_s))
        r_of = dtab[d["r_of"]:d["r_of"] + n_r].reshape(K, n_q)
        assert sorted(r_of.reshape(-1).tolist()) == list(range(n_r))
        mg, mk = dtab[d["mg"]:d["mg"] + n_g], dtab[d["mk"]:d["mk"] + K]
        W = U[s_of[:, :, None, None], r_of[None, None, :, :]].sum(axis=3)     # [n_g, n_i, K]
        rows = work[m[sch.M_OFF] + mg[:, None] + mk[None, :]]                 # [n_g, K, B]
        got = np.empty_like(direct)
        got[s_of] = np.einsum("gik,gkb->gib", W, rows)
        np.testing.assert_allclose(got, direct, rtol=1e-13)
        # fragment layout bookkeeping
        assert d["MT"] == min(4, -(-n_i // 8)) and d["n_it"] == -(-n_i // (8 * d["MT"])) and d["n_k4"] == -(-K // 4)
        assert d["w_size"] == n_g * d["n_it"] * d["n_k4"] * d["MT"] * 32 and d["w_off"] % 32 == 0
        assert plan.launches_arr[d["launch"]][0] in (sch.PHASE_COLLECT_INSTANCE, sch.PHASE_DIST_PRE_INSTANCE,
                                                      sch.PHASE_DIST_MAIN_MESSAGES, sch.PHASE_MARGINAL_DIRECT)
    return len(tasks)


def test_dense_contractions_equal_the_projection_tasks_they_replace():
    assert _dense_check(wl.large_state_tree((8, 12, 16, 8, 12, 16))) >= 2
    assert _dense_check(wl.dag37()) >= 3
    assert _dense_check(wl.random_dag(60, 3, 2, 5, 8, 3)) >= 1
    assert _dense_check(wl.dag37(), with_evidence=False) == 0          # everything uniform: nothing per instance


def test_gpu_dense_nets_reach_every_variant_of_the_dense_kernel():
    """jt_dense_kernel is instantiated per m-tile count (1-4) with two stage layouts (several short
    units per stage / one unit over several stages) and zeroes its ring only when K is not a
    multiple of 4: the networks tests/test_gpu_dense.py runs on the GPU must reach all of them."""
    import junctiontree as jt
    from test_gpu_dense import NETS
    seen = set()
    for _, make in NETS:
        net = make()
        tree = jt.create_junction_tree(net["factors"], net["sizes"])
        tasks, _ = _native.DevicePlan(tree.plan(net["evidence_vars"]).to_blob()).dense_tasks()
        for d in tasks:
            seen.add(("MT", d["MT"]))
            seen.add(("several i-tiles", d["n_it"] > 1))
            seen.add(("K % 4 == 0", d["K"] % 4 == 0))
            seen.add(("short", d["K"] <= 8))
    for want in [("MT", 1), ("MT", 2), ("MT", 3), ("MT", 4), ("several i-tiles", True), ("several i-tiles", False),
                 ("K % 4 == 0", True), ("K % 4 == 0", False), ("short", True), ("short", False)]:
        assert want in seen, want


def test_plan_load_derivation_is_pinned():
    """What jt_plan_create derives from a blob (dense contractions, the table behind them incl. the
    belief kernel's walk tables) is read by the kernels as is.  Its fingerprints on the five
    BASELINE configs and the nets of the dense GPU tests are committed
    (tests/golden/derived_tables.json, make_derived.py): a plan-load change meant as a pure
    speed-up must not move them, and an intended change has to regenerate the file -- together
    with a GPU parity run."""
    import json
    knobs = [k for k in ("JT_DISABLE_DENSE", "JT_BETA_WALK", "JT_DENSE_MIN_GAIN", "JT_DENSE_WAVES", "JT_DENSE_BALANCE",
                         "JT_BETA_BLOCK", "JT_DISABLE_BETA", "JT_BETA_MIN_MB", "JT_BETA_MAX_B") if os.environ.get(k)]
    if knobs:
        pytest.skip("selection-rule knobs set in the environment: %s" % knobs)
    sys_path_entry = os.path.join(ROOT, "tests", "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_derived", os.path.join(sys_path_entry, "make_derived.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(mod.OUT) as fh:
        want = json.load(fh)
    got = mod.fingerprints()
    assert set(got) == set(want)
    for name in sorted(want):
        assert got[name] == want[name], "derived tables of %s changed (see tests/golden/make_derived.py)" % name
    assert mod.fingerprints() == got                    # built on all host cores: still deterministic
