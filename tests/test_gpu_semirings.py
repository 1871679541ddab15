"""GPU parity of the other distributive laws (max-product, log-sum-exp, max-sum) against the
semiring oracle (``oracle/ref_fixed.py``, pinned in ``tests/test_semirings_cpu.py``): same
kernels, same schedule, another (+, x) pair.  Tolerances as for sum-product: 1e-12 (float64) /
1e-5 (float32) relative on every potential; for log-domain laws that is an absolute error on
the log value, and -inf (probability zero) must match exactly."""

import numpy as np
import pytest

import jt_workloads as wl
from helpers import RTOL_F32, RTOL_F64, SEMIRING_NAMES, assert_close, assert_close_semiring, semiring_inputs

pytestmark = pytest.mark.gpu


def _law(name):
    from junctiontree import semirings as sr
    return {"max_product": sr.max_product, "log_sum_exp": sr.log_sum_exp, "max_sum": sr.max_sum}[name]


def _nets():
    return [wl.sprinkler(), wl.huang_darwiche(), wl.random_dag(12, 3, 2, 3, 8, 5),
            wl.random_dag(16, 3, 2, 4, 6, 11), wl.ising(4), wl.large_state_tree((4, 6, 8, 4, 6, 8))]


def _oracle(tree, net, vals, evars, ev, B, semiring):
    from oracle import ref_fixed
    ct = tree.clique_tree
    return ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                     net["factors"], net["sizes"], vals, evars, ev, n=B, semiring=semiring)


@pytest.mark.parametrize("semiring", SEMIRING_NAMES)
@pytest.mark.parametrize("uniform", [True, False], ids=["uniform", "per_instance"])
@pytest.mark.parametrize("net", _nets(), ids=lambda n: n["name"])
@pytest.mark.parametrize("B", [1, 7, 70, 300])
def test_batched_propagation_semiring_f64(net, B, uniform, semiring):
    """LDG kernel (small / odd batches) and TMA kernel (B >= 128), with and without uniform mode."""
    import junctiontree as jt
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = net.get("evidence_vars", [])
    ev = wl.draw_evidence(net, B) if evars else None
    vals = semiring_inputs(net["values"], semiring)
    outs, nodes = tree.propagate_batch(vals, evars, ev, batch=B, nodes=True, uniform=uniform, dl=_law(semiring))
    want_f, want_n = _oracle(tree, net, vals, evars, ev, B, semiring)
    for k, (g, w) in enumerate(zip(nodes, want_n)):
        assert_close_semiring(g, w, RTOL_F64, semiring, "node %d" % k)
    for f, (g, w) in enumerate(zip(outs, want_f)):
        assert_close_semiring(g, w, RTOL_F64, semiring, "factor %d" % f)


@pytest.mark.parametrize("semiring", SEMIRING_NAMES)
@pytest.mark.parametrize("B", [6, 520])
def test_batched_propagation_semiring_f32(B, semiring):
    import junctiontree as jt
    net = wl.random_dag(16, 3, 2, 4, 6, 11)
    vals32 = [np.asarray(v, np.float32) for v in semiring_inputs(net["values"], semiring)]
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    ev = wl.draw_evidence(net, B)
    outs, nodes = tree.propagate_batch(vals32, evars, ev, nodes=True, dl=_law(semiring))
    assert all(o.dtype == np.float32 for o in outs)
    want_f, want_n = _oracle(tree, net, [np.asarray(v, np.float64) for v in vals32], evars, ev, B, semiring)
    # log-domain float32: the error of a log value scales with its magnitude (|log p| up to ~40)
    for k, (g, w) in enumerate(zip(list(nodes) + list(outs), list(want_n) + list(want_f))):
        assert_close_semiring(g, w, RTOL_F32, semiring, "array %d" % k)


@pytest.mark.parametrize("semiring", SEMIRING_NAMES)
def test_outputs_without_clique_beliefs_and_single_propagate(semiring):
    """JT_NO_BELIEFS path (direct marginals) and the CUDA-graph single-instance ``propagate``."""
    import junctiontree as jt
    for net in (wl.sprinkler(), wl.huang_darwiche(), wl.random_dag(12, 3, 2, 3, 8, 5)):
        tree = jt.create_junction_tree(net["factors"], net["sizes"])
        vals = semiring_inputs(net["values"], semiring)
        want_f, _ = _oracle(tree, net, vals, [], None, 1, semiring)
        got = tree.propagate(vals, dl=_law(semiring))
        for f, (g, w) in enumerate(zip(got, want_f)):
            assert_close_semiring(g, w[0], RTOL_F64, semiring, "%s factor %d" % (net["name"], f))
        B = 130
        evars = net.get("evidence_vars", [])
        ev = wl.draw_evidence(net, B) if evars else None
        outs = tree.propagate_batch(vals, evars, ev, batch=B, dl=_law(semiring))
        want_f, _ = _oracle(tree, net, vals, evars, ev[:3] if ev is not None else None, 3, semiring)
        for f, w in enumerate(want_f):
            assert_close_semiring(outs[f][:3], w, RTOL_F64, semiring, "%s factor %d" % (net["name"], f))


@pytest.mark.parametrize("semiring", SEMIRING_NAMES)
def test_compute_beliefs_and_einsum_in_other_semirings(semiring):
    """``compute_beliefs(tree, potentials, clique_vars, dl)`` and ``dl.einsum`` on the device."""
    from junctiontree import computation as comp
    from oracle import brute, ref_fixed
    law = _law(semiring)
    rng = np.random.default_rng(3)
    node_vars = [["a", "b", "c"], ["c", "d"], ["b", "e", "f"], ["c"], ["b"]]
    sizes = dict(a=3, b=4, c=2, d=5, e=3, f=2)
    tree = [0, (3, [1]), (4, [2])]
    pots = [rng.random([sizes[v] for v in vs]) + 0.1 for vs in node_vars[:3]]
    one = ref_fixed.semiring_ops(semiring)[2]
    pots += [np.full([sizes[v] for v in vs], one) for vs in node_vars[3:]]
    if semiring in ("log_sum_exp", "max_sum"):
        pots[:3] = [np.log(p) for p in pots[:3]]
    got = comp.compute_beliefs(tree, pots, node_vars, law)
    want = brute.tree_beliefs(tree, node_vars, pots, semiring)
    for k, (g, w) in enumerate(zip(got, want)):
        assert_close_semiring(g, w, RTOL_F64, semiring, "node %d" % k)
    x, y = pots[0], pots[2]
    got = law.einsum(x, ["a", "b", "c"], y, ["b", "e", "f"], ["f", "a"])
    want = ref_fixed._einsum(x, ["a", "b", "c"], y, ["b", "e", "f"], ["f", "a"], semiring=semiring)
    assert_close_semiring(got, want, RTOL_F64, semiring, "einsum")
    assert_close_semiring(law.project(x, ["a", "b", "c"], ["c", "a"]),
                          ref_fixed._einsum(x, ["a", "b", "c"], ["c", "a"], semiring=semiring), RTOL_F64, semiring,
                          "project")


def test_split_r_kernel_in_other_semirings():
    """Few instances with long reductions (split-r kernel, tree reduction in shared memory)."""
    import junctiontree as jt
    rng = np.random.default_rng(4)
    big = ["v%02d" % i for i in range(13)]
    factors = [big, ["v00", "w0"], ["v05", "v06", "w1"], ["w1", "w2"], ["v12"]]
    sizes = {v: 2 for v in big}
    sizes.update(w0=3, w1=2, w2=4)
    values = [rng.random([sizes[v] for v in f]) + 0.05 for f in factors]
    net = {"factors": factors, "sizes": sizes, "values": values, "evidence_vars": ["w2"], "seed": 3}
    tree = jt.create_junction_tree(factors, sizes)
    for semiring in SEMIRING_NAMES:
        for B in (1, 5):
            ev = wl.draw_evidence(net, B)
            vals = semiring_inputs(values, semiring)
            outs, nodes = tree.propagate_batch(vals, ["w2"], ev, nodes=True, dl=_law(semiring))
            want_f, want_n = _oracle(tree, net, vals, ["w2"], ev, B, semiring)
            for k, (g, w) in enumerate(zip(list(nodes) + list(outs), list(want_n) + list(want_f))):
                assert_close_semiring(g, w, RTOL_F64, semiring, "%s B=%d array %d" % (semiring, B, k))


def test_map_and_log_partition_through_marginals_batch():
    """Output stage per semiring: max-marginals normalised by their maximum have their argmax at
    the MAP state and log_z = log of the best joint probability; log-sum-exp returns log
    posteriors and the same log Z as sum-product."""
    import junctiontree as jt
    from junctiontree import semirings as sr
    from oracle import brute
    net = wl.random_dag(11, 3, 2, 3, 8, 9)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars, B = net["evidence_vars"], 9
    ev = wl.draw_evidence(net, B)
    free = [v for v in sorted(net["sizes"]) if v not in evars]
    post, log_z = tree.marginals_batch(net["values"], free, evars, ev)
    logv = semiring_inputs(net["values"], "log_sum_exp")
    lpost, llog_z = tree.marginals_batch(logv, free, evars, ev, dl=sr.log_sum_exp)
    assert_close(llog_z, log_z, 1e-11, "log Z")
    for v in free:
        assert_close(np.exp(lpost[v]), post[v], 1e-11, "posterior %s" % v)
    mm, mlog = tree.marginals_batch(net["values"], free, evars, ev, dl=sr.max_product)
    ms, mslog = tree.marginals_batch(logv, free, evars, ev, dl=sr.max_sum)
    assert_close(mslog, mlog, 1e-11, "log of the best joint state")
    for b in range(B):
        evd = {v: int(ev[b][i]) for i, v in enumerate(evars)}
        joint = brute.factor_graph_marginals(net["factors"], net["values"], [free], evd)[0]
        best = np.unravel_index(np.argmax(joint), joint.shape)
        assert tuple(int(np.argmax(mm[v][b])) for v in free) == tuple(int(i) for i in best)
        assert tuple(int(np.argmax(ms[v][b])) for v in free) == tuple(int(i) for i in best)
        assert_close(mlog[b], np.log(joint.max()), 1e-11, "MAP value instance %d" % b)
        for v in free:
            assert np.isclose(mm[v][b].max(), 1.0, rtol=1e-12) and np.isclose(ms[v][b].max(), 0.0, atol=1e-12)


def test_log_domain_survives_underflow():
    """A chain whose sum-product underflows float64 (Z ~ 1e-400): log-sum-exp returns the exact
    log partition function and log marginals."""
    import junctiontree as jt
    from junctiontree import semirings as sr
    n = 40
    rng = np.random.default_rng(1)
    labels = ["x%02d" % i for i in range(n)]
    factors = [[labels[i], labels[i + 1]] for i in range(n - 1)]
    sizes = {v: 3 for v in labels}
    logv = [np.log(rng.random((3, 3)) + 0.1) - 25.0 for _ in factors]      # each entry ~ e^-25
    tree = jt.create_junction_tree(factors, sizes)
    lmarg, log_z = tree.marginals_batch(logv, [labels[0], labels[-1]], batch=2, dl=sr.log_sum_exp)
    # exact reference: the same chain with the -25 offsets removed, then added back in log space
    marg, log_z0 = tree.marginals_batch([np.exp(v + 25.0) for v in logv], [labels[0], labels[-1]], batch=2)
    assert_close(log_z, log_z0 - 25.0 * (n - 1), 1e-12, "log Z")
    assert log_z[0] < -900 and np.exp(log_z[0]) == 0.0                      # underflows as a plain product
    for v in (labels[0], labels[-1]):
        assert_close(np.exp(lmarg[v]), marg[v], 1e-11, "posterior %s" % v)


def test_streamed_pipeline_in_another_semiring():
    """Host-in / host-out batches above the streaming threshold (chunked two-stream pipeline)."""
    import junctiontree as jt
    from junctiontree import semirings as sr
    net = wl.random_dag(14, 3, 2, 3, 8, 2)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    B = 8192 + 300
    ev = wl.draw_evidence(net, B)
    outs = tree.propagate_batch(net["values"], net["evidence_vars"], ev, dl=sr.max_product)
    pick = [0, 8191, 8192, B - 1]
    want_f, _ = _oracle(tree, net, net["values"], net["evidence_vars"], ev[pick], len(pick), "max_product")
    for f, w in enumerate(want_f):
        assert_close(outs[f][pick], w, RTOL_F64, "factor %d" % f)


@pytest.mark.parametrize("semiring", SEMIRING_NAMES)
def test_semiring_properties_at_scale_config2(semiring):
    """Config 2 at 16,384 instances under the other laws, size-independent properties on every
    instance on the device: all cliques and separators agree on the semiring total (max-product:
    the value of the MAP state; log-sum-exp: log Z, equal to the sum-product log Z), a separator
    belief is the semiring marginal of its child clique; the first instances against the oracle."""
    import torch
    import junctiontree as jt
    net = wl.dag37()
    B = 16384
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    ev = wl.draw_evidence(net, B)
    vals = semiring_inputs(net["values"], semiring)
    outs, nodes = tree.propagate_batch(vals, evars, ev, nodes=True, device_output=True, dl=_law(semiring))
    plan = tree.plan(evars)

    def total(x, dims):
        if semiring == "log_sum_exp":
            return torch.logsumexp(x, dim=dims)
        return torch.amax(x, dim=dims)

    log_domain = semiring in ("log_sum_exp", "max_sum")
    Z = total(nodes[plan.root].reshape(B, -1), [1])
    assert bool(torch.isfinite(Z).all())
    for k, nd in enumerate(nodes):
        zk = total(nd.reshape(B, -1), [1])
        err = (zk - Z).abs() if log_domain else (zk - Z).abs() / Z
        assert float(err.max()) < 1e-11, "node %d" % k
    for c in plan.order[::3]:
        for s, k in plan.children[c]:
            kv, sv = plan.node_vars[k], plan.node_vars[s]
            axes = [1 + i for i, v in enumerate(kv) if v not in sv]
            marg = total(nodes[k], axes) if axes else nodes[k]
            kept = [v for v in kv if v in sv]
            marg = marg.permute([0] + [1 + kept.index(v) for v in sv])
            err = (marg - nodes[s]).abs() if log_domain else (marg - nodes[s]).abs() / nodes[s].abs().clamp_min(1e-300)
            assert float(err.max()) < 1e-10
    if semiring == "log_sum_exp":
        _, logz = tree.marginals_batch(net["values"], [sorted(net["sizes"])[0]], evars, ev[:512])
        assert_close(Z[:512].cpu().numpy(), logz, 1e-11, "log Z against the sum-product output stage")
    want_f, want_n = _oracle(tree, net, vals, evars, ev[:3], 3, semiring)
    for f, w in enumerate(want_f):
        assert_close_semiring(outs[f][:3].cpu().numpy(), w, RTOL_F64, semiring, "factor %d" % f)
    for k in range(0, len(nodes), 5):
        assert_close_semiring(nodes[k][:3].cpu().numpy(), want_n[k], RTOL_F64, semiring, "node %d" % k)
    del outs, nodes
    tree.clique_tree._engines.clear()
    torch.cuda.empty_cache()
