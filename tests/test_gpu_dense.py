"""Dense contractions (jt_dense.cu: shared potential x one per-instance message on the FP64 tensor
pipe) against the projection kernels they replace and against the oracle, through the C ABI."""

import numpy as np
import pytest

import jt_workloads as wl
from helpers import RTOL_F64, assert_close
from junctiontree import _native

pytestmark = pytest.mark.gpu


def _oracle(tree, net, evars, ev, n):
    from oracle import ref_fixed
    ct = tree.clique_tree
    return ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                     net["factors"], net["sizes"], net["values"], evars, ev, n=n)


NETS = [
    ("large_state_small", lambda: wl.large_state_tree((8, 12, 16, 8, 12, 16))),   # config 4's shape, 1/512 of the entries
    ("large_state_odd", lambda: wl.large_state_tree((5, 7, 9, 3, 11, 6))),        # nothing a multiple of 8 or 4
    ("dag37", wl.dag37),
    ("dag60", lambda: wl.random_dag(60, 3, 2, 5, 8, 3)),
    ("dag24_wide", lambda: wl.random_dag(24, 4, 3, 6, 6, 11)),
]


@pytest.mark.parametrize("B", [128, 130, 300, 1024])
@pytest.mark.parametrize("name,make", NETS, ids=[n for n, _ in NETS])
def test_dense_equals_projection_and_oracle_with_beliefs(name, make, B):
    import junctiontree as jt
    net = make()
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    ev = wl.draw_evidence(net, B)
    engine = tree._engine(tree.plan(evars).sizes, evars, tree.plan(evars).full_sizes)
    n_dense = len(engine.dev.dense_tasks()[0])
    assert n_dense > 0, "the plan has no dense contraction: the test would not exercise jt_dense_kernel"
    l0 = _native.launch_count()
    outs_d, nodes_d = tree.propagate_batch(net["values"], evars, ev, nodes=True, dense=True)
    l1 = _native.launch_count()
    outs_p, nodes_p = tree.propagate_batch(net["values"], evars, ev, nodes=True, dense=False)
    l2 = _native.launch_count()
    assert l1 - l0 > l2 - l1, "dense launches did not run"
    for k, (a, b) in enumerate(zip(list(outs_d) + list(nodes_d), list(outs_p) + list(nodes_p))):
        assert_close(a, b, 1e-13, "%s entry %d dense vs projection" % (name, k))
    n = min(B, 3)
    want_f, want_n = _oracle(tree, net, evars, ev[:n], n)
    for k, w in enumerate(want_n):
        assert_close(nodes_d[k][:n], w, RTOL_F64, "node %d" % k)
    for f, w in enumerate(want_f):
        assert_close(outs_d[f][:n], w, RTOL_F64, "factor %d" % f)


@pytest.mark.parametrize("B", [256, 1000])
@pytest.mark.parametrize("name,make", NETS, ids=[n for n, _ in NETS])
def test_dense_without_beliefs_outputs_only(name, make, B):
    """The pipelines' mode (JT_NO_BELIEFS): the message-sending tasks of the writers and the direct
    marginals are dense contractions too."""
    import junctiontree as jt
    net = make()
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    ev = wl.draw_evidence(net, B)
    outs_d = tree.propagate_batch(net["values"], evars, ev, dense=True)
    outs_p = tree.propagate_batch(net["values"], evars, ev, dense=False)
    for f, (a, b) in enumerate(zip(outs_d, outs_p)):
        assert_close(a, b, 1e-13, "%s factor %d dense vs projection" % (name, f))
    want_f, _ = _oracle(tree, net, evars, ev[:3], 3)
    for f, w in enumerate(want_f):
        assert_close(outs_d[f][:3], w, RTOL_F64, "factor %d" % f)


def test_dense_reuses_the_w_region_across_chunks_of_a_session():
    """JT_UNIFORM_VALID: later chunks of a pipeline skip the uniform phases and the W rebuild."""
    import junctiontree as jt
    net = wl.random_dag(60, 3, 2, 5, 8, 3)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    B = 4096
    ev = wl.draw_evidence(net, B)
    free = [v for v in sorted(net["sizes"]) if v not in evars]
    with tree.marginals_session(net["values"], B, free, evars, chunk=512) as session:
        for _ in range(2):
            marg, logz = session.run(ev)
    want, want_logz = tree.marginals_batch(net["values"], free, evars, ev[:300])
    for v in free:
        assert_close(marg[v][:300], want[v], 1e-12, "marginal %s" % v)
    assert_close(logz[:300], want_logz, 1e-12, "log Z")
    # and against brute force on a few instances
    from oracle import ref_fixed
    ct = tree.clique_tree
    outs, _ = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                        net["factors"], net["sizes"], net["values"], evars, ev[:2], n=2)
    z = outs[0].reshape(2, -1).sum(axis=1)
    assert_close(np.exp(logz[:2]), z, 1e-12, "Z")


def test_config4_dense_full_size():
    """Config 4 at its stated size: [128 x 96] and [96 x 64] contractions per separator slice."""
    import torch
    import junctiontree as jt
    net = wl.large_state_tree()
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    B = 512
    ev = wl.draw_evidence(net, B)
    outs_d, nodes_d = tree.propagate_batch(net["values"], evars, ev, nodes=True, device_output=True, dense=True)
    outs_p, nodes_p = tree.propagate_batch(net["values"], evars, ev, nodes=True, device_output=True, dense=False)
    for k, (a, b) in enumerate(zip(nodes_d, nodes_p)):
        rel = ((a - b).abs() / b.abs().clamp_min(1e-300)).max().item()
        assert rel < 1e-13, "node %d: %g" % (k, rel)
    want_f, want_n = _oracle(tree, net, evars, ev[:2], 2)
    for k, w in enumerate(want_n):
        assert_close(nodes_d[k][:2].cpu().numpy(), w, RTOL_F64, "node %d" % k)
    del outs_d, nodes_d, outs_p, nodes_p
    tree.clique_tree._engines.clear()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("B", [256, 520])
@pytest.mark.parametrize("name,make", NETS, ids=[n for n, _ in NETS])
def test_dense_float32_storage(name, make, B):
    """The float32 pipeline: rows are stored in float32, the contractions still run on the FP64
    tensor pipe (converted on the way into the fragments, rounded once at the store): <= 1e-5
    against the float64 oracle on float32-rounded inputs, and against the projection kernels."""
    import junctiontree as jt
    from helpers import RTOL_F32
    net = make()
    vals32 = [np.asarray(v, np.float32) for v in net["values"]]
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    ev = wl.draw_evidence(net, B)
    l0 = _native.launch_count()
    outs_d, nodes_d = tree.propagate_batch(vals32, evars, ev, nodes=True, dense=True)
    l1 = _native.launch_count()
    outs_p, nodes_p = tree.propagate_batch(vals32, evars, ev, nodes=True, dense=False)
    l2 = _native.launch_count()
    assert l1 - l0 > l2 - l1, "dense launches did not run"
    assert all(a.dtype == np.float32 for a in outs_d)
    for k, (a, b) in enumerate(zip(list(outs_d) + list(nodes_d), list(outs_p) + list(nodes_p))):
        assert_close(a, b, RTOL_F32, "%s f32 entry %d dense vs projection" % (name, k))
    net64 = dict(net)
    net64["values"] = [np.asarray(v, np.float64) for v in vals32]
    want_f, want_n = _oracle(tree, net64, evars, ev[:3], 3)
    for k, w in enumerate(want_n):
        assert_close(nodes_d[k][:3], w, RTOL_F32, "f32 node %d" % k)
    for f, w in enumerate(want_f):
        assert_close(outs_d[f][:3], w, RTOL_F32, "f32 factor %d" % f)
    # and the pipelines' mode
    o_d = tree.propagate_batch(vals32, evars, ev, dense=True)
    for f, w in enumerate(want_f):
        assert_close(o_d[f][:3], w, RTOL_F32, "f32 outputs-only factor %d" % f)
