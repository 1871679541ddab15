"""GPU parity tests: the sm_100a path (through the C ABI) against the oracles.

Tolerances are BASELINE.json's: relative 1e-12 (float64) / 1e-5 (float32) on every clique and
separator potential and on every per-factor output; evidence slicing and index maps bit-exact.
"""

import numpy as np
import pytest

import jt_workloads as wl
from helpers import RTOL_F32, RTOL_F64, assert_close, load_golden, tuplify

pytestmark = pytest.mark.gpu


def _nets():
    return [
        wl.sprinkler(), wl.huang_darwiche(), wl.wisconsin(),
        wl.random_dag(12, 3, 2, 3, 8, 5), wl.random_dag(16, 3, 2, 4, 6, 11),
        wl.ising(4), wl.large_state_tree((4, 6, 8, 4, 6, 8)),
    ]


def _oracle(tree, net, evars, ev, B):
    from oracle import ref_fixed
    ct = tree.clique_tree
    return ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                     net["factors"], net["sizes"], net["values"], evars, ev, n=B)


@pytest.mark.parametrize("uniform", [True, False], ids=["uniform", "per_instance"])
@pytest.mark.parametrize("net", _nets(), ids=lambda n: n["name"])
@pytest.mark.parametrize("B", [1, 3, 8, 70, 300])
def test_batched_propagation_f64(net, B, uniform):
    import junctiontree as jt
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = net.get("evidence_vars", [])
    ev = wl.draw_evidence(net, B) if evars else None
    outs, nodes = tree.propagate_batch(net["values"], evars, ev, batch=B, nodes=True, uniform=uniform)
    want_f, want_n = _oracle(tree, net, evars, ev, B)
    for k, (g, w) in enumerate(zip(nodes, want_n)):
        assert_close(g, w, RTOL_F64, "node %d" % k)
    for f, (g, w) in enumerate(zip(outs, want_f)):
        assert_close(g, w, RTOL_F64, "factor %d" % f)


@pytest.mark.parametrize("net", _nets(), ids=lambda n: n["name"])
@pytest.mark.parametrize("B", [1, 6, 64, 520])
def test_batched_propagation_f32(net, B):
    """float32 pipeline vs the float64 oracle on float32-rounded inputs."""
    import junctiontree as jt
    net = dict(net)
    net["values"] = [np.asarray(v, np.float32) for v in net["values"]]
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = net.get("evidence_vars", [])
    ev = wl.draw_evidence(net, B) if evars else None
    outs, nodes = tree.propagate_batch(net["values"], evars, ev, batch=B, nodes=True)
    assert all(o.dtype == np.float32 for o in outs)
    net64 = dict(net)
    net64["values"] = [np.asarray(v, np.float64) for v in net["values"]]
    want_f, want_n = _oracle(tree, net64, evars, ev, B)
    for k, (g, w) in enumerate(zip(nodes, want_n)):
        assert_close(g, w, RTOL_F32, "node %d" % k)
    for f, (g, w) in enumerate(zip(outs, want_f)):
        assert_close(g, w, RTOL_F32, "factor %d" % f)


@pytest.mark.parametrize("B", [10, 40, 300])
def test_per_instance_factor_tables(B):
    """Leading batch axis on the factor arrays: every instance has its own tables (B = 10: the
    whole-propagation kernel; 40: LDG kernels; 300: init_rows and TMA kernels)."""
    import junctiontree as jt
    from oracle import ref_fixed
    net = wl.random_dag(10, 3, 2, 3, 8, 3)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    rng = np.random.default_rng(7)
    vals = [rng.random((B,) + v.shape) + 0.05 for v in net["values"]]
    outs, nodes = tree.propagate_batch(vals, nodes=True)
    ct = tree.clique_tree
    for b in sorted(set(range(0, B, max(1, B // 10))) | {B - 1}):
        fo, ys = ref_fixed.propagate(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                     net["factors"], net["sizes"], [v[b] for v in vals])
        for k, y in enumerate(ys):
            assert_close(nodes[k][b], y, RTOL_F64, "node %d instance %d" % (k, b))
        for f, o in enumerate(fo):
            assert_close(outs[f][b], o, RTOL_F64, "factor %d instance %d" % (f, b))


def test_golden_end_to_end_vs_reference():
    """propagate() through our own host compile against the unmodified reference's outputs."""
    import junctiontree as jt
    cases, arrays = load_golden()
    n = 0
    for case in cases:
        if case["kind"] != "end_to_end":
            continue
        values = [arrays[k] for k in case["values"]]
        tree = jt.create_junction_tree(case["factors"], dict(case["sizes"]))
        outs = tree.propagate(values)      # sliced arrays carry the conditioning
        for f, key in enumerate(case["outputs"]):
            if case["outputs_valid"][f]:
                assert_close(outs[f], arrays[key], RTOL_F64, "%s factor %d" % (case["name"], f))
                n += 1
    assert n >= 250


def test_golden_compute_beliefs_on_reference_trees():
    """compute_beliefs on the trees the reference built (its axis orders), vs its outputs."""
    from junctiontree import computation as comp
    cases, arrays = load_golden()
    n = 0
    for case in cases:
        if "beliefs" not in case:
            continue
        if case["kind"] == "operator":
            pots = [arrays[k] for k in case["potentials"]]
            node_vars = case["variables"]
        else:
            node_vars = case["maxcliques"] + case["separators"]
            sizes = dict(case["sizes"])
            sizes.update({v: 1 for v in case["slices"]})
            pots = [arrays[k] for k in case["psi"]] + \
                   [np.ones(tuple(sizes[v] for v in s)) for s in case["separators"]]
        got = comp.compute_beliefs(tuplify(case["tree"]), pots, node_vars)
        for k, key in enumerate(case["beliefs"]):
            if case["beliefs_valid"][k]:
                want = arrays[key]
                want = np.broadcast_to(want, got[k].shape) if want.shape != got[k].shape else want
                assert_close(got[k], want, RTOL_F64, "%s node %d" % (case["name"], k), signed=case["kind"] == "operator")
                n += 1
    assert n >= 320


def test_compute_beliefs_does_not_modify_inputs():
    from junctiontree import computation as comp
    rng = np.random.default_rng(0)
    pots = [rng.standard_normal((2, 3)), rng.standard_normal((3, 4)), np.ones((3,))]
    keep = [p.copy() for p in pots]
    comp.compute_beliefs([0, (2, [1])], pots, [[3, 5], [5, 9], [5]])
    for p, k in zip(pots, keep):
        assert np.array_equal(p, k)


def test_conditioned_readme_example():
    """README conditioning flow (sizes mutated, arrays sliced; reference
    tests/test_junctiontree.py:345-419) -- and the clique the reference gets wrong (D3)."""
    import junctiontree as jt
    from oracle import brute
    net = wl.sprinkler()
    tree = jt.create_junction_tree(net["factors"], dict(net["sizes"]))
    tree.clique_tree.factor_graph.sizes["wet_grass"] = 1
    vals = [v.copy() for v in net["values"]]
    vals[3] = vals[3][:, :, 1:]
    out = tree.propagate(vals)
    marg = out[1].sum(axis=0)
    np.testing.assert_allclose(marg / marg.sum(), [0.57024, 0.42976], atol=0.01)
    truth = brute.joint_marginals(vals, net["factors"], net["factors"])
    for f in range(4):
        assert_close(out[f], truth[f], RTOL_F64, "factor %d" % f)
    tree.clique_tree.factor_graph.sizes["rain"] = 1
    vals[3] = vals[3][1:, :, :]
    vals[2] = vals[2][:, 1:]
    out = tree.propagate(vals)
    marg = out[1].sum(axis=0)
    np.testing.assert_allclose(marg / marg.sum(), [0.8055, 0.1945], atol=0.01)


def test_evidence_slicing_is_bit_exact():
    """Stage V1 + E0: per-instance factor offsets and the initial clique potentials equal NumPy
    slicing exactly (integers and products of the same doubles in the same order)."""
    import torch
    import junctiontree as jt
    from junctiontree import computation as comp
    from oracle import plan_interp
    net = wl.random_dag(14, 3, 2, 4, 8, 9)
    B = 33
    ev = wl.draw_evidence(net, B)
    evars = net["evidence_vars"]
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    plan = tree.plan(evars)
    engine = tree._engine(plan.sizes, evars, plan.full_sizes)
    fdev, _ = engine.factors_to_device(net["values"], np.float64)
    edev = engine.evidence_to_device(ev, B)
    ws = engine.workspace(B, np.float64)
    engine.dev.upload()
    engine.dev.init(fdev.data_ptr(), False, edev.data_ptr(), B, np.float64, ws.data_ptr(), 0, engine._stream())
    torch.cuda.synchronize()
    fbase = engine.evidence_offsets_view(ws, B, np.float64).cpu().numpy()
    assert np.array_equal(fbase, plan_interp.evidence_offsets(plan, ev, B))
    work = engine.work_view(ws, B, np.float64).cpu().numpy()
    # reference semantics: apply_evidence slices, then evaluate multiplies the factors in order
    for b in range(B):
        sliced = [p[0] for p in comp.apply_evidence(net["values"], net["factors"],
                                                    {v: int(ev[b, i]) for i, v in enumerate(evars)})]
        for c in range(plan.n_cliques):
            psi = np.ones(plan.node_shape[c])
            for f in plan.clique_factors[c]:
                axes = [plan.node_vars[c].index(v) for v in net["factors"][f]]
                shape = [1] * len(plan.node_vars[c])
                perm = np.argsort(axes)
                arr = np.transpose(sliced[f], perm)
                for ax, n in zip(sorted(axes), arr.shape):
                    shape[ax] = n
                psi = psi * arr.reshape(shape)
            got = work[plan.node_off[c]:plan.node_off[c] + plan.node_size[c], b].reshape(plan.node_shape[c])
            assert np.array_equal(got, psi), (b, c)


def test_out_of_range_evidence_is_reported():
    import junctiontree as jt
    net = wl.random_dag(8, 2, 2, 3, 8, 1)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    ev = wl.draw_evidence(net, 4)
    ev[2, 0] = 99
    with pytest.raises(ValueError):
        tree.propagate_batch(net["values"], net["evidence_vars"], ev)


def test_evaluate_and_marginalize():
    """CliqueGraph.evaluate / marginalize (reference tests/test_junctiontree.py:9-111)."""
    import junctiontree as jt
    rng = np.random.default_rng(3)
    sizes = {"a": 2, "b": 3, "c": 4, "d": 5, "e": 6}
    factors = [["a", "b"], ["b", "c"], ["c", "d"], ["a", "e"]]
    g = jt.CliqueGraph(maxcliques=[["a", "b", "c"], ["a", "c", "d", "e"], ["a", "d", "e"]],
                       factor_to_maxclique=[0, 0, 1, 2],
                       factor_graph=jt.FactorGraph(factors=factors, sizes=sizes))
    xs = [rng.standard_normal([sizes[v] for v in f]) for f in factors]
    ys = g.evaluate(xs)
    want = [np.einsum("ab,bc->abc", xs[0], xs[1]),
            np.einsum("ae,cd->acde", [[1]], xs[2]),
            np.einsum("d,ae->ade", [1], xs[3])]
    for y, w in zip(ys, want):
        # the reference's shapes: a clique variable no assigned factor covers is a size-1 axis
        # (junctiontree.py:52-61; its tests/test_junctiontree.py:88-109 compare against exactly these einsums)
        assert y.shape == w.shape
        np.testing.assert_allclose(y, w, rtol=RTOL_F64)
    back = g.marginalize(ys)
    full = g.evaluate(xs, reference_shapes=False)          # what the propagation stages work on
    for y, f in zip(ys, full):
        np.testing.assert_array_equal(np.broadcast_to(y, f.shape), f)
    assert_close(back[0], np.einsum("abc->ab", ys[0]), RTOL_F64, signed=True)
    assert_close(back[1], np.einsum("abc->bc", ys[0]), RTOL_F64, signed=True)
    assert_close(back[2], np.einsum("acde->cd", ys[1]), RTOL_F64, signed=True)
    assert_close(back[3], np.einsum("ade->ae", ys[2]), RTOL_F64, signed=True)


def test_sum_product_operator_surface():
    """SumProduct.einsum / project / absorb on the device vs np.einsum."""
    from junctiontree import computation as comp
    rng = np.random.default_rng(5)
    A, B_, C = rng.random((3, 4, 2)), rng.random((4, 2)), rng.random((3,))
    sp = comp.sum_product
    assert sp.on_device
    assert_close(sp.einsum(A, ["a", "b", "c"], B_, ["b", "c"], ["b", "c"]),
                 np.einsum("abc,bc->bc", A, B_), RTOL_F64)
    assert_close(sp.einsum(A, [0, 1, 2], C, [0], []), np.einsum("abc,a->", A, C), RTOL_F64)
    assert_close(sp.einsum(A, [0, 1, 2], C, [0], [2, 0]), np.einsum("abc,a->ca", A, C), RTOL_F64)
    assert_close(sp.project(A, "abc", "ca"), np.einsum("abc->ca", A), RTOL_F64)
    assert_close(sp.absorb(A, "abc", B_, "bc"), A * B_[None], RTOL_F64)
    old = B_.copy()
    old[1, 0] = 0.0
    ratio = np.divide(B_, old, out=np.zeros_like(B_), where=old != 0)
    assert_close(sp.absorb(A, "abc", B_, "bc", old=old), A * ratio[None], RTOL_F64)
    # evidence shrinking equivalence, reference tests/test_computation.py:411-459
    a = np.zeros(3)
    a[2] = 1
    upd = sp.einsum(A, [0, 1, 2], a, [0], [0, 1, 2])
    assert_close(sp.einsum(upd, [0, 1, 2], B_, [1, 2], [1, 2]),
                 sp.einsum(upd[2], [1, 2], B_, [1, 2], [1, 2]), RTOL_F64)
    # diagonal and broadcast
    M = rng.random((4, 4))
    assert_close(sp.einsum(M, [0, 0], [0]), np.diagonal(M), RTOL_F64)
    assert_close(sp.einsum(M, [0, 1], np.ones((1, 4)), [0, 1], [1]), M.sum(axis=0), RTOL_F64)


def test_large_marginal_summation_accuracy():
    """A 2^16-entry clique marginalised to one variable: long sums stay within 1e-12."""
    import junctiontree as jt
    from oracle import ref_fixed
    net = wl.ising(4)
    rng = np.random.default_rng(1)
    factors = [["v%d" % i for i in range(16)], ["v0"]]
    sizes = {"v%d" % i: 2 for i in range(16)}
    values = [rng.random((2,) * 16) + 0.05, rng.random(2) + 0.05]
    tree = jt.create_junction_tree(factors, sizes)
    out = tree.propagate(values)
    want = np.einsum(values[0], list(range(16)), values[1], [0], [0], dtype=np.longdouble)
    assert_close(out[1], np.asarray(want, np.float64), RTOL_F64)


def test_properties_at_scale_dag37():
    """Config 2 shape at a reduced batch: size-independent invariants on every instance --
    every node sums to the same Z, neighbouring nodes agree on separator marginals."""
    import junctiontree as jt
    net = wl.dag37()
    B = 256
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    ev = wl.draw_evidence(net, B)
    outs, nodes = tree.propagate_batch(net["values"], net["evidence_vars"], ev, nodes=True)
    plan = tree.plan(net["evidence_vars"])
    Z = nodes[0].reshape(B, -1).sum(axis=1)
    for k, nd in enumerate(nodes):
        np.testing.assert_allclose(nd.reshape(B, -1).sum(axis=1), Z, rtol=1e-11, err_msg="node %d" % k)
    for c in plan.order:
        for sep, child in plan.children[c]:
            for clique in (c, child):
                cv, sv = plan.node_vars[clique], plan.node_vars[sep]
                axes = tuple(1 + i for i, v in enumerate(cv) if v not in sv)
                marg = nodes[clique].sum(axis=axes)
                kept = [v for v in cv if v in sv]
                marg = np.transpose(marg, [0] + [1 + kept.index(v) for v in sv])
                np.testing.assert_allclose(marg, nodes[sep], rtol=1e-11)
    # and a sample of instances against the oracle
    want_f, _ = _oracle(tree, net, net["evidence_vars"], ev[:4], 4)
    for f, w in enumerate(want_f):
        assert_close(outs[f][:4], w, RTOL_F64, "factor %d" % f)


# ---------------------------------------------------------------------------------------------
# BASELINE.json's configs at (or near) full size: size-independent properties checked on the
# device for every instance, and the first instances against the oracle


def _check_full_size(net, B, dtype, n_oracle, rtol_z, rtol, all_nodes=False, ev_offset=0):
    """``n_oracle`` instances spread over the batch (first, last and evenly between) against the
    NumPy oracle -- every factor output and every node belief (``all_nodes``) or every ~40th --
    and the size-independent invariants on every instance, on the device."""
    import torch
    import junctiontree as jt
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = net.get("evidence_vars", [])
    ev = wl.draw_evidence(net, B + ev_offset)[ev_offset:] if evars else None
    vals = [np.asarray(v, dtype) for v in net["values"]]
    outs, nodes = tree.propagate_batch(vals, evars, ev, batch=B, nodes=True, device_output=True, dtype=dtype)
    plan = tree.plan(evars)
    # every clique and separator belief sums to the same partition function, per instance
    Z = nodes[plan.root].reshape(B, -1).sum(dim=1, dtype=torch.float64)
    assert bool(torch.isfinite(Z).all()) and bool((Z > 0).all())
    for k, nd in enumerate(nodes):
        zk = nd.reshape(B, -1).sum(dim=1, dtype=torch.float64)
        assert float(((zk - Z).abs() / Z).max()) < rtol_z, "node %d" % k
    # separator belief = marginal of the child clique belief (checked on a few edges)
    edges = [(c, s, k) for c in plan.order for s, k in plan.children[c]]
    for c, s, k in edges[:: max(1, len(edges) // 12)]:
        kv, sv = plan.node_vars[k], plan.node_vars[s]
        axes = [1 + i for i, v in enumerate(kv) if v not in sv]
        marg = nodes[k].sum(dim=axes, dtype=torch.float64) if axes else nodes[k].double()
        kept = [v for v in kv if v in sv]
        marg = marg.permute([0] + [1 + kept.index(v) for v in sv])
        ref = nodes[s].double()
        assert float(((marg - ref).abs() / ref.abs().clamp_min(1e-300)).max()) < rtol_z * 10
    # instances spread over the batch against the NumPy oracle
    if n_oracle:
        net64 = dict(net)
        net64["values"] = [np.asarray(v, np.float64) for v in vals]
        pick = sorted(set(int(round(x)) for x in np.linspace(0, B - 1, n_oracle)))
        want_f, want_n = _oracle(tree, net64, evars, ev[pick] if ev is not None else None, len(pick))
        label = "%s B=%d %s" % (net.get("name", "net"), B, np.dtype(dtype).name)
        idx = torch.as_tensor(pick, device="cuda")
        for f, w in enumerate(want_f):
            assert_close(outs[f][idx].cpu().numpy(), w, rtol, "%s factor outputs" % label)
        step = 1 if all_nodes else max(1, len(nodes) // 40)
        for k in range(0, len(nodes), step):
            kind = "clique" if k < plan.n_cliques else "separator"
            assert_close(nodes[k][idx].cpu().numpy(), want_n[k], rtol, "%s %s beliefs" % (label, kind))
    del outs, nodes
    tree.clique_tree._engines.clear()
    torch.cuda.empty_cache()


def test_config3_ising_16x16_float64():
    """Binary 16 x 16 Ising grid, row-sweep order: 240 cliques of up to 2^17 entries, BASELINE's
    batch of 256 in one 142 GB workspace (two halves of 128 with different evidence when the box
    does not have that much free); 4 instances and every clique / separator against the oracle."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free > 150e9:
        _check_full_size(wl.ising(16), 256, np.float64, 4, 1e-11, RTOL_F64, all_nodes=True)
    else:
        _check_full_size(wl.ising(16), 128, np.float64, 2, 1e-11, RTOL_F64, all_nodes=True)
        _check_full_size(wl.ising(16), 128, np.float64, 2, 1e-11, RTOL_F64, all_nodes=True, ev_offset=128)


def test_config4_large_state_tree_float64_and_float32():
    """Six variables of 64-128 states, four 3-variable cliques of 786,432 entries, batch 512."""
    _check_full_size(wl.large_state_tree(), 512, np.float64, 2, 1e-11, RTOL_F64)
    _check_full_size(wl.large_state_tree(), 512, np.float32, 2, 2e-5, RTOL_F32)


def test_config5_dag500_chunk():
    """500-node DAG, 424 cliques up to 1.8M entries: one 1024-instance chunk of the 1M batch."""
    _check_full_size(wl.dag500(), 1024, np.float64, 4, 1e-11, RTOL_F64, all_nodes=True)


def test_config2_dag37_full_batch():
    """Config 2 at its full batch of 65,536 instances."""
    _check_full_size(wl.dag37(), 65536, np.float64, 8, 1e-11, RTOL_F64)


def test_star_tree_with_many_messages():
    """A clique with nine neighbours whose separators all depend on summed-out axes: more
    r-dependent messages than the kernels keep in registers / ring rows (generic paths)."""
    from junctiontree import computation as comp
    from oracle import brute
    rng = np.random.default_rng(11)
    centre = list("abcdefgh")
    sizes = {v: 2 for v in centre}
    node_vars = [centre]
    pots = [rng.random((2,) * 8) + 0.1]
    tree = [0]
    pairs = [("a", "b"), ("b", "c"), ("c", "d"), ("d", "e"), ("e", "f"), ("f", "g"), ("g", "h"), ("h", "a"), ("a", "e")]
    n = len(pairs)
    for i, (u, v) in enumerate(pairs):
        leaf = "z%d" % i
        sizes[leaf] = 3
        node_vars.append([u, leaf, v])
        pots.append(rng.random((2, 3, 2)) + 0.1)
    for i, (u, v) in enumerate(pairs):
        node_vars.append([v, u])                      # separator axis order differs from the cliques
        pots.append(np.ones((2, 2)))
        tree.append((1 + n + i, [1 + i]))
    got = comp.compute_beliefs(tree, pots, node_vars)
    want = brute.tree_beliefs(tree, node_vars, pots)
    for k, (g, w) in enumerate(zip(got, want)):
        assert_close(g, w, RTOL_F64, "node %d" % k)


@pytest.mark.parametrize("B", [1, 5, 64])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_split_r_small_batches(B, dtype):
    """Few instances, long reductions (a 2^13-entry clique with 2- and 4-entry neighbours): the
    split-r kernel, including belief writes and uniform operands."""
    import junctiontree as jt
    rng = np.random.default_rng(4)
    big = ["v%02d" % i for i in range(13)]
    factors = [big, ["v00", "w0"], ["v05", "v06", "w1"], ["w1", "w2"], ["v12"]]
    sizes = {v: 2 for v in big}
    sizes.update(w0=3, w1=2, w2=4)
    values = [np.asarray(rng.random([sizes[v] for v in f]) + 0.05, dtype) for f in factors]
    net = {"factors": factors, "sizes": sizes, "values": values, "evidence_vars": ["w2"], "seed": 3}
    tree = jt.create_junction_tree(factors, sizes)
    ev = wl.draw_evidence(net, B)
    outs, nodes = tree.propagate_batch(values, ["w2"], ev, nodes=True)
    net64 = dict(net, values=[np.asarray(v, np.float64) for v in values])
    want_f, want_n = _oracle(tree, net64, ["w2"], ev, B)
    rtol = RTOL_F64 if dtype == np.float64 else RTOL_F32
    for k, (g, w) in enumerate(zip(nodes, want_n)):
        assert_close(g, w, rtol, "node %d" % k)
    for f, (g, w) in enumerate(zip(outs, want_f)):
        assert_close(g, w, rtol, "factor %d" % f)


def test_large_batch_is_streamed_in_chunks():
    """Host-in / host-out batches above the streaming threshold go through the chunked
    two-stream pipeline (ragged last chunk included) and give the same numbers."""
    import junctiontree as jt
    net = wl.random_dag(14, 3, 2, 3, 8, 2)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    B = 3 * 8192 + 1234
    ev = wl.draw_evidence(net, B)
    outs = tree.propagate_batch(net["values"], net["evidence_vars"], ev)
    assert all(o.shape[0] == B for o in outs)
    direct = tree.propagate_batch(net["values"], net["evidence_vars"], ev, device_output=True)
    for f, (a, b) in enumerate(zip(outs, direct)):
        assert_close(a, b.cpu().numpy(), 1e-14, "factor %d" % f)
    pick = [0, 1, 8191, 8192, 3 * 8192, B - 1]
    want_f, _ = _oracle(tree, net, net["evidence_vars"], ev[pick], len(pick))
    for f, w in enumerate(want_f):
        assert_close(outs[f][pick], w, RTOL_F64, "factor %d" % f)
    ev[B - 3, 1] = -1
    with pytest.raises(ValueError):
        tree.propagate_batch(net["values"], net["evidence_vars"], ev)


def test_marginals_batch_output_stage():
    """Normalised single-variable posteriors and log Z per instance (device output stage)."""
    import junctiontree as jt
    from oracle import brute
    net = wl.random_dag(14, 3, 2, 4, 8, 6)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    B = 37
    ev = wl.draw_evidence(net, B)
    evars = net["evidence_vars"]
    marg, log_z = tree.marginals_batch(net["values"], None, evars, ev)
    free = [v for v in sorted(net["sizes"]) if v not in evars]
    assert sorted(marg) == free and log_z.shape == (B,)
    for b in range(0, B, 6):
        evd = {v: int(ev[b][i]) for i, v in enumerate(evars)}
        truth = brute.factor_graph_marginals(net["factors"], net["values"], [[v] for v in free], evd)
        for v, tr in zip(free, truth):
            assert_close(marg[v][b], tr / tr.sum(), 1e-11, "P(%s | e) instance %d" % (v, b))
            assert_close(log_z[b], np.log(tr.sum()), 1e-11, "log Z instance %d" % b)
    raw, raw_log_z = tree.marginals_batch(net["values"], free[:3], evars, ev, normalize=False)
    for v in free[:3]:
        assert_close(raw[v] / raw[v].sum(axis=1, keepdims=True), marg[v], 1e-12, v)
    # log Z alone (jt_normalize with JT_LOGZ_ONLY): the outputs stay unnormalised, log Z is the same
    assert_close(raw_log_z, log_z, 1e-12, "log Z without normalisation")
    assert_close(np.log(raw[free[0]].sum(axis=1)), log_z, 1e-12, "unnormalised totals")


@pytest.mark.parametrize("seed", range(24))
def test_random_networks_fuzz(seed):
    """Random DAGs (2-5 states, up to 5 parents), random evidence sets (not only leaves), random
    batch sizes crossing the kernel selection thresholds (LDG / split-r / TMA, odd and even B)."""
    import junctiontree as jt
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(5, 15))
    net = wl.random_dag(n, int(rng.integers(1, 6)), 2, int(rng.integers(2, 6)), int(rng.integers(2, n + 1)), 500 + seed)
    labels = sorted(net["sizes"])
    k = int(rng.integers(0, max(1, n // 2)))
    net["evidence_vars"] = sorted(rng.choice(labels, size=k, replace=False).tolist())
    B = int(rng.choice([1, 2, 7, 33, 64, 129, 256, 301, 640]))
    dtype = np.float64 if seed % 3 else np.float32
    vals = [np.asarray(v, dtype) for v in net["values"]]
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    ev = wl.draw_evidence(net, B) if evars else None
    outs, nodes = tree.propagate_batch(vals, evars, ev, batch=B, nodes=True, uniform=bool(seed % 2))
    pick = sorted(set([0, B - 1, B // 2]))
    net64 = dict(net, values=[np.asarray(v, np.float64) for v in vals])
    want_f, want_n = _oracle(tree, net64, evars, ev[pick] if ev is not None else None, len(pick))
    rtol = RTOL_F64 if dtype == np.float64 else RTOL_F32
    for kk, w in enumerate(want_n):
        assert_close(nodes[kk][pick], w, rtol, "node %d" % kk)
    for f, w in enumerate(want_f):
        assert_close(outs[f][pick], w, rtol, "factor %d" % f)


def test_integration_stub_from_the_docs():
    """The ctypes stub printed in INTEGRATION.md section 2 (what a maintainer of the reference
    would add) is executed as is -- only the library path is substituted -- on a junction tree
    object with the reference's attributes, and reproduces the oracle."""
    import os
    import re
    import junctiontree as jt
    from junctiontree import _native
    from oracle import ref_fixed
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# junctiontree/_b200\.py.*?)```", text, flags=re.S).group(1)
    block = block.replace('"libjt_b200.so"', repr(_native.library_path()))
    scope = {}
    exec(compile(block, "INTEGRATION.md", "exec"), scope)
    for net in (wl.sprinkler(), wl.huang_darwiche(), wl.random_dag(12, 3, 2, 3, 8, 5)):
        tree = jt.create_junction_tree(net["factors"], net["sizes"])
        got = scope["propagate"](tree, net["values"])
        ct = tree.clique_tree
        want, _ = ref_fixed.propagate(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                      net["factors"], net["sizes"], net["values"])
        for f, (g, w) in enumerate(zip(got, want)):
            assert_close(g, w, RTOL_F64, "%s factor %d" % (net["name"], f))


def test_propagate_evidence_with_a_different_pattern_per_instance():
    """Batched apply_evidence front-end: every instance has its own set of observed variables;
    each result equals a propagation of the sliced network (reference semantics:
    ``apply_evidence`` slicing ``e:e+1``, ``computation.py:11-34``, then ``propagate``)."""
    import junctiontree as jt
    from junctiontree import computation as comp
    from oracle import ref_fixed
    net = wl.random_dag(13, 3, 2, 4, 8, 21)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    ct = tree.clique_tree
    rng = np.random.default_rng(5)
    labels = sorted(net["sizes"])
    pool = [labels[i] for i in (1, 4, 7, 11)]
    B = 41
    evidence = []
    for b in range(B):
        chosen = [v for v in pool if rng.random() < 0.5]
        evidence.append({v: int(rng.integers(0, net["sizes"][v])) for v in chosen})
    evidence[3] = {}                                   # nothing observed
    got = tree.propagate_evidence(net["values"], evidence)
    assert len(got) == B
    for b in range(B):
        sliced = [a[0] for a in comp.apply_evidence(net["values"], net["factors"], evidence[b])]
        eff = dict(net["sizes"], **{v: 1 for v in evidence[b]})
        want, _ = ref_fixed.propagate(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                      net["factors"], eff, sliced)
        for f, (g, w) in enumerate(zip(got[b], want)):
            assert_close(g, w, RTOL_F64, "instance %d factor %d" % (b, f))
    # the same evidence as an int table with -1 = not observed
    table = np.full((B, len(pool)), -1, np.int64)
    for b, ev in enumerate(evidence):
        for v, s in ev.items():
            table[b, pool.index(v)] = s
    again = tree.propagate_evidence(net["values"], table, variables=pool)
    for b in range(B):
        for g, w in zip(again[b], got[b]):
            assert np.array_equal(g, w)
    with pytest.raises(ValueError):
        tree.propagate_evidence(net["values"], [{"nope": 0}])


@pytest.mark.parametrize("B", [5, 300])
@pytest.mark.parametrize("uniform", [True, False], ids=["uniform", "per_instance"])
def test_soft_evidence_likelihood_vectors(B, uniform):
    """Soft evidence: per-instance likelihood vectors (the workspace's likelihood region) are one
    more single-variable factor per instance -- compared with the oracle run on the extended
    factor graph; combined with hard evidence on other variables."""
    import junctiontree as jt
    from oracle import ref_fixed
    net = wl.random_dag(14, 3, 2, 4, 8, 6)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    ct = tree.clique_tree
    evars = net["evidence_vars"]
    ev = wl.draw_evidence(net, B)
    rng = np.random.default_rng(8)
    free = [v for v in sorted(net["sizes"]) if v not in evars]
    lik = {v: rng.random((B, net["sizes"][v])) + 0.05 for v in (free[0], free[3], free[7])}
    outs, nodes = tree.propagate_batch(net["values"], evars, ev, nodes=True, uniform=uniform, likelihoods=lik)
    for b in sorted(set([0, B // 2, B - 1])):
        fx, f2cx, vx = ref_fixed.with_likelihood_factors(net["factors"], ct.factor_to_maxclique, ct.maxcliques,
                                                         net["values"], lik, b)
        want_f, want_n = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, f2cx, fx, net["sizes"],
                                                   vx, evars, ev[b:b + 1], n=1)
        for k, w in enumerate(want_n):
            assert_close(nodes[k][b], w[0], RTOL_F64, "node %d instance %d" % (k, b))
        for f in range(len(net["factors"])):
            assert_close(outs[f][b], want_f[f][0], RTOL_F64, "factor %d instance %d" % (f, b))


def test_soft_evidence_one_hot_equals_slicing_and_streams():
    """One-hot likelihoods reproduce hard evidence (reference tests/test_computation.py:411-459:
    slicing == one-hot); the chunked pipeline carries likelihoods; log-domain laws take
    log-likelihoods; posteriors through marginals_batch."""
    import junctiontree as jt
    from junctiontree import semirings as sr
    net = wl.random_dag(12, 3, 2, 3, 8, 5)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    B = 4096 + 500                                     # above the streaming threshold
    ev = wl.draw_evidence(net, B)
    onehot = {v: np.eye(net["sizes"][v])[ev[:, i]] for i, v in enumerate(evars)}
    hard = tree.propagate_batch(net["values"], evars, ev)
    soft = tree.propagate_batch(net["values"], likelihoods=onehot)
    for f, fv in enumerate(net["factors"]):
        for b in (0, 4095, 4096, B - 1):
            ix = tuple(slice(int(ev[b, evars.index(v)]), int(ev[b, evars.index(v)]) + 1) if v in evars else slice(None)
                       for v in fv)
            assert_close(soft[f][b][ix], hard[f][b], RTOL_F64, "factor %d instance %d" % (f, b))
            mask = np.ones(soft[f][b].shape, bool)
            mask[ix] = False
            assert np.all(soft[f][b][mask] == 0.0)             # no belief outside the observed state
    free = [v for v in sorted(net["sizes"]) if v not in evars]
    post_h, logz_h = tree.marginals_batch(net["values"], free, evars, ev)
    post_s, logz_s = tree.marginals_batch(net["values"], free, likelihoods=onehot)
    assert_close(logz_s, logz_h, 1e-11, "log Z")
    for v in free:
        assert_close(post_s[v], post_h[v], 1e-11, "posterior %s" % v)
    with np.errstate(divide="ignore"):
        logv = [np.log(x) for x in net["values"]]
        post_l, logz_l = tree.marginals_batch(logv, free, likelihoods={v: np.log(x) for v, x in onehot.items()},
                                              dl=sr.log_sum_exp)
    assert_close(logz_l, logz_h, 1e-11, "log Z (log domain)")
    with pytest.raises(ValueError):
        tree.propagate_batch(net["values"], evars, ev, likelihoods={evars[0]: onehot[evars[0]]})
    with pytest.raises(ValueError):
        tree.propagate_batch(net["values"], likelihoods={"nope": np.ones((B, 2))})


def test_small_propagations_run_as_one_launch():
    """A few instances of a small tree: init + collect + distribute + marginal are one launch of
    the whole-propagation kernel (one CTA per instance), with and without clique beliefs, with
    evidence, in every semiring -- and give the oracle's numbers."""
    import junctiontree as jt
    from junctiontree import _native, semirings as sr
    from oracle import ref_fixed
    net = wl.random_dag(12, 3, 2, 3, 8, 5)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    ct = tree.clique_tree
    evars = net["evidence_vars"]
    for B in (1, 5, 16):
        ev = wl.draw_evidence(net, B)
        for nodes in (False, True):
            before = _native.launch_count()
            res = tree.propagate_batch(net["values"], evars, ev, nodes=nodes)
            assert _native.launch_count() - before == 1
            outs, got_nodes = res if nodes else (res, None)
            want_f, want_n = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                                       net["factors"], net["sizes"], net["values"], evars, ev, n=B)
            for f, (g, w) in enumerate(zip(outs, want_f)):
                assert_close(g, w, RTOL_F64, "factor %d" % f)
            if nodes:
                for k, (g, w) in enumerate(zip(got_nodes, want_n)):
                    assert_close(g, w, RTOL_F64, "node %d" % k)
    before = _native.launch_count()
    tree.propagate_batch(net["values"], evars, wl.draw_evidence(net, 17))
    assert _native.launch_count() - before > 1                   # larger batches: per-level launches
    ev = wl.draw_evidence(net, 4)
    ev[2, 0] = 99
    with pytest.raises(ValueError):
        tree.propagate_batch(net["values"], evars, ev)           # out-of-range evidence is still reported
    before = _native.launch_count()
    got = tree.propagate_batch(net["values"], evars, wl.draw_evidence(net, 3), dl=sr.max_product)
    assert _native.launch_count() - before == 1
    want_f, _ = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                          net["factors"], net["sizes"], net["values"], evars,
                                          wl.draw_evidence(net, 3), n=3, semiring="max_product")
    for f, (g, w) in enumerate(zip(got, want_f)):
        assert_close(g, w, RTOL_F64, "max-product factor %d" % f)


def test_per_level_kernels_at_tiny_batches_without_the_whole_propagation_kernel():
    """JT_DISABLE_WALK=1 keeps the per-level launches for tiny batches: the LDG and split-r
    kernels at B = 1..8 stay covered (the variable is read once per process, hence a subprocess)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, JT_DISABLE_WALK="1")
    sel = "test_batched_propagation_f64 or test_batched_propagation_f32 or test_golden or test_split_r or test_soft_evidence"
    run = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-k", sel, "-p", "no:cacheprovider",
                          os.path.join(root, "tests", "test_gpu_parity.py")], env=env, cwd=root,
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-2000:]


def test_streaming_pipeline_on_sparse_workspaces():
    """The chunked pipeline backs only the rows its stages touch (uniform mode, no clique
    beliefs) with device memory -- CUDA virtual memory management, jt_workspace_sparse_* -- and
    gives the same numbers as the dense path, for factor scopes and for the output stage."""
    import junctiontree as jt
    from junctiontree import _native
    net = wl.large_state_tree((12, 16, 20, 12, 16, 20))
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    B = 4096 + 600
    ev = wl.draw_evidence(net, B)
    plan = tree.plan(evars)
    engine = tree._engine(plan.sizes, evars, plan.full_sizes)
    mapped, dense = engine.dev.sparse_bytes(4096, np.float64, engine.PIPELINE_FLAGS)
    assert mapped < 0.5 * dense
    pipe = engine.pipeline(B, np.float64, chunk=4096)
    assert isinstance(pipe.slots[0]["ws"], _native.SparseWorkspace)
    assert pipe.slots[0]["ws"].mapped_bytes == mapped
    del pipe
    outs = tree.propagate_batch(net["values"], evars, ev)                 # streamed: B > 4096
    direct = tree.propagate_batch(net["values"], evars, ev[:64], nodes=True)[0]   # dense workspace, all beliefs
    for f, (a, b) in enumerate(zip(outs, direct)):
        assert_close(a[:64], b, 1e-13, "factor %d" % f)
    pick = [0, 4095, 4096, B - 1]
    want_f, _ = _oracle(tree, net, evars, ev[pick], len(pick))
    for f, w in enumerate(want_f):
        assert_close(outs[f][pick], w, RTOL_F64, "factor %d" % f)
    marg, log_z = tree.marginals_batch(net["values"], None, evars, ev)
    dense_marg, dense_log_z = tree.marginals_batch(net["values"], None, evars, ev[:40])   # one small dense chunk
    assert_close(log_z[:40], dense_log_z, 1e-12, "log Z")
    for v in marg:
        assert_close(marg[v][:40], dense_marg[v], 1e-12, "posterior %s" % v)
    ev[B - 2, 0] = 77
    with pytest.raises(ValueError):
        tree.propagate_batch(net["values"], evars, ev)                    # the error counter lives in the mapped tail


def test_marginals_session_reuses_the_pipeline():
    """Serving: one session, several batches of the same size -- same numbers as marginals_batch,
    fresh result arrays per call, soft evidence and out-of-range evidence handled."""
    import junctiontree as jt
    net = wl.random_dag(14, 3, 2, 4, 8, 6)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    B = 4096 + 200
    free = [v for v in sorted(net["sizes"]) if v not in evars]
    soft = [free[2]]
    rng = np.random.default_rng(3)
    with tree.marginals_session(net["values"], B, free, evars, likelihood_vars=soft) as session:
        results = []
        for seed in (11, 12):
            ev = wl.draw_evidence(net, B, seed=seed)
            lik = {soft[0]: rng.random((B, net["sizes"][soft[0]])) + 0.1}
            got, log_z = session.run(ev, lik)
            want, want_z = tree.marginals_batch(net["values"], free, evars, ev, likelihoods=lik)
            assert_close(log_z, want_z, 1e-13, "log Z")
            for v in free:
                assert_close(got[v], want[v], 1e-13, "posterior %s" % v)
            results.append(got)
        assert not np.shares_memory(results[0][free[0]], results[1][free[0]])
        bad = wl.draw_evidence(net, B)
        bad[7, 0] = -3
        with pytest.raises(ValueError):
            session.run(bad, lik)
    with pytest.raises(RuntimeError):
        session.run(ev, lik)


def test_readme_usage_snippet_runs():
    """The usage block of README.md, executed as printed on the README network of the reference."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "README.md")).read()
    block = re.search(r"```python\n(import junctiontree as jt.*?)```", text, flags=re.S).group(1)
    net = wl.sprinkler()
    B = 65536
    rng = np.random.default_rng(0)
    scope = {"factors": net["factors"], "var_sizes": net["sizes"], "values": net["values"],
             "states": rng.integers(0, 2, size=(B, 1)).astype(np.int32), "lam": rng.random((B, 2)) + 0.1}
    exec(compile(block, "README.md", "exec"), scope)
    assert len(scope["out"]) == 4 and scope["outs"][0].shape == (B, 2)
    assert np.allclose(scope["post"]["rain"].sum(axis=1), 1.0) and scope["log_z"].shape == (B,)
    assert np.allclose(scope["best"]["rain"].max(axis=1), 1.0)
    assert len(scope["per_instance"]) == 3 and scope["per_instance"][1][2].shape == (1, 1)
    assert scope["soft"][3].shape == (B, 2, 2, 2)


def test_propagate_session_returns_factor_beliefs():
    """propagate_session: the per-factor beliefs of propagate_batch through a reusable pipeline,
    as copies or as views of the session's pinned buffer."""
    import junctiontree as jt
    net = wl.random_dag(14, 3, 2, 3, 8, 2)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    evars = net["evidence_vars"]
    B = 4096 + 100
    ev = wl.draw_evidence(net, B)
    want = tree.propagate_batch(net["values"], evars, ev)
    with tree.propagate_session(net["values"], B, evars, chunk=2048) as session:
        got = session.run(ev)
        views = session.run(ev, copy=False)
        for f, (g, v, w) in enumerate(zip(got, views, want)):
            assert_close(g, w, 1e-13, "factor %d" % f)
            assert np.array_equal(g, v) and not np.shares_memory(g, v)


@pytest.mark.parametrize("B", [1, 70, 300])
def test_edge_cases_disconnected_scalar_single_and_impossible_evidence(B):
    """Edge cases of the reference's domain against brute force over the joint: unconnected
    components (empty separators, scalar messages: every output is scaled by the other
    components' partition sums, reference construction.py:530), a scalar factor, a single-clique
    net, zeros in the tables, evidence of probability zero (Z = 0: normalised output all zeros,
    log Z = -inf)."""
    import junctiontree as jt
    from oracle import brute
    rng = np.random.default_rng(B)
    # three components, one of them a single variable, plus a scalar factor
    factors = [["x", "y"], ["y", "w"], ["z"], ["u", "v"], []]
    sizes = dict(x=2, y=3, w=2, z=4, u=3, v=2)
    values = [rng.random([sizes[v] for v in f]) + 0.1 for f in factors[:-1]] + [np.array(2.5)]
    values[3][1, :] = 0.0                                    # zeros in a table
    tree = jt.create_junction_tree(factors, sizes)
    assert any(len(s) == 0 for s in tree.separators)         # components are joined by empty separators
    ev = rng.integers(0, 3, size=(B, 1)).astype(np.int32)    # evidence on u; state 1 has probability zero
    ev[0, 0] = 1
    outs = tree.propagate_batch(values, ["u"], ev)
    for b in sorted(set([0, B // 2, B - 1])):
        truth = brute.factor_graph_marginals(factors, values, factors, {"u": int(ev[b, 0])})
        for f, (g, w) in enumerate(zip(outs, truth)):
            assert_close(g[b], w, RTOL_F64, "factor %d instance %d" % (f, b))
    assert np.all(outs[0][0] == 0.0)                         # impossible evidence: every belief is zero
    marg, log_z = tree.marginals_batch(values, ["x", "z"], ["u"], ev)
    assert np.isneginf(log_z[0]) and np.all(marg["x"][0] == 0.0)
    ok = ev[:, 0] != 1
    assert np.allclose(marg["z"][ok].sum(axis=1), 1.0) and np.all(np.isfinite(log_z[ok]))
    if B == 1:
        # single propagate() calls: a one-clique network and a scalar-only factor graph
        single = jt.create_junction_tree([["a", "b"]], dict(a=2, b=3))
        table = rng.random((2, 3))
        assert_close(single.propagate([table])[0], table, RTOL_F64, "single clique")
        scalar = jt.create_junction_tree([[]], {})
        assert_close(scalar.propagate([np.array(3.0)])[0], np.array(3.0), RTOL_F64, "scalar factor graph")


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("B", [40, 70, 200, 600])
def test_clique_initialisation_with_one_to_eight_factors_per_clique(B, dtype):
    """The clique initialisation instantiates its row loop per factor count (1-6 gathered factors;
    more than six take the generic loop) and for blocks that span 1-4 rows of a narrow batch: seven
    unconnected triples whose single clique holds k = 1..6 and 8 factors, per-instance mode, evidence
    on one variable of every triple (so every factor with it is gathered per instance).  Brute force
    per component: the components only scale each other by their evidence likelihoods."""
    import junctiontree as jt
    from oracle import brute
    rng = np.random.default_rng(100 + B)
    shapes = [["x", "y", "z"], ["x"], ["y", "z"], ["x", "z"], ["z"], ["x", "y"], ["x", "y", "z"], ["y"]]
    factors, sizes, comp = [], {}, []
    for c, k in enumerate([1, 2, 3, 4, 5, 6, 8]):
        names = {v: "%s%d" % (v, c) for v in "xyz"}
        sizes.update({names["x"]: 2, names["y"]: 3, names["z"]: 2 + c % 2})
        for f in shapes[:k]:
            factors.append([names[v] for v in f])
            comp.append(c)
    np_dtype = np.float64 if dtype == "f64" else np.float32
    values = [(rng.random([sizes[v] for v in f]) + 0.25).astype(np_dtype) for f in factors]
    tree = jt.create_junction_tree(factors, sizes)
    counts = sorted(np.bincount(tree.clique_tree.factor_to_maxclique).tolist())
    assert counts == [1, 2, 3, 4, 5, 6, 8], counts
    evars = ["z%d" % c for c in range(7)]
    ev = np.stack([rng.integers(0, sizes[v], size=B) for v in evars], axis=1).astype(np.int32)
    outs = tree.propagate_batch(values, evars, ev, uniform=False)
    rtol = RTOL_F64 if dtype == "f64" else 2e-5
    vals64 = [np.asarray(v, np.float64) for v in values]
    for b in sorted(set([0, 1, B // 2, B - 1])):
        evidence = {v: int(ev[b, j]) for j, v in enumerate(evars)}
        # partition sum of every component under this instance's evidence
        z = []
        for c in range(7):
            idx = [i for i in range(len(factors)) if comp[i] == c]
            total = brute.factor_graph_marginals([factors[i] for i in idx], [vals64[i] for i in idx], [[]],
                                                 {"z%d" % c: evidence["z%d" % c]})[0]
            z.append(float(total))
        for c in range(7):
            idx = [i for i in range(len(factors)) if comp[i] == c]
            truth = brute.factor_graph_marginals([factors[i] for i in idx], [vals64[i] for i in idx],
                                                 [factors[i] for i in idx], {"z%d" % c: evidence["z%d" % c]})
            others = float(np.prod([z[o] for o in range(7) if o != c]))
            for i, w in zip(idx, truth):
                assert_close(outs[i][b], w * others, rtol, "factor %d (component %d) instance %d" % (i, c, b))
