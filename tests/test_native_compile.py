"""Host compile phase in C++ (``csrc/jt_compile.cpp`` behind jt_triangulate / jt_junction_tree /
jt_plan_build) against the Python implementations of the same algorithms
(``construction.py``, ``schedule.py``), which stay as the cross-check: identical cliques,
factor assignment, fill-in edges and tree; byte-identical plan blobs.  No GPU needed: the compile
phase makes no CUDA call."""

import numpy as np
import pytest

import jt_workloads as wl
from helpers import compile_net
from junctiontree import _native
from junctiontree import construction as cons
from junctiontree import schedule as sch

NETS = [wl.sprinkler(), wl.huang_darwiche(), wl.wisconsin(), wl.random_dag(12, 3, 2, 3, 8, 5),
        wl.random_dag(16, 3, 2, 4, 6, 11), wl.ising(4), wl.large_state_tree((4, 6, 8, 4, 6, 8)), wl.dag37()]


def _both(net):
    tri_p = cons.find_triangulation(net["factors"], net["sizes"], net.get("order"), impl="python")
    tri_n = cons.find_triangulation(net["factors"], net["sizes"], net.get("order"), impl="native")
    return tri_p, tri_n


@pytest.mark.parametrize("net", NETS + [wl.dag500(), wl.ising(16)], ids=lambda n: n["name"])
def test_native_triangulation_and_tree_equal_python(net):
    (tri_p, mc_p, f2c_p), (tri_n, mc_n, f2c_n) = _both(net)
    assert mc_p == mc_n and list(f2c_p) == list(f2c_n)
    assert [tuple(e) for e in tri_p] == [tuple(e) for e in tri_n]
    tree_p, seps_p = cons.construct_junction_tree(mc_p, net["sizes"], impl="python")
    tree_n, seps_n = cons.construct_junction_tree(mc_n, net["sizes"], impl="native")
    assert seps_p == seps_n and repr(tree_p) == repr(tree_n)
    assert cons.check_running_intersection(tree_n, mc_n + seps_n)
    for root in (0, len(mc_p) - 1):
        a = cons.construct_junction_tree(mc_p, net["sizes"], root=root, impl="python")
        b = cons.construct_junction_tree(mc_p, net["sizes"], root=root, impl="native")
        assert repr(a) == repr(b) and b[0][0] == root


@pytest.mark.parametrize("seed", range(40))
def test_native_compile_fuzz(seed):
    """Random DAGs with integer, string or tuple labels, disconnected components, scalar and
    single-variable factors, repeated factors, user elimination orders."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 18))
    kind = seed % 3
    label = (lambda i: i * 7 % 101) if kind == 0 else (lambda i: "n%d" % i) if kind == 1 else (lambda i: (i % 3, i))
    factors, sizes = [], {}
    for i in range(n):
        sizes[label(i)] = int(rng.integers(1, 5))
        k = int(rng.integers(0, min(i, 3) + 1)) if rng.random() < 0.85 else 0
        parents = sorted(rng.choice(i, size=k, replace=False).tolist()) if k else []
        factors.append([label(p) for p in parents] + [label(i)])
    if seed % 5 == 0:
        factors.append([])                                    # scalar factor
        factors.append(list(factors[0]))                      # repeated factor
    order = None
    if seed % 4 == 1:
        order = [label(int(i)) for i in rng.permutation(n)]
    tri_p = cons.find_triangulation(factors, sizes, order, impl="python")
    tri_n = cons.find_triangulation(factors, sizes, order, impl="native")
    assert tri_p[1] == tri_n[1] and list(tri_p[2]) == list(tri_n[2])
    assert [tuple(e) for e in tri_p[0]] == [tuple(e) for e in tri_n[0]]
    tp = cons.construct_junction_tree(tri_p[1], sizes, impl="python")
    tn = cons.construct_junction_tree(tri_n[1], sizes, impl="native")
    assert repr(tp) == repr(tn)
    if tri_n[1] != [[]]:
        assert cons.check_running_intersection(tn[0], tri_n[1] + tn[1])
    # plan blobs: no evidence, random evidence set, explicit output scopes
    mc, f2c = tri_n[1], tri_n[2]
    tree, seps = tn
    labels = sorted(sizes, key=repr)
    k = int(rng.integers(0, max(1, n // 2)))
    evars = [labels[i] for i in sorted(rng.choice(len(labels), size=k, replace=False).tolist())]
    eff = dict(sizes)
    for v in evars:
        eff[v] = 1
    outputs = [[v] for v in labels if v not in evars][:5] or None
    soft = [v for v in labels if v not in evars][1::3]        # soft evidence on some free variables
    for ev, out, lik in (([], None, ()), (evars, None, ()), (evars, outputs, ()), (evars, None, soft)):
        e = eff if ev else sizes
        if lik and (tree is None or mc == [[]]):
            continue
        a = sch.Plan(tree, mc + seps, e, factors, f2c, ev, sizes, out, emitter="python", likelihood_vars=lik)
        b = sch.Plan(tree, mc + seps, e, factors, f2c, ev, sizes, out, emitter="native", likelihood_vars=lik)
        assert a.lik_entries == sum(e[v] for v in lik) and a.work_entries == a.lik_base + a.lik_entries
        assert a.to_blob() == b.to_blob()
        assert np.array_equal(a.tasks_arr, b.tasks_arr) and np.array_equal(a.tables, b.tables)
        assert np.array_equal(a.msgs_arr, b.msgs_arr) and np.array_equal(a.launches_arr, b.launches_arr)


@pytest.mark.parametrize("net", NETS + [wl.dag500()], ids=lambda n: n["name"])
def test_native_plan_blob_is_byte_identical(net):
    tree, seps, mc, f2c, eff, evars = compile_net(net)
    variants = [
        dict(tree=tree, node_vars=mc + seps, sizes=eff, factors=net["factors"], factor_to_clique=f2c,
             evidence_vars=evars, full_sizes=net["sizes"]),
        dict(tree=tree, node_vars=mc + seps, sizes=net["sizes"], factors=net["factors"], factor_to_clique=f2c),
        dict(tree=tree, node_vars=mc + seps, sizes=net["sizes"]),                      # compute_beliefs: no factors
        dict(tree=None, node_vars=mc, sizes=net["sizes"], factors=net["factors"], factor_to_clique=f2c),  # evaluate
        dict(tree=tree, node_vars=mc + seps, sizes=eff, factors=net["factors"], factor_to_clique=f2c,
             evidence_vars=evars, full_sizes=net["sizes"],
             outputs=[[v] for v in sorted(net["sizes"]) if v not in evars]),
    ]
    for kw in variants:
        a = sch.Plan(emitter="python", **kw)
        b = sch.Plan(emitter="native", **kw)
        assert a.to_blob() == b.to_blob()
        # the blob the library parses is the one it emitted
        dp = _native.DevicePlan(b.to_blob())
        assert dp.query(sch.H_NTASKS) == len(b.tasks_arr) and dp.query(sch.H_NTAB) == b.tables.size
        dp.close()


def test_native_compile_on_hand_built_trees_with_arbitrary_axis_orders():
    """compute_beliefs-style input: user tree, clique and separator axes in any order."""
    node_vars = [["c", "a", "b"], ["d", "c"], ["f", "b", "e"], ["g", "e"], ["c"], ["b"], ["e"]]
    tree = [0, (4, [1]), (5, [2, (6, [3])])]
    sizes = dict(a=3, b=4, c=2, d=5, e=3, f=2, g=7)
    a = sch.Plan(tree, node_vars, sizes, emitter="python")
    b = sch.Plan(tree, node_vars, sizes, emitter="native")
    assert a.to_blob() == b.to_blob()


def test_native_compile_error_reporting():
    net = wl.huang_darwiche()
    with pytest.raises(ValueError):
        cons.find_triangulation(net["factors"], net["sizes"], ["A", "B"], impl="native")
    with pytest.raises(ValueError):
        cons.find_triangulation(net["factors"], net["sizes"], ["A", "B"], impl="python")
    with pytest.raises(_native.NativeError, match="permutation"):
        _native.triangulate([2, 2, 2], [[0, 1], [1, 2]], [0, 0, 1])
    with pytest.raises(_native.NativeError, match="factor lists"):
        _native.triangulate([2, 2], [[0, 5]])
    with pytest.raises(_native.NativeError, match="positive"):
        _native.triangulate([2, 0], [[0, 1]])
    # malformed trees are rejected by the emitter
    with pytest.raises(_native.NativeError, match="separator"):
        _native.plan_build([2, 2, 2], [2, 2, 2], 2, [[0, 1], [1, 2], [0]], ([0, 1], [-1, 0], [-1, 2]),
                           None, None, [], None)
    with pytest.raises(_native.NativeError, match="parents before children"):
        _native.plan_build([2, 2, 2], [2, 2, 2], 2, [[0, 1], [1, 2], [1]], ([0, 1], [-1, 1], [-1, 2]),
                           None, None, [], None)
    with pytest.raises(_native.NativeError, match="effective size 1"):
        _native.plan_build([2, 2, 2], [2, 2, 2], 2, [[0, 1], [1, 2], [1]], ([0, 1], [-1, 0], [-1, 2]),
                           [[0, 1], [1, 2]], [0, 1], [2], None)


def test_five_thousand_variables_compile_quickly():
    import time
    net = wl.random_dag(5000, 3, 2, 4, 8, 3)
    t0 = time.perf_counter()
    _, mc, f2c = cons.find_triangulation(net["factors"], net["sizes"])
    tree, seps = cons.construct_junction_tree(mc, net["sizes"])
    plan = sch.Plan(tree, mc + seps, net["sizes"], net["factors"], f2c)
    dt = time.perf_counter() - t0
    assert cons.check_running_intersection(tree, mc + seps)
    assert plan.n_cliques == len(mc) and dt < 20.0
