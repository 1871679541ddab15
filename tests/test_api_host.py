"""Host-side API mirror (CPU only): names and call surface of the reference, evidence slicing,
the SumProduct plugin hook, and loud failure of the compute path without a CUDA device."""

import numpy as np
import pytest

import junctiontree as jt
from helpers import RTOL_F64, assert_close, load_golden, tuplify
from junctiontree import computation as comp
from junctiontree import sum_product as sp_mod
from junctiontree._native import NativeError


def _no_cuda():
    import torch
    return not torch.cuda.is_available()


def test_public_names_of_the_reference_exist():
    """reference junctiontree/__init__.py re-exports junctiontree.junctiontree.*"""
    for name in ("create_junction_tree", "FactorGraph", "CliqueGraph", "JunctionTree", "einsum"):
        assert hasattr(jt, name)
    for name in ("compute_beliefs", "apply_evidence", "sum_product"):
        assert hasattr(comp, name)
    assert isinstance(comp.sum_product, sp_mod.SumProduct) and comp.sum_product.on_device
    import inspect
    assert list(inspect.signature(comp.compute_beliefs).parameters) == ["tree", "potentials", "clique_vars", "dl"]
    assert list(inspect.signature(comp.apply_evidence).parameters) == ["potentials", "variables", "evidence"]


def test_create_junction_tree_structure_and_assertion():
    sizes = {"cloudy": 2, "sprinkler": 2, "rain": 2, "wet_grass": 2}
    factors = [["cloudy"], ["cloudy", "sprinkler"], ["cloudy", "rain"], ["rain", "sprinkler", "wet_grass"]]
    tree = jt.create_junction_tree(factors, sizes)
    assert isinstance(tree, jt.JunctionTree)
    ct = tree.clique_tree
    assert ct.factor_graph.factors == factors and ct.factor_graph.sizes is sizes
    assert len(ct.maxcliques) == 2 and len(tree.separators) == 1
    assert sorted(tree.separators[0]) == ["rain", "sprinkler"]
    assert tree.tree[0] in (0, 1) and tree.tree[1][0] == 2
    assert len(ct.factor_to_maxclique) == 4
    with pytest.raises(AssertionError):
        jt.create_junction_tree([("a", "b")], {"a": 2, "b": 2})       # reference junctiontree.py:14
    plan = tree.plan()
    assert plan.n_cliques == 2 and 8 * plan.algorithmic_entries() == 640


def test_apply_evidence_slices_views_bit_exact():
    """reference computation.py:11-34, including the one-element list wrapping (D4)."""
    rng = np.random.default_rng(0)
    pots = [rng.standard_normal((2, 3, 6)), rng.standard_normal((3, 4)), rng.standard_normal((2, 5)),
            np.ones((3,)), 2.5]
    variables = [[3, 5, 7], [5, 9], [3, 1], [5], []]
    out = comp.apply_evidence(pots, variables, {3: 0, 9: 2})
    assert all(isinstance(o, list) and len(o) == 1 for o in out)
    assert np.array_equal(out[0][0], pots[0][0:1, :, :]) and out[0][0].base is pots[0]
    assert np.array_equal(out[1][0], pots[1][:, 2:3])
    assert np.array_equal(out[2][0], pots[2][0:1, :])
    assert np.array_equal(out[3][0], pots[3])
    assert out[4][0] == 2.5


def test_sum_product_plugin_hook_remaps_labels():
    """SumProduct(einsum, *args, **kwargs) forwards to the injected function with integer
    sublists (reference sum_product.py:6-35)."""
    calls = []

    def fake_einsum(*args, **kwargs):
        calls.append((args, kwargs))
        return np.einsum(*args[:-1]) if kwargs.get("drop_last") else np.einsum(*args)

    sp = sp_mod.SumProduct(np.einsum, optimize=True)
    assert not sp.on_device
    A, B = np.random.rand(3, 4), np.random.rand(4, 5)
    np.testing.assert_allclose(sp.einsum(A, ["x", "y"], B, ["y", ("z", 1)], ["x", ("z", 1)]), A @ B)
    sp2 = sp_mod.SumProduct(fake_einsum, "extra", flag=1)
    sp2.einsum(A, ["a", "b"], ["b"], drop_last=True)
    (args, kwargs), = calls
    assert args[1] == [0, 1] and args[2] == [1] and args[3] == "extra" and kwargs == {"drop_last": True, "flag": 1}
    # implicit-output form only for label-free operands, like the reference (D16)
    assert sp.einsum(3.0, []) == 3.0


def test_compute_beliefs_through_an_injected_distributive_law_matches_the_reference():
    """The plugin path calls dl.einsum once per operator (collect, exclude-one distribute,
    separator and clique beliefs); with np.einsum injected it must reproduce the reference's
    recorded outputs.  This pins the orchestration order on the CPU."""
    cases, arrays = load_golden()
    dl = sp_mod.SumProduct(np.einsum)
    checked = 0
    for case in cases:
        if "beliefs" not in case:
            continue
        if case["kind"] == "operator":
            pots = [arrays[k] for k in case["potentials"]]
            node_vars = case["variables"]
        else:
            continue
        got = comp.compute_beliefs(tuplify(case["tree"]), pots, node_vars, dl)
        for k, key in enumerate(case["beliefs"]):
            if case["beliefs_valid"][k]:
                assert_close(got[k], arrays[key], RTOL_F64, "%s node %d" % (case["name"], k), signed=case["kind"] == "operator")
                checked += 1
        for p, key in zip(pots, case["potentials"]):
            assert np.array_equal(p, arrays[key])          # inputs untouched
    assert checked >= 30


@pytest.mark.skipif(not _no_cuda(), reason="checks the behaviour on a machine without a GPU")
def test_compute_path_fails_loudly_without_cuda():
    sizes = {"a": 2, "b": 3}
    tree = jt.create_junction_tree([["a", "b"], ["b"]], sizes)
    vals = [np.ones((2, 3)), np.ones(3)]
    with pytest.raises(NativeError):
        tree.propagate(vals)
    with pytest.raises(NativeError):
        tree.propagate_batch(vals, batch=4)
    with pytest.raises(NativeError):
        tree.clique_tree.evaluate(vals)
    with pytest.raises(NativeError):
        comp.compute_beliefs([0], [np.ones((2, 3))], [["a", "b"]])
    with pytest.raises(NativeError):
        comp.sum_product.einsum(np.ones((2, 3)), [0, 1], [0])


def test_no_numpy_compute_in_the_product_package():
    """The package must not route the hot path through NumPy (np.einsum & co)."""
    import os
    import re
    pkg = os.path.dirname(jt.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            code = re.sub(r'""".*?"""|\'\'\'.*?\'\'\'|#.*', "", text, flags=re.S)
            assert "np.einsum" not in code and "numpy.einsum" not in code, fn
            assert "oracle" not in code, fn
            assert not re.search(r"np\.(tensordot|dot|matmul|sum|prod)\(", code), fn


def test_shape_validation_errors():
    sizes = {"a": 2, "b": 3}
    tree = jt.create_junction_tree([["a", "b"], ["b"]], sizes)
    with pytest.raises(ValueError):
        jt.junctiontree._effective_sizes([["a", "b"], ["b"]], [np.ones((2, 3)), np.ones(4)])
    with pytest.raises(ValueError):
        jt.junctiontree._effective_sizes([["a", "b"]], [np.ones((2, 3, 1))])
    from junctiontree import schedule as sch
    with pytest.raises(ValueError):      # separator not contained in its cliques
        sch.Plan([0, (2, [1])], [["a", "b"], ["b", "c"], ["a"]], {"a": 2, "b": 2, "c": 2})
    with pytest.raises(ValueError):      # observed variable must have effective size 1
        sch.Plan([0], [["a"]], {"a": 2}, [["a"]], [0], ["a"], {"a": 2})


def test_two_trees_over_one_clique_graph_get_their_own_plans():
    """the engine cache is keyed by the tree (root, shape), the separator axis orders,
    the sizes and the device -- a second JunctionTree over the same CliqueGraph must not reuse the
    first one's compiled plan."""
    import attr
    sizes = {"a": 2, "b": 3, "c": 4, "d": 2, "e": 3}
    tree = jt.create_junction_tree([["a", "b", "c"], ["b", "c", "d"], ["d", "e"]], sizes)
    ct = tree.clique_tree
    assert any(len(s) == 2 for s in tree.separators)
    assert attr.fields(type(ct)) and type(ct).__attrs_attrs__[0].name == "maxcliques"
    with pytest.raises(attr.exceptions.FrozenInstanceError):       # frozen, as in the reference (junctiontree.py:120)
        ct.maxcliques = []
    plan_a = tree.plan()
    # the same clique graph, re-rooted at another clique
    from junctiontree import construction as cons
    new_root = next(c for c in range(len(ct.maxcliques)) if c != tree.tree[0])
    tree_b, seps_b = cons.construct_junction_tree(ct.maxcliques, sizes, root=new_root)
    assert tree_b[0] == new_root != tree.tree[0]
    other = jt.JunctionTree(tree=tree_b, separators=seps_b, clique_tree=ct)
    plan_b = other.plan()
    assert plan_b is not plan_a and plan_b.root == new_root and plan_a.root == tree.tree[0]
    # separators listed with reversed axis order: a different plan again (other separator shapes)
    flipped = [list(reversed(s)) for s in tree.separators]
    third = jt.JunctionTree(tree=tree.tree, separators=flipped, clique_tree=ct)
    plan_c = third.plan()
    assert plan_c is not plan_a
    assert [list(v) for v in plan_c.node_vars[len(ct.maxcliques):]] == flipped
    assert tree.plan() is plan_a                                    # and the first tree still gets its own
    # caches live outside the frozen objects and go away with them
    import gc
    key = id(ct)
    assert key in jt.junctiontree._caches
    del tree, other, third, ct, plan_a, plan_b, plan_c
    gc.collect()
    assert key not in jt.junctiontree._caches


def test_engine_and_workspace_caches_are_bounded(monkeypatch):
    """Every distinct evidence pattern (propagate_evidence) compiles its own plan and every batch
    size gets its own workspace: both caches drop their least recently used entry instead of
    growing with the number of patterns / sizes ever seen."""
    from junctiontree import engine as eng
    import jt_workloads as wl
    net = wl.random_dag(12, 3, 2, 3, 8, 5)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    monkeypatch.setattr(jt.junctiontree, "_MAX_ENGINES", 3)
    labels = sorted(net["sizes"])
    plans = [tree.plan([v]) for v in labels[:3]]
    assert len(tree.clique_tree._engines) == 3
    assert tree.plan([labels[0]]) is plans[0]                 # a hit: now the most recently used
    fourth = tree.plan([labels[3]])
    assert len(tree.clique_tree._engines) == 3
    assert tree.plan([labels[0]]) is plans[0]                 # survived, the oldest (labels[1]) went
    assert tree.plan([labels[3]]) is fourth
    assert tree.plan([labels[1]]) is not plans[1]             # recompiled
    assert len(tree.clique_tree._engines) == 3

    engine = tree._engine(plans[0].sizes, [labels[0]], plans[0].full_sizes)
    made = []
    monkeypatch.setattr(eng.Engine, "new_workspace", lambda self, B, dtype: made.append(B) or object())
    first = engine.workspace(8, np.float64)
    for B in (16, 32, 64):
        engine.workspace(B, np.float64)
    assert engine.workspace(8, np.float64) is first and made == [8, 16, 32, 64]
    engine.workspace(128, np.float64)                         # fifth size: the least recently used (16) goes
    assert engine.workspace(8, np.float64) is first
    assert sum(isinstance(k[0], int) for k in engine._workspaces) == eng.Engine.MAX_CACHED_WORKSPACES
    engine.workspace(16, np.float64)
    assert made == [8, 16, 32, 64, 128, 16]
    engine._workspaces[("graph", 1)] = "runner"               # tagged entries (runners) are not counted or evicted
    for B in (256, 512, 1024, 2048, 4096):
        engine.workspace(B, np.float32)
    assert engine._workspaces[("graph", 1)] == "runner"
    assert sum(isinstance(k[0], int) for k in engine._workspaces) == eng.Engine.MAX_CACHED_WORKSPACES
