"""Host compile: triangulation, maximal cliques, junction tree (CPU only).

Behaviour spec: the reference's tests/test_construction.py (chordality, maximal cliques of the
textbook graph, integer labels, no-edge graphs) plus the properties the reference violates on
random inputs (running intersection)."""

import json
import os
import subprocess
import sys

import numpy as np
import pytest

import jt_workloads as wl
from junctiontree import construction as cons

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(factors, sizes, order=None):
    tri, mc, f2c = cons.find_triangulation(factors, sizes, order)
    tree, seps = cons.construct_junction_tree(mc, sizes)
    return tri, mc, f2c, tree, seps


def _is_chordal(variables, edges):
    """Perfect elimination ordering exists (simplicial vertices can be removed one by one)."""
    adj = {v: set() for v in variables}
    for a, b in edges:
        adj[a].add(b)
        adj[b].add(a)
    left = set(variables)
    while left:
        for v in list(left):
            nb = adj[v] & left
            if all(b in adj[a] for a in nb for b in nb if a != b):
                left.remove(v)
                break
        else:
            return False
    return True


@pytest.mark.parametrize("seed", range(40))
def test_random_dags_give_valid_junction_trees(seed):
    """The reference loses running intersection on 13-52 % of such inputs (defect D8)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(6, 22))
    net = wl.random_dag(n, int(rng.integers(1, 5)), 2, 4, int(rng.integers(3, n + 1)), seed)
    tri, mc, f2c, tree, seps = _compile(net["factors"], net["sizes"])
    nodes = mc + seps
    assert cons.check_running_intersection(tree, nodes)
    # every factor sits inside its clique, cliques are maximal, separator ids follow the cliques
    for fv, home in zip(net["factors"], f2c):
        assert set(fv) <= set(mc[home])
    for i, a in enumerate(mc):
        assert not any(set(a) < set(b) for j, b in enumerate(mc) if i != j)
    order, parent, parent_sep, depth, children = cons.tree_edges(tree)
    assert sorted(order) == list(range(len(mc)))
    assert sorted(s for c in order for s, _ in children[c]) == list(range(len(mc), len(nodes)))
    # triangulated moral graph is chordal
    variables = sorted(net["sizes"])
    edges = [tuple(e) for e in cons.factors_to_undirected_graph(net["factors"])] + list(tri)
    assert _is_chordal(variables, edges)


def test_four_cycle_gets_one_fill_edge():
    factors = [["a", "b"], ["b", "c"], ["c", "d"], ["d", "a"]]
    sizes = {v: 2 for v in "abcd"}
    tri, mc, f2c, tree, seps = _compile(factors, sizes)
    assert len(tri) == 1
    assert sorted(len(c) for c in mc) == [3, 3]
    assert len(seps) == 1 and len(seps[0]) == 2


def test_textbook_maximal_cliques():
    """Huang & Darwiche's network: six 3-variable cliques (reference
    tests/test_junctiontree.py:149-161)."""
    net = wl.huang_darwiche()
    _, mc, f2c, tree, seps = _compile(net["factors"], net["sizes"])
    # min-fill ties may pick another (equally small) triangulation than the textbook's
    # ABD ACE ADE CEG DEF EGH; the shape is fixed: six triangles, five 2-variable separators
    assert sorted(len(c) for c in mc) == [3] * 6
    assert sorted(len(s) for s in seps) == [2, 2, 2, 2, 2]
    for must in (("D", "E", "F"), ("E", "G", "H")):      # the two 3-variable factors are cliques
        assert must in set(map(tuple, mc))


def test_integer_and_mixed_labels():
    """Integer labels crash the reference's heap on ties (defect D1)."""
    net = wl.random_dag(37, 4, 2, 4, 8, 0)
    relabel = {v: i for i, v in enumerate(sorted(net["sizes"]))}
    factors = [[relabel[v] for v in f] for f in net["factors"]]
    sizes = {relabel[v]: n for v, n in net["sizes"].items()}
    _, mc, _, tree, seps = _compile(factors, sizes)
    assert cons.check_running_intersection(tree, mc + seps)
    _, mc2, _, tree2, seps2 = _compile([[1, "x"], ["x", (2, 3)]], {1: 2, "x": 3, (2, 3): 2})
    assert cons.check_running_intersection(tree2, mc2 + seps2)


def test_no_edges_and_disconnected_graphs():
    """reference tests/test_junctiontree.py:610-612 and defect D13."""
    tri, mc, f2c, tree, seps = _compile([["x"]], {"x": 42})
    assert mc == [["x"]] and tree == [0] and seps == [] and [f2c[0]] == [0]
    tri, mc, f2c, tree, seps = _compile([["x"], ["y"]], {"x": 2, "y": 3})
    assert sorted(map(tuple, mc)) == [("x",), ("y",)]
    assert seps == [[]]                       # empty separator joins the components
    assert cons.check_running_intersection(tree, mc + seps)


def test_unused_sizes_are_ignored_and_scalar_factors_allowed():
    _, mc, f2c, tree, seps = _compile([["a", "b"], []], {"a": 2, "b": 2, "zzz": 7})
    assert mc == [["a", "b"]] and f2c == [0, 0]


def test_user_elimination_order_reaches_grid_treewidth():
    """Row sweep on an n x n grid gives cliques of n + 1 variables."""
    net = wl.ising(6)
    _, mc, _, tree, seps = _compile(net["factors"], net["sizes"], net["order"])
    assert max(len(c) for c in mc) == 7
    assert cons.check_running_intersection(tree, mc + seps)
    with pytest.raises(ValueError):
        cons.find_triangulation(net["factors"], net["sizes"], net["order"][:-1])


def test_config_sizes_match_the_survey_table():
    """Sum n_C / sum n_S of configs 2, 4, 5: same elimination, same sizes."""
    from junctiontree import schedule as sch
    expect = {"dag37": (27, 90686, 20486), "large_state_tree": (4, 3145728, 26624)}
    for net in (wl.dag37(), wl.large_state_tree()):
        _, mc, f2c, tree, seps = _compile(net["factors"], net["sizes"], net.get("order"))
        plan = sch.Plan(tree, mc + seps, net["sizes"], net["factors"], f2c)
        assert (plan.n_cliques, plan.clique_entries, plan.sep_entries) == expect[net["name"]]


def test_tree_is_rooted_at_its_centre():
    factors = [["v%d" % i, "v%d" % (i + 1)] for i in range(20)]       # a chain of 20 cliques
    sizes = {"v%d" % i: 2 for i in range(21)}
    _, mc, _, tree, seps = _compile(factors, sizes)
    _, _, _, depth, _ = cons.tree_edges(tree)
    assert max(depth.values()) == 10


def test_traversal_helpers():
    tree = [0, (5, [1, (7, [3])]), (6, [2])]
    assert list(cons.bf_traverse(tree)) == [0, 5, 6, 1, 2, 7, 3]
    assert list(cons.df_traverse(tree)) == [0, 5, 1, 7, 3, 6, 2]
    assert list(cons.bf_traverse(tree, clique_ix=6)) == [0, 5, 6]          # no RuntimeError (D5)
    # as in the reference, separators are nodes of the traversal too: clique->sep and sep->clique
    assert cons.generate_potential_pairs(tree) == [(0, 5), (0, 6), (5, 1), (6, 2), (1, 7), (7, 3)]
    nodes = [["a", "b"], ["b", "c"], ["a", "d"], ["c", "e"], None, ["b"], ["a"], ["c"]]
    assert cons.get_clique(tree, nodes, "e") == (3, ["c", "e"])
    assert cons.get_clique(tree, nodes, "zz") is None
    assert cons.find_subtree(tree, 3) and not cons.find_subtree(tree, 9)


def test_deep_chain_needs_no_recursion():
    n = 3000                                   # deeper than Python's recursion limit (defect D15)
    factors = [["v%d" % i, "v%d" % (i + 1)] for i in range(n)]
    sizes = {"v%d" % i: 2 for i in range(n + 1)}
    _, mc, _, tree, seps = _compile(factors, sizes)
    order, _, _, depth, _ = cons.tree_edges(tree)
    assert len(order) == n and max(depth.values()) == n // 2


def test_compile_is_independent_of_the_hash_seed():
    """The reference's separator axis order changes with PYTHONHASHSEED."""
    code = (
        "import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import jt_workloads as wl\n"
        "from junctiontree import construction as cons\n"
        "net = wl.random_dag(25, 3, 2, 4, 8, 3)\n"
        "_, mc, f2c = cons.find_triangulation(net['factors'], net['sizes'])\n"
        "tree, seps = cons.construct_junction_tree(mc, net['sizes'])\n"
        "print(json.dumps([mc, f2c, seps, repr(tree)]))\n"
    ) % (os.path.join(ROOT, "junction-tree_b200"), ROOT)
    outs = set()
    for seed in ("0", "1", "12345"):
        env = dict(os.environ, PYTHONHASHSEED=seed)
        outs.add(subprocess.check_output([sys.executable, "-c", code], env=env).decode())
    assert len(outs) == 1
    json.loads(outs.pop())


def test_compile_speed_at_500_variables():
    import time
    net = wl.dag500()
    t0 = time.perf_counter()
    _, mc, _, tree, seps = _compile(net["factors"], net["sizes"])
    assert time.perf_counter() - t0 < 5.0       # the reference needs 2.2 s for 250 variables
    assert len(mc) == 424
