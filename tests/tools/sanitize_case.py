"""Small propagation exercising every kernel (TMA ring, uniform warp, LDG, split-r, init), meant
to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tests/tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tests/tools/sanitize_case.py
"""

import os
import sys

import numpy as np

os.environ.setdefault("JT_BETA_MIN_MB", "0")      # run jt_beta_kernel on these small launches too

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "junction-tree_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import jt_workloads as wl  # noqa: E402
import junctiontree as jt  # noqa: E402
from oracle import ref_fixed  # noqa: E402


def late_round2():
    """The kernels changed last in round 2: the clique initialisation on narrow batches (blocks of
    1-4 rows, gathers-only row loop per factor count, float64 and float32 vectors) and
    jt_dense_kernel's per-m-tile-count MMA warps (a net that reaches MT = 1, 3 and 4 with several
    i-tiles per group)."""
    net = wl.random_dag(14, 3, 2, 3, 8, 2)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    ct = tree.clique_tree
    for B, dtype in ((70, np.float64), (200, np.float32), (300, np.float64)):
        ev = wl.draw_evidence(net, B)
        vals = [np.asarray(v, dtype) for v in net["values"]]
        outs, nodes = tree.propagate_batch(vals, net["evidence_vars"], ev, nodes=True, uniform=False)
        want_f, want_n = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                                   net["factors"], net["sizes"], net["values"], net["evidence_vars"],
                                                   ev[:2], n=2)
        rtol = 1e-12 if dtype == np.float64 else 1e-5
        for g, w in zip(list(outs) + list(nodes), list(want_f) + list(want_n)):
            np.testing.assert_allclose(g[:2], w, rtol=rtol)
        print("ok init rows", B, np.dtype(dtype).name, "per-instance")
    net2 = wl.random_dag(24, 4, 3, 6, 6, 11)
    tree2 = jt.create_junction_tree(net2["factors"], net2["sizes"])
    ct2 = tree2.clique_tree
    for B in (256, 300):
        ev2 = wl.draw_evidence(net2, B)
        outs, nodes = tree2.propagate_batch(net2["values"], net2["evidence_vars"], ev2, nodes=True, dense=True)
        want_f, want_n = ref_fixed.propagate_batch(tree2.tree, tree2.separators, ct2.maxcliques, ct2.factor_to_maxclique,
                                                   net2["factors"], net2["sizes"], net2["values"], net2["evidence_vars"],
                                                   ev2[:2], n=2)
        for g, w in zip(list(outs) + list(nodes), list(want_f) + list(want_n)):
            np.testing.assert_allclose(g[:2], w, rtol=1e-12)
        print("ok dense m-tile variants", net2["name"], B)


def main():
    if "--quick" in sys.argv:          # only the kernels changed last (fits a short GPU slot under memcheck)
        late_round2()
        print("done")
        return
    late_round2()
    net = wl.random_dag(14, 3, 2, 3, 8, 2)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    ct = tree.clique_tree
    for B, dtype, uniform in ((300, np.float64, True), (1030, np.float64, True), (520, np.float32, False), (5, np.float64, True)):
        ev = wl.draw_evidence(net, B)
        vals = [np.asarray(v, dtype) for v in net["values"]]
        outs, nodes = tree.propagate_batch(vals, net["evidence_vars"], ev, nodes=True, uniform=uniform)
        want_f, want_n = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                                   net["factors"], net["sizes"], net["values"], net["evidence_vars"],
                                                   ev[:2], n=2)
        rtol = 1e-12 if dtype == np.float64 else 1e-5
        for g, w in zip(list(outs) + list(nodes), list(want_f) + list(want_n)):
            np.testing.assert_allclose(g[:2], w, rtol=rtol)
        print("ok", B, np.dtype(dtype).name, "uniform" if uniform else "per-instance")
    tree.propagate(net["values"])
    # other semirings (log domain), soft evidence (likelihood region, init_rows kernel at B >= 512),
    # direct marginals without clique beliefs, output stage
    from junctiontree import semirings as sr
    B = 520
    ev = wl.draw_evidence(net, B)
    rng = np.random.default_rng(0)
    free = [v for v in sorted(net["sizes"]) if v not in net["evidence_vars"]]
    lik = {v: rng.random((B, net["sizes"][v])) + 0.1 for v in free[:2]}
    for law, name, vals, lk in ((sr.max_product, "max_product", net["values"], lik),
                                (sr.log_sum_exp, "log_sum_exp", [np.log(v) for v in net["values"]],
                                 {v: np.log(x) for v, x in lik.items()})):
        outs, nodes = tree.propagate_batch(vals, net["evidence_vars"], ev, nodes=True, dl=law, likelihoods=lk)
        b = 3
        fx, f2cx, vx = ref_fixed.with_likelihood_factors(net["factors"], ct.factor_to_maxclique, ct.maxcliques,
                                                         vals, lk, b)
        want_f, want_n = ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, f2cx, fx, net["sizes"],
                                                   vx, net["evidence_vars"], ev[b:b + 1], n=1, semiring=name)
        for g, w in zip(list(outs) + list(nodes), list(want_f[:len(outs)]) + list(want_n)):
            np.testing.assert_allclose(g[b], w[0], rtol=1e-11, atol=1e-11)
        print("ok", name, "with soft evidence")
    marg, log_z = tree.marginals_batch(net["values"], None, net["evidence_vars"], ev, likelihoods=lik)
    assert np.all(np.isfinite(log_z))
    # round 2: the accelerated uniform-mode kernels (jt_dense_kernel incl. multi-unit stages and the
    # float32-storage form, jt_dense_prep_kernel, jt_scalar_kernel, jt_beta_kernel on its side
    # stream, level fork / join) against the projection kernels and the oracle
    for make in (lambda: wl.large_state_tree((8, 12, 16, 8, 12, 16)), lambda: wl.random_dag(60, 3, 2, 5, 8, 3)):
        net2 = make()
        tree2 = jt.create_junction_tree(net2["factors"], net2["sizes"])
        ct2 = tree2.clique_tree
        for B, dtype in ((256, np.float64), (1000, np.float64), (520, np.float32)):
            ev2 = wl.draw_evidence(net2, B)
            vals = [np.asarray(v, dtype) for v in net2["values"]]
            outs, nodes = tree2.propagate_batch(vals, net2["evidence_vars"], ev2, nodes=True, dense=True)
            outs_p, nodes_p = tree2.propagate_batch(vals, net2["evidence_vars"], ev2, nodes=True, dense=False)
            rtol = 1e-12 if dtype == np.float64 else 1e-5
            for g, w in zip(list(outs) + list(nodes), list(outs_p) + list(nodes_p)):
                np.testing.assert_allclose(g, w, rtol=rtol)
            want_f, want_n = ref_fixed.propagate_batch(tree2.tree, tree2.separators, ct2.maxcliques,
                                                       ct2.factor_to_maxclique, net2["factors"], net2["sizes"],
                                                       [np.asarray(v, np.float64) for v in vals],
                                                       net2["evidence_vars"], ev2[:2], n=2)
            for g, w in zip(list(outs) + list(nodes), list(want_f) + list(want_n)):
                np.testing.assert_allclose(g[:2], w, rtol=rtol)
            tree2.propagate_batch(vals, net2["evidence_vars"], ev2, dense=True)        # outputs only
            print("ok accelerated", net2["name"], B, np.dtype(dtype).name)
    print("done")


if __name__ == "__main__":
    main()
