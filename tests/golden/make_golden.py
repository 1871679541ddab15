"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py          # writes tests/golden/reference_golden.npz + .json
    python tests/golden/make_golden.py --extra  # writes tests/golden/reference_golden_extra.npz + .json:
                                                # more random nets of 6-10 variables, each also conditioned
                                                # on its last variable (round 2; the first file is unchanged)

It imports ``junctiontree`` from ``/root/reference`` (oracle tier T1) and records,
for every case, the inputs, the structure the reference built (tree, maxcliques, separators,
factor_to_maxclique -- separator axis order is hash-seed dependent there, so it is stored, never
recomputed) and the reference's outputs.  Each output is also compared with brute force (T0); the
flag ``valid`` says whether the reference agreed with brute force to 1e-9, i.e. whether the case
lies inside the reference-valid domain.  Nothing here is imported by the product.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "junction-tree_b200"))   # jt_workloads only
sys.path.insert(0, "/root/reference")                          # `junctiontree` = the reference

import junctiontree as ref_jt                                   # noqa: E402
from junctiontree import computation as ref_comp                # noqa: E402

assert ref_jt.__file__.startswith("/root/reference"), ref_jt.__file__

import jt_workloads as wl                                       # noqa: E402


def brute(arrays, var_lists, scopes):
    labels = {}
    ops = []
    for a, vs in zip(arrays, var_lists):
        ops += [np.asarray(a, np.float64), [labels.setdefault(v, len(labels)) for v in vs]]
    return [np.einsum(*(ops + [[labels[v] for v in s]])) for s in scopes]


def listify(tree):
    return [tree[0]] + [[int(s), listify(t)] for s, t in tree[1:]]


def close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.allclose(a, b, rtol=1e-9, atol=1e-300))


arrays = {}
meta = {"cases": []}


def put(name, arr):
    arrays[name] = np.asarray(arr, np.float64)
    return name


def end_to_end(name, net, slices=None):
    """create_junction_tree + propagate + compute_beliefs through the reference."""
    factors, sizes, values = net["factors"], dict(net["sizes"]), [np.array(v) for v in net["values"]]
    tree = ref_jt.create_junction_tree(factors, sizes)
    if slices:
        # conditioning as the reference tests do it: mutate sizes, slice arrays
        # (tests/test_junctiontree.py:393-411)
        for var, state in slices.items():
            tree.clique_tree.factor_graph.sizes[var] = 1
            for f, fv in enumerate(factors):
                if var in fv:
                    ix = [slice(None)] * len(fv)
                    ix[fv.index(var)] = slice(state, state + 1)
                    values[f] = values[f][tuple(ix)]
    outs = tree.propagate(values)
    ct = tree.clique_tree
    psi = ct.evaluate(values)
    eff_sizes = tree.clique_tree.factor_graph.sizes
    seps = [np.ones(tuple(eff_sizes[v] for v in s)) for s in tree.separators]
    node_vars = [list(c) for c in ct.maxcliques] + [list(s) for s in tree.separators]
    try:
        beliefs = ref_comp.compute_beliefs(tree.tree, psi + seps, node_vars)
    except Exception as exc:     # reference defect D7 on some conditioned inputs
        beliefs = None
        print("  compute_beliefs failed in the reference:", type(exc).__name__, exc)
    truth = brute(values, factors, factors)
    case = {
        "name": name, "kind": "end_to_end", "factors": factors, "sizes": {k: int(v) for k, v in sizes.items()},
        "slices": slices or {}, "tree": listify(tree.tree), "maxcliques": [list(c) for c in ct.maxcliques],
        "separators": [list(s) for s in tree.separators],
        "factor_to_maxclique": [int(ct.factor_to_maxclique[i]) for i in range(len(factors))],
        "values": [put("%s/value%d" % (name, f), v) for f, v in enumerate(values)],
        "outputs": [put("%s/out%d" % (name, f), o) for f, o in enumerate(outs)],
        "outputs_valid": [close(o, t) for o, t in zip(outs, truth)],
        "psi": [put("%s/psi%d" % (name, c), p) for c, p in enumerate(psi)],
    }
    if beliefs is not None:
        full_psi_truth = brute(values, factors, node_vars)
        case["beliefs"] = [put("%s/belief%d" % (name, k), b) for k, b in enumerate(beliefs)]
        case["beliefs_valid"] = [close(b, t) for b, t in zip(beliefs, full_psi_truth)]
    meta["cases"].append(case)
    print(name, "outputs valid:", case["outputs_valid"], "beliefs valid:", case.get("beliefs_valid"))


def operator_case(name, tree, potentials, variables):
    """compute_beliefs on a hand-built tree (shapes of tests/test_computation.py:51-322)."""
    beliefs = ref_comp.compute_beliefs(tree, potentials, variables)
    ids = []

    def walk(t):
        ids.append(t[0])
        for s, sub in t[1:]:
            ids.append(s)
            walk(sub)
    walk(tree)
    truth = brute([potentials[i] for i in ids], [variables[i] for i in ids], variables)
    case = {
        "name": name, "kind": "operator", "tree": listify(tree), "variables": variables,
        "potentials": [put("%s/pot%d" % (name, k), p) for k, p in enumerate(potentials)],
        "beliefs": [put("%s/belief%d" % (name, k), b) for k, b in enumerate(beliefs)],
        "beliefs_valid": [close(b, t) for b, t in zip(beliefs, truth)],
    }
    meta["cases"].append(case)
    print(name, "beliefs valid:", case["beliefs_valid"])


EXTRA = "--extra" in sys.argv
rng = np.random.default_rng(20261017)
R = lambda *shape: rng.standard_normal(shape)     # signed values, like the reference tests


def write(stem):
    np.savez_compressed(os.path.join(HERE, stem + ".npz"), **arrays)
    meta["numpy"] = np.__version__
    meta["hashseed"] = os.environ.get("PYTHONHASHSEED", "random")
    with open(os.path.join(HERE, stem + ".json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("wrote %d arrays, %d cases" % (len(arrays), len(meta["cases"])))


if EXTRA:
    # random nets of 6-10 variables, kept only when the unmodified reference runs and agrees with
    # brute force on every output; each is recorded unconditioned and conditioned on its last
    # variable (state 1) the way the reference tests condition (sizes mutated, arrays sliced) --
    # the conditioned outputs carry per-factor validity flags (reference defects D3 / D7)
    for n, max_par, smax, first_seed in ((6, 2, 3, 300), (8, 3, 3, 400), (9, 2, 4, 500), (10, 3, 3, 600)):
        kept = 0
        for seed in range(first_seed, first_seed + 60):
            net = wl.random_dag(n, max_par, 2, smax, n, seed)
            try:
                tree = ref_jt.create_junction_tree(net["factors"], dict(net["sizes"]))
                outs = tree.propagate(net["values"])
            except Exception:
                continue
            truth = brute(net["values"], net["factors"], net["factors"])
            if not all(close(o, t) for o, t in zip(outs, truth)):
                continue
            name = "dag%d_seed%d" % (n, seed)
            end_to_end(name, net)
            last = "v%03d" % (n - 1)
            try:
                end_to_end(name + "_cond", net, {last: 1})
            except Exception as exc:
                print("  conditioned case failed in the reference:", type(exc).__name__, exc)
            kept += 1
            if kept == 3:
                break
    write("reference_golden_extra")
    sys.exit(0)

end_to_end("sprinkler", wl.sprinkler())
end_to_end("sprinkler_wet", wl.sprinkler(), {"wet_grass": 1})
end_to_end("sprinkler_wet_rain", wl.sprinkler(), {"wet_grass": 1, "rain": 1})
end_to_end("huang_darwiche", wl.huang_darwiche())
end_to_end("wisconsin", wl.wisconsin())

one_child = [0, (2, [1])]
chain3 = [0, (3, [1, (4, [2])])]
two_kids = [0, (3, [1]), (4, [2])]
operator_case("scalar_node", [0], [R()], [[]])
operator_case("matrix_node", [0], [R(2, 3)], [[3, 5]])
operator_case("child_all_shared", one_child, [R(2, 3), R(3, 2), np.ones((3, 2))], [[3, 5], [5, 3], [5, 3]])
operator_case("child_one_common", one_child, [R(2, 3), R(3, 4), np.ones((3,))], [[3, 5], [5, 9], [5]])
operator_case("child_no_common", one_child, [R(2), R(3), np.ones(())], [[3], [9], []])
operator_case("grandchild_not_shared", chain3, [R(2, 3), R(3, 4), R(4, 5), np.ones((3,)), np.ones((4,))],
              [[3, 5], [5, 9], [9, 1], [5], [9]])
operator_case("grandchild_shared", chain3, [R(2, 3), R(3, 4), R(6, 3), np.ones((3,)), np.ones((3,))],
              [[3, 5], [5, 9], [1, 5], [5], [5]])
operator_case("two_children_not_shared", two_kids, [R(2, 3), R(3, 4), R(2, 5), np.ones((3,)), np.ones((2,))],
              [[3, 5], [5, 9], [3, 1], [5], [3]])
operator_case("two_children_shared", two_kids, [R(2, 3), R(3, 4), R(3), np.ones((3,)), np.ones((3,))],
              [[3, 5], [5, 9], [5], [5], [5]])
operator_case("two_children_3d", two_kids, [R(2, 3, 4), R(3, 4, 5), R(3, 6), np.ones((3, 4)), np.ones((3,))],
              [[3, 5, 7], [5, 7, 9], [5, 1], [5, 7], [5]])

# small random nets, kept only when the unmodified reference runs and agrees with brute force
kept = 0
for seed in range(40):
    net = wl.random_dag(7, 2, 2, 3, 7, 100 + seed)
    try:
        tree = ref_jt.create_junction_tree(net["factors"], dict(net["sizes"]))
        outs = tree.propagate(net["values"])
    except Exception:
        continue
    truth = brute(net["values"], net["factors"], net["factors"])
    if all(close(o, t) for o, t in zip(outs, truth)):
        end_to_end("dag7_seed%d" % (100 + seed), net)
        kept += 1
    if kept == 4:
        break

write("reference_golden")
