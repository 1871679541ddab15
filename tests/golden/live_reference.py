"""Run the UNMODIFIED reference live, beside the oracle restatement (T2) and brute force (T0), on
cases handed over in a file.

    PYTHONHASHSEED=0 python tests/golden/live_reference.py CASES.json CASES.npz [/root/reference]

The cases (factors, sizes, values and a VALID junction tree built by this repository's host
compile -- the reference's own construction drops maximal cliques) are
written by ``tests/test_oracle.py::test_restatement_equals_the_live_reference_on_its_valid_domain``,
which starts this script as a subprocess: the reference and this repository's package share the
import name ``junctiontree``, so the reference gets a process of its own in which only
``oracle/ref_fixed.py`` (NumPy only) comes from this repository.

Per case: T1 = the reference's ``JunctionTree.propagate`` on the given structure (it may raise:
defects D2 / D7), T0 = joint einsum, T2 = ``oracle.ref_fixed.propagate``.  A case is inside the
reference-valid domain when T1 ran and agrees with T0 to 1e-9; there T2 must equal T1.  Prints
one JSON line.  Build container only; nothing on the GPU box reads the reference.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASES_JSON, CASES_NPZ = sys.argv[1], sys.argv[2]
REF = sys.argv[3] if len(sys.argv) > 3 else "/root/reference"
sys.path.insert(0, ROOT)                 # oracle/ (namespace package, NumPy only)
sys.path.insert(0, REF)                  # `junctiontree` = the reference

import junctiontree as ref_jt                                   # noqa: E402

assert os.path.abspath(ref_jt.__file__).startswith(os.path.abspath(REF)), ref_jt.__file__

from oracle import ref_fixed                                    # noqa: E402


def tuplify(tree):
    return [tree[0]] + [(s, tuplify(t)) for s, t in tree[1:]]


def brute(arrays, var_lists, scopes):
    labels, ops = {}, []
    for a, vs in zip(arrays, var_lists):
        ops += [np.asarray(a, np.float64), [labels.setdefault(v, len(labels)) for v in vs]]
    return [np.einsum(*(ops + [[labels[v] for v in s]])) for s in scopes]


def rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    if got.shape != want.shape:
        return float("inf")
    if not want.size:
        return 0.0
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), np.finfo(np.float64).tiny)))


def main():
    with open(CASES_JSON) as fh:
        cases = json.load(fh)
    with np.load(CASES_NPZ) as npz:
        arrays = {k: npz[k] for k in npz.files}
    ran = valid = 0
    worst_t2_t1 = worst_t2_t0 = 0.0
    raised = {}
    for case in cases:
        values = [arrays["%s/value%d" % (case["name"], f)] for f in range(len(case["factors"]))]
        tree = tuplify(case["tree"])
        truth = brute(values, case["factors"], case["factors"])
        t2, _ = ref_fixed.propagate(tree, case["separators"], case["maxcliques"], case["factor_to_maxclique"],
                                    case["factors"], case["sizes"], values)
        worst_t2_t0 = max(worst_t2_t0, max(rel(a, b) for a, b in zip(t2, truth)))
        graph = ref_jt.FactorGraph(factors=case["factors"], sizes=dict(case["sizes"]))
        clique_graph = ref_jt.CliqueGraph(maxcliques=case["maxcliques"], factor_to_maxclique=case["factor_to_maxclique"],
                                          factor_graph=graph)
        jtree = ref_jt.JunctionTree(tree=tree, separators=case["separators"], clique_tree=clique_graph)
        try:
            t1 = jtree.propagate(values)
        except Exception as exc:
            raised[type(exc).__name__] = raised.get(type(exc).__name__, 0) + 1
            continue
        ran += 1
        if max(rel(a, b) for a, b in zip(t1, truth)) > 1e-9:
            continue                                              # outside the reference-valid domain (D2 / D3)
        valid += 1
        worst_t2_t1 = max(worst_t2_t1, max(rel(a, b) for a, b in zip(t2, t1)))
    print(json.dumps({"cases": len(cases), "reference_ran": ran, "reference_valid": valid, "reference_raised": raised,
                      "max_rel_t2_vs_t1_on_valid": worst_t2_t1, "max_rel_t2_vs_t0": worst_t2_t0,
                      "reference": os.path.abspath(ref_jt.__file__),
                      "hashseed": os.environ.get("PYTHONHASHSEED", "random")}))


if __name__ == "__main__":
    main()
