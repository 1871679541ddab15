"""Replay the golden vectors through the UNMODIFIED reference and compare with what was recorded.

    PYTHONHASHSEED=0 python tests/golden/replay_reference.py [/root/reference] [--case=NAME ...]

The committed vectors (reference_golden*.json / .npz, written by make_golden.py) hold, per case,
the inputs, the structure the reference built and the reference's outputs.  This script rebuilds
the reference's own objects from the recorded structure, calls the reference's ``propagate`` /
``evaluate`` / ``compute_beliefs`` on the recorded inputs and checks every recorded output -- so
anyone with the reference checkout can confirm that the vectors are the reference's, not ours.
It runs in the build container only (``tests/test_oracle.py::test_golden_vectors_replay_through_
the_unmodified_reference`` starts it as a subprocess when the checkout is present; the GPU box
has no reference and nothing there reads it).  Prints one JSON line: the recorded arrays this
hash seed reproduces (<= 1e-12), those it does not, and the cases in which the reference raised.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_args = [a for a in sys.argv[1:] if not a.startswith("--case=")]
REF = _args[0] if _args else "/root/reference"
ONLY = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--case=")]      # replay these cases only
sys.path.insert(0, REF)

import junctiontree as ref_jt                                   # noqa: E402
from junctiontree import computation as ref_comp                # noqa: E402

assert os.path.abspath(ref_jt.__file__).startswith(os.path.abspath(REF)), ref_jt.__file__


def tuplify(tree):
    return [tree[0]] + [(s, tuplify(t)) for s, t in tree[1:]]


def differ(got, want):
    """Largest relative difference (inf on a shape mismatch); recorded values are float64."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    if got.shape != want.shape:
        return float("inf")
    if not want.size:
        return 0.0
    scale = np.maximum(np.abs(want), np.finfo(np.float64).tiny)
    return float(np.max(np.abs(got - want) / scale))


def replay(case, arrays):
    """[(key of the recorded array, what the reference returns now)] for one case."""
    tree = tuplify(case["tree"])
    if case["kind"] == "operator":
        potentials = [arrays[k] for k in case["potentials"]]
        return list(zip(case["beliefs"], ref_comp.compute_beliefs(tree, potentials, case["variables"])))
    values = [arrays[k] for k in case["values"]]
    graph = ref_jt.FactorGraph(factors=case["factors"], sizes=dict(case["sizes"]))
    clique_graph = ref_jt.CliqueGraph(maxcliques=case["maxcliques"], factor_to_maxclique=case["factor_to_maxclique"],
                                      factor_graph=graph)
    jtree = ref_jt.JunctionTree(tree=tree, separators=case["separators"], clique_tree=clique_graph)
    psi = clique_graph.evaluate(values)
    results = list(zip(case["psi"], psi))
    results += list(zip(case["outputs"], jtree.propagate(values)))
    if "beliefs" in case:
        seps = [np.ones(tuple(case["sizes"][v] for v in s)) for s in case["separators"]]
        node_vars = [list(c) for c in case["maxcliques"]] + [list(s) for s in case["separators"]]
        results += list(zip(case["beliefs"], ref_comp.compute_beliefs(tree, psi + seps, node_vars)))
    return results


def main():
    """The reference's message division depends on the iteration order of Python sets of variable
    labels, i.e. on PYTHONHASHSEED: under an unlucky seed a case raises
    a broadcast error or returns other numbers.  So a run reports, per recorded array, whether
    THIS seed reproduces it; the test takes the union over a few seeds."""
    reproduced, differs, raised, worst, n_cases = [], [], [], 0.0, 0
    for stem in ("reference_golden", "reference_golden_extra"):
        with open(os.path.join(HERE, stem + ".json")) as fh:
            cases = json.load(fh)["cases"]
        with np.load(os.path.join(HERE, stem + ".npz")) as npz:
            arrays = {k: npz[k] for k in npz.files}
        for case in cases:
            if ONLY and case["name"] not in ONLY:
                continue
            n_cases += 1
            recorded = case.get("beliefs", []) + case.get("outputs", []) + case.get("psi", [])
            try:
                results = replay(case, arrays)
            except Exception as exc:                  # a reference defect under this hash seed
                raised.append([stem + ":" + case["name"], "%s: %s" % (type(exc).__name__, str(exc)[:80])])
                continue
            assert sorted(k for k, _ in results) == sorted(recorded)
            for key, got in results:
                err = differ(got, arrays[key])
                if err <= 1e-12:
                    reproduced.append(key)
                    worst = max(worst, err)
                else:
                    differs.append([key, err])
    print(json.dumps({"cases": n_cases, "reproduced": reproduced, "differs": differs, "raised": raised,
                      "max_rel_diff_of_reproduced": worst, "reference": os.path.abspath(ref_jt.__file__),
                      "hashseed": os.environ.get("PYTHONHASHSEED", "random")}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
