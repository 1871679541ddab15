"""Shared helpers for the test-suite (golden vectors, network compilation, comparisons)."""

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_JSON = os.path.join(HERE, "golden", "reference_golden.json")
GOLDEN_NPZ = os.path.join(HERE, "golden", "reference_golden.npz")

# tolerances stated by BASELINE.json's north_star
RTOL_F64 = 1e-12
RTOL_F32 = 1e-5


def load_golden():
    """The recorded reference runs: the reference's own fixtures (reference_golden.*) and, since
    round 2, twelve more random nets of 6-10 variables, each also conditioned on its last variable
    (reference_golden_extra.*, ``make_golden.py --extra``).  Returns (cases, arrays by key)."""
    cases, arrays = [], {}
    for stem in ("reference_golden", "reference_golden_extra"):
        with open(os.path.join(HERE, "golden", stem + ".json")) as fh:
            cases += json.load(fh)["cases"]
        with np.load(os.path.join(HERE, "golden", stem + ".npz")) as npz:
            arrays.update({k: npz[k] for k in npz.files})
    return cases, arrays


def tuplify(tree):
    """JSON nested lists -> the reference's tree format [clique, (sep, subtree), ...]."""
    return [tree[0]] + [(s, tuplify(t)) for s, t in tree[1:]]


#: max per-entry relative error of every comparison, by label (dumped by conftest at session end
#: to gpurun_out/parity_report.json when that directory exists)
PARITY_REPORT = {}


def max_rel_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    if not want.size:
        return 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(got - want) / np.abs(want)
    rel = np.where(want == 0, np.where(got == 0, 0.0, np.inf), rel)
    return float(np.max(rel))


def assert_close(got, want, rtol, what="", signed=False):
    """Relative error <= rtol on EVERY entry (north_star: "relative error <= 1e-12 / 1e-5 on every
    clique and separator potential"); a zero must come back as an exact zero.  ``signed=True`` is
    for the tests that feed signed random potentials (the reference's own operator tests do,
    tests/test_computation.py:51-322): sums then cancel, and an entry that is tiny relative to
    its array is compared against the array scale instead."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    if signed:
        scale = np.max(np.abs(want)) if want.size else 0.0
        np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * scale * 1e-3, err_msg=what)
        return
    err = max_rel_err(got, want)
    if what:
        PARITY_REPORT[what] = max(err, PARITY_REPORT.get(what, 0.0))
    assert err <= rtol, "%s: max per-entry relative error %.3g > %g" % (what, err, rtol)


def compile_net(net, with_evidence=True):
    """(tree, separators, maxcliques, factor_to_clique, effective sizes, evidence_vars)."""
    from junctiontree import construction as cons
    _, mc, f2c = cons.find_triangulation(net["factors"], net["sizes"], net.get("order"))
    tree, seps = cons.construct_junction_tree(mc, net["sizes"])
    evars = list(net.get("evidence_vars", [])) if with_evidence else []
    eff = dict(net["sizes"])
    for v in evars:
        eff[v] = 1
    return tree, seps, mc, f2c, eff, evars


SEMIRING_NAMES = ("max_product", "log_sum_exp", "max_sum")


def semiring_inputs(values, semiring):
    """Factor tables in the domain of ``semiring``: log potentials for the log-domain laws
    (zeros become -inf)."""
    if semiring in ("log_sum_exp", "max_sum"):
        with np.errstate(divide="ignore"):
            return [np.log(np.asarray(v, np.float64)) for v in values]
    return [np.asarray(v, np.float64) for v in values]


def assert_close_semiring(got, want, rtol, semiring, what=""):
    """Product-domain laws: relative error as ``assert_close``.  Log-domain laws: an absolute
    error of ``rtol`` on a log value *is* a relative error of ``rtol`` on the potential;
    -inf (probability zero) must match exactly."""
    if semiring in ("sum_product", "max_product"):
        return assert_close(got, want, rtol, what)
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    ninf = np.isneginf(want)
    assert np.array_equal(np.isneginf(got), ninf), "%s: -inf pattern differs" % what
    g, w = got[~ninf], want[~ninf]
    assert np.all(np.abs(g - w) <= rtol * np.maximum(1.0, np.abs(w))), \
        "%s: max abs error %g" % (what, np.max(np.abs(g - w)) if g.size else 0.0)
