"""The oracles are pinned here (CPU only): the NumPy restatement against the unmodified
reference's recorded outputs (tests/golden, produced by tests/golden/make_golden.py), against
brute force, and the schedule interpreter against both."""

import numpy as np
import pytest

import jt_workloads as wl
from helpers import RTOL_F64, assert_close, compile_net, load_golden, tuplify
from oracle import brute, plan_interp, ref_fixed


def test_golden_file_covers_the_reference_fixtures():
    cases, arrays = load_golden()
    names = {c["name"] for c in cases}
    for expected in ("sprinkler", "sprinkler_wet", "sprinkler_wet_rain", "huang_darwiche", "wisconsin",
                     "scalar_node", "child_all_shared", "two_children_3d", "grandchild_shared"):
        assert expected in names
    # the reference's own known-answer numbers (tests/test_junctiontree.py:245-292, 422-525)
    hd = next(c for c in cases if c["name"] == "huang_darwiche")
    np.testing.assert_allclose(arrays[hd["outputs"][0]], [0.5, 0.5])
    np.testing.assert_allclose(arrays[hd["outputs"][3]].sum(axis=0), [0.32, 0.68])
    np.testing.assert_allclose(arrays[hd["outputs"][4]].sum(axis=0), [0.535, 0.465])
    wi = next(c for c in cases if c["name"] == "wisconsin")
    np.testing.assert_allclose(arrays[wi["outputs"][2]].sum(axis=1), [0.75, 0.25])
    np.testing.assert_allclose(arrays[wi["outputs"][3]].sum(axis=0), [0.546, 0.454])
    # the reference defect D3 is visible in the recorded vectors
    wet = next(c for c in cases if c["name"] == "sprinkler_wet")
    assert wet["outputs_valid"] == [True, True, True, False]


def test_ref_fixed_compute_beliefs_matches_reference_on_its_trees():
    cases, arrays = load_golden()
    checked = 0
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
        got = ref_fixed.compute_beliefs(tuplify(case["tree"]), pots, node_vars)
        for k, key in enumerate(case["beliefs"]):
            if case["beliefs_valid"][k]:
                assert_close(got[k], arrays[key], RTOL_F64, "%s node %d" % (case["name"], k), signed=case["kind"] == "operator")
                checked += 1
    assert checked >= 320


def test_ref_fixed_propagate_matches_reference_outputs():
    """Factor outputs do not depend on the tree, so our own host compile can be used."""
    cases, arrays = load_golden()
    checked = 0
    for case in cases:
        if case["kind"] != "end_to_end":
            continue
        values = [arrays[k] for k in case["values"]]
        sizes = dict(case["sizes"])
        sizes.update({v: 1 for v in case["slices"]})
        net = {"factors": case["factors"], "sizes": sizes}
        tree, seps, mc, f2c, eff, _ = compile_net(net)
        outs, _ = ref_fixed.propagate(tree, seps, mc, f2c, case["factors"], eff, values)
        for f, key in enumerate(case["outputs"]):
            if case["outputs_valid"][f]:
                assert_close(outs[f], arrays[key], RTOL_F64, "%s factor %d" % (case["name"], f))
                checked += 1
    assert checked >= 250


NETS = [wl.sprinkler(), wl.huang_darwiche(), wl.wisconsin(), wl.random_dag(12, 3, 2, 3, 8, 5),
        wl.random_dag(10, 4, 2, 4, 10, 21), wl.ising(4), wl.large_state_tree((4, 6, 8, 4, 6, 8))]


@pytest.mark.parametrize("net", NETS, ids=lambda n: n["name"])
def test_ref_fixed_matches_brute_force(net):
    tree, seps, mc, f2c, eff, evars = compile_net(net)
    B = 3
    ev = wl.draw_evidence(net, B) if evars else None
    outs, ys = ref_fixed.propagate_batch(tree, seps, mc, f2c, net["factors"], net["sizes"], net["values"],
                                         evars, ev, n=B)
    nodes = mc + seps
    for b in range(B):
        evd = {v: int(ev[b][i]) for i, v in enumerate(evars)} if evars else None
        truth = brute.factor_graph_marginals(net["factors"], net["values"], nodes + net["factors"], evd)
        for k in range(len(nodes)):
            assert_close(ys[k][b], truth[k], 1e-11, "node %d" % k)
        for f in range(len(net["factors"])):
            assert_close(outs[f][b], truth[len(nodes) + f], 1e-11, "factor %d" % f)


@pytest.mark.parametrize("uniform", [False, True], ids=["per_instance_psi", "uniform_psi"])
@pytest.mark.parametrize("net", NETS + [wl.dag37()], ids=lambda n: n["name"])
def test_schedule_interpreter_matches_ref_fixed(net, uniform):
    """The compiled plan (index tables, task wiring, launch order), interpreted in NumPy,
    reproduces the oracle -- this is what the CUDA kernels execute."""
    from junctiontree import schedule as sch
    tree, seps, mc, f2c, eff, evars = compile_net(net)
    plan = sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"])
    B = 2
    ev = wl.draw_evidence(net, B) if evars else None
    work, fout = plan_interp.run(plan, B, factor_in=plan_interp.flatten_factors(plan, net["values"]), evidence=ev,
                                 uniform=uniform)
    outs, ys = ref_fixed.propagate_batch(tree, seps, mc, f2c, net["factors"], net["sizes"], net["values"],
                                         evars, ev, n=B)
    assert plan.uni_entries > 0
    for k in range(len(mc) + len(seps)):
        assert_close(plan_interp.node_array(plan, work, k, B), ys[k], 1e-13, "node %d" % k)
    for f in range(len(net["factors"])):
        assert_close(plan_interp.factor_array(plan, fout, f, B), outs[f], 1e-13, "factor %d" % f)


def test_split_tables_equal_direct_index_arithmetic():
    """hi/lo table pairs reproduce sum_v digit_v * stride_v exactly, for spaces larger than one
    table (bit-exact index maps)."""
    from junctiontree import schedule as sch
    rng = np.random.default_rng(0)
    for _ in range(50):
        nv = int(rng.integers(1, 7))
        variables = list(range(nv))
        sizes = {v: int(rng.integers(1, 9)) for v in variables}
        space = sch._Space(variables, sizes)
        target = [v for v in variables if rng.random() < 0.6]
        rng.shuffle(target)
        strides = dict(zip(target, sch._row_major_strides([sizes[v] for v in target])))
        hi, lo = space.tables(strides)
        assert len(hi) == space.n_hi and len(lo) == space.n_lo and space.n_hi * space.n_lo == space.n
        x = np.arange(space.n)
        got = hi[x // space.n_lo] + lo[x % space.n_lo]
        digits = np.unravel_index(x, space.shape) if nv else ()
        want = sum((digits[i] * strides.get(v, 0) for i, v in enumerate(variables)), np.zeros(space.n, np.int64))
        assert np.array_equal(got, want)


def test_invariants_on_a_large_tree():
    """Size-independent properties on config 2's shape: all nodes share one Z and neighbours
    agree on the separator marginal."""
    net = wl.dag37()
    tree, seps, mc, f2c, eff, evars = compile_net(net, with_evidence=False)
    _, ys = ref_fixed.propagate(tree, seps, mc, f2c, net["factors"], net["sizes"], net["values"])
    Z = ys[0].sum()
    np.testing.assert_allclose(Z, 1.0, rtol=1e-12)       # CPTs of a Bayesian network
    for y in ys:
        np.testing.assert_allclose(y.sum(), Z, rtol=1e-12)


@pytest.mark.parametrize("uniform", [False, True], ids=["per_instance_psi", "uniform_psi"])
@pytest.mark.parametrize("net", NETS, ids=lambda n: n["name"])
def test_direct_marginals_without_clique_beliefs(net, uniform):
    """JT_NO_BELIEFS schedule: message-sending distribute tasks only, outputs computed straight
    from psi_C and the incoming messages -- same factor outputs, no clique belief written."""
    from junctiontree import schedule as sch
    tree, seps, mc, f2c, eff, evars = compile_net(net)
    plan = sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"])
    B = 2
    ev = wl.draw_evidence(net, B) if evars else None
    work, fout = plan_interp.run(plan, B, factor_in=plan_interp.flatten_factors(plan, net["values"]), evidence=ev,
                                 uniform=uniform, beliefs=False)
    outs, _ = ref_fixed.propagate_batch(tree, seps, mc, f2c, net["factors"], net["sizes"], net["values"],
                                        evars, ev, n=B)
    for f in range(len(net["factors"])):
        assert_close(plan_interp.factor_array(plan, fout, f, B), outs[f], 1e-13, "factor %d" % f)


def test_soft_evidence_is_one_more_single_variable_factor():
    """Soft (likelihood) evidence: the plan multiplies a per-instance vector from the
    workspace's likelihood region into a clique; the oracle states it in the reference's own
    terms -- an extra factor [v] per instance -- and one-hot likelihoods reproduce hard evidence
    (the reference's own test of slicing vs one-hot, tests/test_computation.py:411-459)."""
    from junctiontree import schedule as sch
    net = wl.random_dag(12, 3, 2, 3, 8, 5)
    tree, seps, mc, f2c, eff, evars = compile_net(net)
    B = 3
    rng = np.random.default_rng(0)
    free = [v for v in sorted(net["sizes"]) if v not in evars]
    soft = [free[1], free[4], free[6]]
    lik = {v: rng.random((B, net["sizes"][v])) + 0.1 for v in soft}
    ev = wl.draw_evidence(net, B)
    for uniform in (False, True):
        plan = sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"], likelihood_vars=soft)
        work = np.zeros((plan.work_entries, B))
        plan_interp.load_likelihoods(plan, work, lik)
        work, fout = plan_interp.run(plan, B, work=work, factor_in=plan_interp.flatten_factors(plan, net["values"]),
                                     evidence=ev, uniform=uniform)
        for b in range(B):
            fx, f2cx, vx = ref_fixed.with_likelihood_factors(net["factors"], f2c, mc, net["values"], lik, b)
            outs, ys = ref_fixed.propagate_batch(tree, seps, mc, f2cx, fx, net["sizes"], vx, evars, ev[b:b + 1], n=1)
            for k in range(len(mc) + len(seps)):
                assert_close(plan_interp.node_array(plan, work, k, B)[b], ys[k][0], 1e-13, "node %d" % k)
            for f in range(len(net["factors"])):
                assert_close(plan_interp.factor_array(plan, fout, f, B)[b], outs[f][0], 1e-13, "factor %d" % f)
    # one-hot likelihoods on the observed variables == slicing them
    plan = sch.Plan(tree, mc + seps, net["sizes"], net["factors"], f2c, likelihood_vars=evars)
    onehot = {v: np.eye(net["sizes"][v])[ev[:, i]] for i, v in enumerate(evars)}
    work = plan_interp.load_likelihoods(plan, np.zeros((plan.work_entries, B)), onehot)
    work, fout = plan_interp.run(plan, B, work=work, factor_in=plan_interp.flatten_factors(plan, net["values"]))
    outs, _ = ref_fixed.propagate_batch(tree, seps, mc, f2c, net["factors"], net["sizes"], net["values"], evars, ev, n=B)
    for f, fv in enumerate(net["factors"]):
        full = plan_interp.factor_array(plan, fout, f, B)
        for b in range(B):
            ix = tuple(slice(int(ev[b, evars.index(v)]), int(ev[b, evars.index(v)]) + 1) if v in evars else slice(None)
                       for v in fv)
            assert_close(full[b][ix], outs[f][b], 1e-13, "factor %d instance %d" % (f, b))


def test_golden_vectors_replay_through_the_unmodified_reference():
    """The pin itself, re-checked live: every committed golden array (738 in 43 cases) is what the
    UNMODIFIED reference returns on the recorded inputs and structures
    (tests/golden/replay_reference.py, run as a subprocess with the reference checkout first on
    its path).  The reference's message division depends on Python's set iteration order
    (SURVEY.md section 9, D2), i.e. on PYTHONHASHSEED: an unlucky seed makes a case raise or return
    other numbers, so the vectors of the first file (recorded under seed 0) must all reproduce
    under seed 0 and every other array under at least one of a fixed handful of seeds.  Build
    container only: the reference does not travel to the GPU box."""
    import json
    import os
    import subprocess
    import sys
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "junctiontree")):
        pytest.skip("no reference checkout at %s" % ref)
    here = os.path.dirname(os.path.abspath(__file__))
    script = os.path.join(here, "golden", "replay_reference.py")
    wanted, first_file = set(), set()
    for stem in ("reference_golden", "reference_golden_extra"):
        with open(os.path.join(here, "golden", stem + ".json")) as fh:
            for case in json.load(fh)["cases"]:
                keys = set(case.get("beliefs", []) + case.get("outputs", []) + case.get("psi", []))
                wanted |= keys
                if stem == "reference_golden":
                    first_file |= keys
    assert len(wanted) == 738 and len(first_file) == 186
    covered = set()
    for seed in (0, 2, 5, 6, 8, 9, 20, 98):
        env = dict(os.environ, PYTHONHASHSEED=str(seed))
        env.pop("PYTHONPATH", None)
        res = subprocess.run([sys.executable, "-P", script, ref], capture_output=True, text=True, env=env,
                             cwd="/tmp", timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        report = json.loads(res.stdout.strip().splitlines()[-1])
        assert report["reference"].startswith(ref) and report["cases"] == 43
        assert report["max_rel_diff_of_reproduced"] <= 1e-12
        covered |= set(report["reproduced"])
        if seed == 0:
            assert first_file <= covered, sorted(first_file - covered)
    assert covered == wanted, sorted(wanted - covered)


def test_restatement_equals_the_live_reference_on_its_valid_domain(tmp_path):
    """T2 (oracle/ref_fixed.py) against the UNMODIFIED reference run live on 120 random networks of
    6-11 variables over VALID junction trees from this repository's host compile
    (tests/golden/live_reference.py, a subprocess with the reference checkout first on its path):
    wherever the reference runs and agrees with brute force -- its valid domain, SURVEY.md 8c --
    the restatement returns the same numbers (<= 1e-12), and it agrees with brute force on every
    case, also where the reference raises or is wrong (defects D2 / D3 / D7).  Build container
    only."""
    import json
    import os
    import subprocess
    import sys
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "junctiontree")):
        pytest.skip("no reference checkout at %s" % ref)
    from junctiontree import construction as cons
    cases, arrays = [], {}
    for seed in range(120):
        n = 6 + seed % 6
        net = wl.random_dag(n, 2 + seed % 2, 2, 3 + seed % 2, n, 7000 + seed)
        _, mc, f2c = cons.find_triangulation(net["factors"], net["sizes"])
        tree, seps = cons.construct_junction_tree(mc, net["sizes"])
        name = "live%d" % seed

        def listify(t):
            return [int(t[0])] + [[int(s), listify(sub)] for s, sub in t[1:]]

        cases.append({"name": name, "factors": net["factors"], "sizes": {k: int(v) for k, v in net["sizes"].items()},
                      "tree": listify(tree), "maxcliques": [list(c) for c in mc], "separators": [list(s) for s in seps],
                      "factor_to_maxclique": [int(c) for c in f2c]})
        for f, v in enumerate(net["values"]):
            arrays["%s/value%d" % (name, f)] = np.asarray(v, np.float64)
    cases_json, cases_npz = str(tmp_path / "cases.json"), str(tmp_path / "cases.npz")
    with open(cases_json, "w") as fh:
        json.dump(cases, fh)
    np.savez(cases_npz, **arrays)
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, PYTHONHASHSEED="0")
    env.pop("PYTHONPATH", None)
    res = subprocess.run([sys.executable, "-P", os.path.join(here, "golden", "live_reference.py"), cases_json, cases_npz,
                          ref], capture_output=True, text=True, env=env, cwd="/tmp", timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    report = json.loads(res.stdout.strip().splitlines()[-1])
    assert report["reference"].startswith(ref) and report["cases"] == 120
    assert report["max_rel_t2_vs_t0"] <= 1e-11                      # the restatement is right everywhere
    assert report["reference_valid"] >= 30, report                  # enough of the reference's valid domain was hit
    assert report["max_rel_t2_vs_t1_on_valid"] <= 1e-12, report     # and there it is the reference's numbers
