"""Other distributive laws, CPU side: the oracle's semiring restatements are
pinned (brute force over the joint; exact identities to the reference-pinned sum-product), the
compiled schedule is semiring-agnostic (NumPy interpreter), and the operator surface exists."""

import os

import numpy as np
import pytest

import jt_workloads as wl
from helpers import SEMIRING_NAMES, assert_close, assert_close_semiring, compile_net, semiring_inputs
from oracle import brute, plan_interp, ref_fixed

NETS = [wl.sprinkler(), wl.huang_darwiche(), wl.random_dag(10, 3, 2, 3, 8, 5), wl.ising(3)]


@pytest.mark.parametrize("semiring", SEMIRING_NAMES)
@pytest.mark.parametrize("net", NETS, ids=lambda n: n["name"])
def test_oracle_semiring_matches_brute_force(net, semiring):
    """ref_fixed's collect/distribute in another semiring equals the semiring marginals of the
    joint (every node and every factor scope)."""
    tree, seps, mc, f2c, eff, evars = compile_net(net, with_evidence=False)
    vals = semiring_inputs(net["values"], semiring)
    outs, ys = ref_fixed.propagate(tree, seps, mc, f2c, net["factors"], net["sizes"], vals, semiring)
    scopes = list(mc) + [list(s) for s in seps] + net["factors"]
    truth = brute.factor_graph_marginals(net["factors"], vals, scopes, semiring=semiring)
    for k, (g, w) in enumerate(zip(list(ys) + list(outs), truth)):
        assert_close_semiring(g, w, 1e-12, semiring, "scope %d" % k)


@pytest.mark.parametrize("net", NETS, ids=lambda n: n["name"])
def test_log_domain_identities_pin_the_semirings_to_sum_product(net):
    """log_sum_exp(log x) = log(sum_product(x)) and max_sum(log x) = log(max_product(x)): the
    log-domain laws are pinned to the reference-pinned sum-product oracle."""
    tree, seps, mc, f2c, eff, evars = compile_net(net, with_evidence=False)
    args = (tree, seps, mc, f2c, net["factors"], net["sizes"])
    for prod, logd in (("sum_product", "log_sum_exp"), ("max_product", "max_sum")):
        outs_p, ys_p = ref_fixed.propagate(*args, semiring_inputs(net["values"], prod), prod)
        outs_l, ys_l = ref_fixed.propagate(*args, semiring_inputs(net["values"], logd), logd)
        with np.errstate(divide="ignore"):
            for k, (p, l) in enumerate(zip(list(ys_p) + list(outs_p), list(ys_l) + list(outs_l))):
                assert_close_semiring(l, np.log(p), 1e-12, logd, "%s scope %d" % (logd, k))


def test_max_marginals_give_the_map_state():
    """argmax of the single-variable max-marginals is the MAP assignment (unique maximiser)."""
    net = wl.random_dag(9, 3, 2, 3, 8, 7)
    labels = sorted(net["sizes"])
    joint = brute.joint_marginals(net["values"], net["factors"], [labels])[0]
    best = np.unravel_index(np.argmax(joint), joint.shape)
    mm = brute.factor_graph_marginals(net["factors"], net["values"], [[v] for v in labels], semiring="max_product")
    assert tuple(int(np.argmax(m)) for m in mm) == tuple(int(i) for i in best)
    for m in mm:
        np.testing.assert_allclose(m.max(), joint.max(), rtol=1e-14)


@pytest.mark.parametrize("semiring", SEMIRING_NAMES)
@pytest.mark.parametrize("uniform", [False, True], ids=["per_instance_psi", "uniform_psi"])
def test_schedule_interpreter_in_other_semirings(semiring, uniform):
    """The compiled plan is semiring-agnostic: interpreted with another (+, x) pair it
    reproduces the semiring oracle, with evidence, uniform mode and direct marginals."""
    from junctiontree import schedule as sch
    net = wl.random_dag(12, 3, 2, 3, 8, 5)
    tree, seps, mc, f2c, eff, evars = compile_net(net)
    plan = sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"])
    B = 3
    ev = wl.draw_evidence(net, B)
    vals = semiring_inputs(net["values"], semiring)
    outs, ys = ref_fixed.propagate_batch(tree, seps, mc, f2c, net["factors"], net["sizes"], vals, evars, ev, n=B,
                                         semiring=semiring)
    for beliefs in (True, False):
        work, fout = plan_interp.run(plan, B, factor_in=plan_interp.flatten_factors(plan, vals), evidence=ev,
                                     uniform=uniform, beliefs=beliefs, semiring=semiring)
        if beliefs:
            for k in range(len(mc) + len(seps)):
                assert_close_semiring(plan_interp.node_array(plan, work, k, B), ys[k], 1e-13, semiring, "node %d" % k)
        for f in range(len(net["factors"])):
            assert_close_semiring(plan_interp.factor_array(plan, fout, f, B), outs[f], 1e-13, semiring,
                                  "factor %d" % f)


def test_semiring_operator_surface():
    """Same plugin surface as SumProduct (reference sum_product.py:6-35) for every law; a user
    einsum function is routed per operator, and the module-level instances are device-bound."""
    from junctiontree import _native, computation as comp, semirings as sr
    from junctiontree.sum_product import SumProduct
    for law, flag in ((sr.max_product, _native.JT_SR_MAX_PRODUCT), (sr.log_sum_exp, _native.JT_SR_LOG_SUM_EXP),
                      (sr.max_sum, _native.JT_SR_MAX_SUM)):
        assert isinstance(law, SumProduct) and law.on_device and law.semiring_flag == flag
    assert comp.max_product is sr.max_product and comp.sum_product.semiring_flag == 0

    def np_max_product(*args):
        return ref_fixed._einsum(*args, semiring="max_product")

    law = sr.MaxProduct(np_max_product)
    assert not law.on_device
    net = wl.huang_darwiche()
    tree, seps, mc, f2c, eff, evars = compile_net(net, with_evidence=False)
    psi = ref_fixed.evaluate(net["factors"], net["values"], mc, f2c, net["sizes"], "max_product")
    ones = [np.ones([net["sizes"][v] for v in s]) for s in seps]
    got = comp.compute_beliefs(tree, psi + ones, mc + seps, law)       # plugin path: no device needed
    want = ref_fixed.compute_beliefs(tree, psi + ones, mc + seps, "max_product")
    for k, (g, w) in enumerate(zip(got, want)):
        assert_close(g, w, 1e-13, "node %d" % k)
    with pytest.raises(NotImplementedError):
        sr.log_sum_exp.ratio(np.ones(2), np.ones(2))


def test_header_declares_the_semiring_flags():
    from junctiontree import _native
    header = open(os.path.join(os.path.dirname(__file__), "..", "include", "jt_b200.h")).read()
    for name in ("JT_SR_SUM_PRODUCT", "JT_SR_MAX_PRODUCT", "JT_SR_LOG_SUM_EXP", "JT_SR_MAX_SUM", "JT_SR_MASK"):
        assert ("#define %s 0x%03x" % (name, getattr(_native, name))) in header
