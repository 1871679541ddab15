"""The PyTorch-extension binding (torch.ops.jt_b200.*, csrc/jt_torch.cpp) on the GPU: the same
propagation through both bindings of the C ABI, the staged operators on tensors against the
one-call operator and the oracle, argument checks, and capture into a CUDA graph."""

import numpy as np
import pytest

import jt_workloads as wl
from helpers import RTOL_F64, assert_close
from junctiontree import _native

pytestmark = pytest.mark.gpu


def _oracle(tree, net, evars, ev, n):
    from oracle import ref_fixed
    ct = tree.clique_tree
    return ref_fixed.propagate_batch(tree.tree, tree.separators, ct.maxcliques, ct.factor_to_maxclique,
                                     net["factors"], net["sizes"], net["values"], evars, ev, n=n)


def _flat_factors(plan, values, torch):
    host = np.empty(plan.fin_entries, np.float64)
    for f, v in enumerate(values):
        host[plan.fin_off[f]:plan.fin_off[f] + plan.fin_size[f]] = np.asarray(v, np.float64).reshape(-1)
    return torch.from_numpy(host).to("cuda")


def _factor_arrays(plan, fout, B):
    """[fout_entries, B] device tensor -> per-factor [B, *shape] NumPy arrays."""
    host = fout.cpu().numpy()
    return [np.moveaxis(host[plan.fout_off[f]:plan.fout_off[f] + plan.fout_size[f]]
                        .reshape(tuple(plan.fout_shape[f]) + (B,)), -1, 0) for f in range(len(plan.fout_off))]


@pytest.mark.parametrize("B", [3, 300])
def test_propagate_batch_through_the_torch_binding_equals_ctypes_and_the_oracle(B, monkeypatch):
    import junctiontree as jt
    net = wl.random_dag(60, 3, 2, 5, 8, 3)
    evars = net["evidence_vars"]
    ev = wl.draw_evidence(net, B)
    tree_c = jt.create_junction_tree(net["factors"], net["sizes"])
    outs_c, nodes_c = tree_c.propagate_batch(net["values"], evars, ev, nodes=True)
    monkeypatch.setenv("JT_BINDING", "torch")
    tree_t = jt.create_junction_tree(net["factors"], net["sizes"])          # its own engines
    plan = tree_t.plan(evars)
    assert tree_t._engine(plan.sizes, evars, plan.full_sizes).binding == "torch"
    before = _native.launch_count()
    outs_t, nodes_t = tree_t.propagate_batch(net["values"], evars, ev, nodes=True)
    assert _native.launch_count() > before, "the torch operators did not launch the library's kernels"
    for k, (a, b) in enumerate(zip(list(outs_t) + list(nodes_t), list(outs_c) + list(nodes_c))):
        assert_close(a, b, 1e-13, "torch binding vs ctypes, entry %d" % k)
    n = min(B, 3)
    want_f, want_n = _oracle(tree_t, net, evars, ev[:n], n)
    for k, w in enumerate(want_n):
        assert_close(nodes_t[k][:n], w, RTOL_F64, "torch binding node %d" % k)
    for f, w in enumerate(want_f):
        assert_close(outs_t[f][:n], w, RTOL_F64, "torch binding factor %d" % f)


def test_staged_operators_on_tensors_equal_the_one_call_operator_and_the_oracle():
    import torch
    import junctiontree as jt
    from junctiontree import torch_ops
    net = wl.random_dag(24, 3, 2, 4, 6, 7)
    evars = net["evidence_vars"]
    B = 200
    ev = wl.draw_evidence(net, B)
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    plan = tree.plan(evars)
    tp = torch_ops.TorchPlan(plan.to_blob())
    tp.upload()
    factors = _flat_factors(plan, net["values"], torch)
    evidence = torch.from_numpy(ev).to("cuda")
    f64 = torch.float64
    # one call
    ws1 = tp.new_workspace(B, f64)
    out1 = torch.empty((plan.fout_entries, B), dtype=f64, device="cuda")
    tp.propagate(factors, evidence, ws1, out1, B, _native.JT_SEP_BELIEFS)
    # the four stages, with the flags jt_propagate passes for shared tables
    flags = _native.JT_UNIFORM | _native.JT_SEP_BELIEFS
    ws2 = tp.new_workspace(B, f64)
    out2 = torch.empty_like(out1)
    tp.init(factors, evidence, ws2, B, flags)
    tp.collect(ws2, B, f64, flags)
    tp.distribute(ws2, B, f64, flags)
    tp.marginal(ws2, out2, B, flags)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out2.cpu().numpy(), out1.cpu().numpy(), rtol=1e-13, atol=0)
    assert tp.evidence_errors(ws1, B, f64) == 0
    got = _factor_arrays(plan, out1, B)
    want_f, want_n = _oracle(tree, net, evars, ev[:4], 4)
    for f, w in enumerate(want_f):
        assert_close(got[f][:4], w, RTOL_F64, "torch staged factor %d" % f)
    # clique and separator beliefs straight out of the workspace tensor ([entries][B] block first)
    work = ws1[: plan.work_entries * B * 8].view(f64).view(plan.work_entries, B).cpu().numpy()
    for k, w in enumerate(want_n):
        rows = work[plan.node_off[k]:plan.node_off[k] + plan.node_size[k]]
        assert_close(np.moveaxis(rows.reshape(tuple(plan.node_shape[k]) + (B,)), -1, 0)[:4], w, RTOL_F64,
                     "torch staged node %d" % k)
    # output stage: every scope sums to one, log Z = log of the unnormalised total of scope 0
    total0 = got[0].reshape(B, -1).sum(axis=1)
    logz = torch.empty(B, dtype=f64, device="cuda")
    tp.normalize(out1, logz, B)
    torch.cuda.synchronize()
    for f, a in enumerate(_factor_arrays(plan, out1, B)):
        np.testing.assert_allclose(a.reshape(B, -1).sum(axis=1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(logz.cpu().numpy(), np.log(total0), rtol=1e-12, atol=1e-12)
    # out-of-range states are counted, not read out of bounds
    bad = ev.copy()
    bad[5, 0] = 99
    tp.propagate(factors, torch.from_numpy(bad).to("cuda"), ws1, out2, B, 0)
    assert tp.evidence_errors(ws1, B, f64) >= 1       # once per factor that contains the variable
    tp.close()


def test_operator_argument_checks_on_the_device():
    import torch
    import junctiontree as jt
    from junctiontree import torch_ops
    net = wl.random_dag(24, 3, 2, 4, 6, 7)
    evars = net["evidence_vars"]
    B = 16
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    plan = tree.plan(evars)
    tp = torch_ops.TorchPlan(plan.to_blob())
    factors = _flat_factors(plan, net["values"], torch)
    evidence = torch.from_numpy(wl.draw_evidence(net, B)).to("cuda")
    out = torch.empty((plan.fout_entries, B), dtype=torch.float64, device="cuda")
    ws = torch.zeros(tp.workspace_bytes(B, torch.float64), dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="not uploaded|upload"):
        tp.propagate(factors, evidence, ws, out, B)
    tp.upload()
    with pytest.raises(RuntimeError, match="workspace holds"):
        tp.propagate(factors, evidence, ws[:-1024], out, B)
    with pytest.raises(RuntimeError, match="factor_tables has"):
        tp.propagate(factors[:-1], evidence, ws, out, B)
    with pytest.raises(RuntimeError, match="evidence must be int32"):
        tp.propagate(factors, evidence.to(torch.int64), ws, out, B)
    with pytest.raises(RuntimeError, match="evidence must hold"):
        tp.propagate(factors, evidence[:-1], ws, out, B)
    with pytest.raises(RuntimeError, match="no evidence was given"):
        tp.propagate(factors, None, ws, out, B)
    with pytest.raises(RuntimeError, match="factor_out has"):
        tp.propagate(factors, evidence, ws, out[:, :-1].contiguous(), B)
    with pytest.raises(RuntimeError, match="needs JT_SKIP_MARGINAL"):
        tp.propagate(factors, evidence, ws, None, B)
    with pytest.raises(RuntimeError, match="must be contiguous"):
        tp.propagate(factors, evidence, ws, out.t(), B)
    tp.propagate(factors, evidence, ws, None, B, _native.JT_SKIP_MARGINAL)       # and the valid forms run
    tp.propagate(factors, evidence, ws, out, B)
    torch.cuda.synchronize()
    # the Hugin ratio operator (SumProduct.absorb(old=...)): x / 0 = 0
    new = torch.tensor([1.0, 2.0, 0.0, 3.0], dtype=torch.float64, device="cuda")
    old = torch.tensor([2.0, 0.0, 0.0, 4.0], dtype=torch.float64, device="cuda")
    assert torch_ops.ops().ratio(new, old).cpu().tolist() == [0.5, 0.0, 0.0, 0.75]
    tp.close()


def test_operators_enqueue_on_the_current_stream_and_capture_into_a_cuda_graph():
    import torch
    import junctiontree as jt
    from junctiontree import torch_ops
    net = wl.dag37()
    evars = net["evidence_vars"]
    B = 1
    tree = jt.create_junction_tree(net["factors"], net["sizes"])
    plan = tree.plan(evars)
    tp = torch_ops.TorchPlan(plan.to_blob())
    tp.upload()
    factors = _flat_factors(plan, net["values"], torch)
    ev_all = wl.draw_evidence(net, 8)
    evidence = torch.from_numpy(ev_all[:1].copy()).to("cuda")
    ws = tp.new_workspace(B, torch.float64)
    out = torch.zeros((plan.fout_entries, B), dtype=torch.float64, device="cuda")
    flags = _native.JT_NO_BELIEFS
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        tp.propagate(factors, evidence, ws, out, B, flags)          # warm-up on the side stream
    stream.synchronize()
    first = out.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        tp.propagate(factors, evidence, ws, out, B, flags)
    want_f, _ = _oracle(tree, net, evars, ev_all, 8)
    for b in range(8):
        evidence.copy_(torch.from_numpy(ev_all[b:b + 1].copy()))
        out.zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            graph.replay()
        stream.synchronize()
        if b == 0:
            np.testing.assert_allclose(out.cpu().numpy(), first.cpu().numpy(), rtol=1e-13, atol=0)
        got = _factor_arrays(plan, out, B)
        for f, w in enumerate(want_f):
            assert_close(got[f][0], w[b], RTOL_F64, "graph replay factor %d" % f)
    tp.close()
