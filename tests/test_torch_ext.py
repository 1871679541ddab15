"""The PyTorch extension over the C ABI (csrc/jt_torch.cpp -> torch.ops.jt_b200.*) on a machine
without a GPU: it loads, registers every operator with the documented schema, shares the library
instance and the plan handles with the ctypes binding, and refuses CPU tensors (there is no CPU
path).  No compute call is made."""

import os

import numpy as np
import pytest
import torch

import jt_workloads as wl
from helpers import compile_net
from junctiontree import _native
from junctiontree import schedule as sch
from junctiontree import torch_ops


def _plan(net, with_evidence=True):
    tree, seps, mc, f2c, eff, evars = compile_net(net, with_evidence)
    return sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"])


def test_extension_loads_and_registers_every_operator_with_its_schema():
    ops = torch_ops.ops()
    assert os.path.exists(torch_ops.extension_path())
    assert ops.abi_version() == _native.ABI_VERSION
    for name, schema in torch_ops.SCHEMAS.items():
        op = getattr(ops, name)
        assert str(op.default._schema) == schema, name
    # nothing registered that the Python side does not know about
    registered = {s.name.split("::")[1] for s in torch._C._jit_get_all_schemas() if s.name.startswith("jt_b200::")}
    assert registered == set(torch_ops.SCHEMAS)


def test_both_bindings_share_one_library_instance():
    """The extension links libjt_b200.so through $ORIGIN; ctypes loads it by absolute path.  Both
    must resolve to ONE mapping (the library keeps process-wide state: launch counter, plans)."""
    torch_ops.ops()
    _native.lib()
    with open("/proc/self/maps") as fh:
        paths = {line.split()[-1] for line in fh if "libjt_b200.so" in line}
    assert paths == {os.path.realpath(_native.library_path())}
    assert torch_ops.ops().launch_count() == _native.launch_count()


def test_errors_of_the_library_surface_as_python_exceptions():
    ops = torch_ops.ops()
    with pytest.raises(RuntimeError, match="plan blob too short or misaligned"):
        ops.plan_create(torch.zeros(100, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="contiguous CPU uint8"):
        ops.plan_create(torch.zeros(100, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="null plan handle"):
        ops.plan_query(0, sch.H_NCLIQUES)


@pytest.mark.parametrize("net", [wl.sprinkler(), wl.dag37()], ids=lambda n: n["name"])
def test_plan_handles_are_interchangeable_between_the_bindings(net):
    plan = _plan(net)
    blob = plan.to_blob()
    tp = torch_ops.TorchPlan(blob)
    dp = _native.DevicePlan(blob)
    try:
        for what in (sch.H_NCLIQUES, sch.H_NSEPS, sch.H_CLIQUE_ENTRIES, sch.H_SEP_ENTRIES, sch.H_NTASKS):
            assert tp.query(what) == dp.query(what)
            # a handle created by ctypes through the torch operators, and the other way round
            assert torch_ops.ops().plan_query(torch_ops.handle_of(dp), what) == dp.query(what)
        for B in (1, 96, 4096):
            for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32)):
                assert tp.workspace_bytes(B, tdt) == dp.workspace_bytes(B, ndt)
        out = np.zeros(1, np.int64)
        import ctypes
        _native.check(_native.lib().jt_plan_query(ctypes.c_void_p(tp.handle_int), sch.H_NCLIQUES,
                                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
        assert out[0] == plan.n_cliques
        with pytest.raises(RuntimeError, match="float32 and float64"):
            tp.workspace_bytes(8, torch.float16)
    finally:
        tp.close()
        dp.close()
    with pytest.raises(_native.NativeError, match="destroyed"):
        torch_ops.handle_of(dp)


def test_stage_operators_refuse_cpu_tensors():
    """No CPU fallback: every stage operator fails on CPU tensors before touching the library."""
    plan = _plan(wl.sprinkler(), with_evidence=False)
    tp = torch_ops.TorchPlan(plan.to_blob())
    B = 4
    ws = torch.zeros(tp.workspace_bytes(B, torch.float64), dtype=torch.uint8)
    factors = torch.ones(plan.fin_entries, dtype=torch.float64)
    fout = torch.zeros((plan.fout_entries, B), dtype=torch.float64)
    try:
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            tp.propagate(factors, None, ws, fout, B)
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            tp.init(factors, None, ws, B)
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            tp.collect(ws, B, torch.float64)
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            tp.distribute(ws, B, torch.float64)
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            tp.marginal(ws, fout, B)
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            tp.normalize(fout, None, B)
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            tp.evidence_errors(ws, B, torch.float64)
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            torch_ops.ops().ratio(factors, factors)
        with pytest.raises(RuntimeError, match="float32 and float64"):
            tp.init(factors.to(torch.int64), None, ws, B)
    finally:
        tp.close()


def test_engine_binding_switch(monkeypatch):
    from junctiontree import engine as eng
    plan = _plan(wl.sprinkler(), with_evidence=False)
    assert eng.Engine(plan).binding == "ctypes"
    monkeypatch.setenv("JT_BINDING", "torch")
    assert eng.Engine(plan).binding == "torch"
    monkeypatch.setenv("JT_BINDING", "numpy")
    with pytest.raises(ValueError, match="JT_BINDING"):
        eng.Engine(plan)


def test_both_bindings_release_the_gil_during_a_call():
    """"the torch binding releases the GIL".  Loading the Ising 16x16 plan takes a few
    hundred milliseconds inside the library (jt_plan_create derives the dense contractions and the
    belief walk tables); a Python thread must keep running meanwhile, through either binding."""
    import threading
    import jt_bench_lib as bl
    net = bl.make_net("ising16")
    tree, seps, mc, f2c, eff, evars = compile_net(net)
    blob = sch.Plan(tree, mc + seps, eff, net["factors"], f2c, evars, net["sizes"]).to_blob()
    tensor = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    ops = torch_ops.ops()

    def through_torch():
        ops.plan_destroy(ops.plan_create(tensor))

    def through_ctypes():
        _native.DevicePlan(blob).close()

    for call in (through_torch, through_ctypes):
        done, failed = [], []

        def work(call=call):
            try:
                call()
            except Exception as exc:          # reported below; the spin loop must still end
                failed.append(exc)
            finally:
                done.append(True)

        worker = threading.Thread(target=work)
        worker.start()
        spins = 0
        while not done:
            spins += 1
        worker.join()
        assert not failed, failed
        assert spins > 1000, "%s held the GIL for the whole call" % call.__name__
