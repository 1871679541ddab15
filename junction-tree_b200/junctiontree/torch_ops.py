"""PyTorch-extension binding of the C ABI: ``torch.ops.jt_b200.*`` (``csrc/jt_torch.cpp``).

The second of the two bindings of ``libjt_b200.so`` (the first is ``_native.py``, ctypes).  The
operators take tensors instead of raw pointers, validate device, dtype and sizes against the
plan in C++, and enqueue on torch's current CUDA stream of the workspace's device -- so they
compose with ``torch.cuda.stream(...)`` blocks and ``torch.cuda.graph`` capture without any
pointer or stream handling in Python.  Both bindings drive the same library instance in the
process: a plan handle created through one is valid in the other.

There is no CPU implementation behind these operators: a CPU tensor is an error.

Select it for the engine's stage calls with ``JT_BINDING=torch`` (or ``engine.binding = "torch"``);
the default stays ctypes, whose per-call cost is lower for single small propagations.
"""

import os

from . import _native

_EXT_NAME = "_jt_torch.so"
_ops = None


def extension_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _EXT_NAME)


def ops():
    """Load (once) the extension and return the ``torch.ops.jt_b200`` namespace."""
    global _ops
    if _ops is None:
        import torch
        path = extension_path()
        if not os.path.exists(path):
            raise _native.NativeError(
                "%s not found: build it with `make -C junction-tree_b200/csrc` (or __graft_entry__.build())" % path)
        _native.lib()                      # same library instance for both bindings
        torch.ops.load_library(path)
        ns = torch.ops.jt_b200
        if ns.abi_version() != _native.ABI_VERSION:
            raise _native.NativeError("_jt_torch.so was built against ABI %d, this package expects %d"
                                      % (ns.abi_version(), _native.ABI_VERSION))
        _ops = ns
    return _ops


#: operator name -> schema, as registered by csrc/jt_torch.cpp (checked by the CPU tests)
SCHEMAS = {
    "abi_version": "jt_b200::abi_version() -> int",
    "launch_count": "jt_b200::launch_count() -> int",
    "plan_create": "jt_b200::plan_create(Tensor blob) -> int",
    "plan_destroy": "jt_b200::plan_destroy(int plan) -> ()",
    "plan_query": "jt_b200::plan_query(int plan, int what) -> int",
    "plan_upload": "jt_b200::plan_upload(int plan, int device) -> ()",
    "workspace_bytes": "jt_b200::workspace_bytes(int plan, int B, ScalarType dtype) -> int",
    "init": "jt_b200::init(int plan, Tensor factor_tables, bool factors_batched, Tensor? evidence, "
            "Tensor(a!) workspace, int B, int flags) -> ()",
    "collect": "jt_b200::collect(int plan, Tensor(a!) workspace, int B, ScalarType dtype, int flags) -> ()",
    "distribute": "jt_b200::distribute(int plan, Tensor(a!) workspace, int B, ScalarType dtype, int flags) -> ()",
    "marginal": "jt_b200::marginal(int plan, Tensor(a!) workspace, Tensor(b!) factor_out, int B, int flags) -> ()",
    "propagate": "jt_b200::propagate(int plan, Tensor factor_tables, bool factors_batched, Tensor? evidence, "
                 "Tensor(a!) workspace, Tensor(b!)? factor_out, int B, int flags) -> ()",
    "normalize": "jt_b200::normalize(int plan, Tensor(a!) factor_out, Tensor(b!)? logz, int B, int flags) -> ()",
    "evidence_errors": "jt_b200::evidence_errors(int plan, Tensor workspace, int B, ScalarType dtype) -> int",
    "ratio": "jt_b200::ratio(Tensor new_values, Tensor old_values) -> Tensor",
}


def handle_of(plan):
    """Integer handle of a ``_native.DevicePlan`` (or an int from ``plan_create``)."""
    if isinstance(plan, int):
        return plan
    value = plan.handle.value
    if not value:
        raise _native.NativeError("the plan has been destroyed")
    return int(value)


class TorchPlan:
    """A plan owned through the extension: ``plan_create`` on a blob (``schedule.Plan.to_blob``),
    stage calls on tensors.  Mirrors ``_native.DevicePlan`` for the calls a tensor program
    needs."""

    def __init__(self, blob):
        import torch
        self._blob = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
        self.handle_int = ops().plan_create(self._blob)

    def close(self):
        if self.handle_int:
            ops().plan_destroy(self.handle_int)
            self.handle_int = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, what):
        return ops().plan_query(self.handle_int, what)

    def workspace_bytes(self, B, dtype):
        return ops().workspace_bytes(self.handle_int, B, dtype)

    def upload(self, device=None):
        import torch
        ops().plan_upload(self.handle_int, torch.cuda.current_device() if device is None else int(device))

    def new_workspace(self, B, dtype, device="cuda"):
        import torch
        return torch.zeros(self.workspace_bytes(B, dtype), dtype=torch.uint8, device=device)

    def propagate(self, factors, evidence, workspace, factor_out, B, flags=0, factors_batched=False):
        ops().propagate(self.handle_int, factors, factors_batched, evidence, workspace, factor_out, B, flags)

    def init(self, factors, evidence, workspace, B, flags=0, factors_batched=False):
        ops().init(self.handle_int, factors, factors_batched, evidence, workspace, B, flags)

    def collect(self, workspace, B, dtype, flags=0):
        ops().collect(self.handle_int, workspace, B, dtype, flags)

    def distribute(self, workspace, B, dtype, flags=0):
        ops().distribute(self.handle_int, workspace, B, dtype, flags)

    def marginal(self, workspace, factor_out, B, flags=0):
        ops().marginal(self.handle_int, workspace, factor_out, B, flags)

    def normalize(self, factor_out, logz, B, flags=0):
        ops().normalize(self.handle_int, factor_out, logz, B, flags)

    def evidence_errors(self, workspace, B, dtype):
        return ops().evidence_errors(self.handle_int, workspace, B, dtype)
