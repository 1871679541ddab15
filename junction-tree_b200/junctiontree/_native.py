"""ctypes binding of ``libjt_b200.so`` (C ABI: ``include/jt_b200.h``).

The library is built in-tree by ``junction-tree_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback: if the library is missing or no CUDA device is present, the compute entry
points raise.
"""

import ctypes
import os

import numpy as np

JT_OK = 0
JT_F32, JT_F64 = 0, 1
JT_SEP_BELIEFS, JT_SKIP_MARGINAL, JT_UNIFORM, JT_NO_UNIFORM, JT_UNIFORM_VALID, JT_NO_BELIEFS = 1, 2, 4, 8, 16, 32
# semiring bits of the stage flags (include/jt_b200.h JT_SR_*)
JT_SR_SUM_PRODUCT, JT_SR_MAX_PRODUCT, JT_SR_LOG_SUM_EXP, JT_SR_MAX_SUM, JT_SR_MASK = 0x000, 0x100, 0x200, 0x300, 0x300
ABI_VERSION = 7

_LIB_NAME = "libjt_b200.so"
_lib = None

_c_void_pp = ctypes.POINTER(ctypes.c_void_p)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)

#: name -> (restype, argtypes); every symbol include/jt_b200.h declares
SIGNATURES = {
    "jt_abi_version": (ctypes.c_int, []),
    "jt_last_error_string": (ctypes.c_char_p, []),
    "jt_launch_count": (ctypes.c_int64, []),
    "jt_plan_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, _c_void_pp]),
    "jt_plan_destroy": (None, [ctypes.c_void_p]),
    "jt_plan_query": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _i64p]),
    "jt_plan_node_range": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _i64p, _i64p]),
    "jt_plan_message_offsets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _i64p, _i64p]),
    "jt_workspace_bytes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_size_t)]),
    "jt_workspace_layout": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, _i64p]),
    "jt_plan_upload": (ctypes.c_int, [ctypes.c_void_p]),
    "jt_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "jt_collect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_int, ctypes.c_void_p]),
    "jt_distribute": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_void_p]),
    "jt_marginal": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "jt_propagate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_int, ctypes.c_void_p]),
    "jt_normalize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "jt_evidence_errors": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, _i64p]),
    "jt_copy_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                    ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "jt_ratio": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                ctypes.c_int, ctypes.c_void_p]),
    "jt_contract": (ctypes.c_int, [_c_void_pp, ctypes.c_int, _i32p, ctypes.c_int64, _i32p,
                                   ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_void_p]),
}


class NativeError(RuntimeError):
    pass


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise NativeError(
                "%s not found: build it with `make -C junction-tree_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU fallback" % path)
        handle = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc):
    if rc != JT_OK:
        raise NativeError("libjt_b200: %s (code %d)" % (lib().jt_last_error_string().decode(), rc))


def dtype_code(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return JT_F64
    if dtype == np.float32:
        return JT_F32
    raise TypeError("only float32 and float64 potentials are supported, got %s" % dtype)


def launch_count():
    return int(lib().jt_launch_count())


class DevicePlan:
    """Owns a ``jt_plan`` created from a schedule blob."""

    def __init__(self, blob):
        self._handle = ctypes.c_void_p()
        self._blob = bytes(blob)
        check(lib().jt_plan_create(self._blob, len(self._blob), ctypes.byref(self._handle)))
        self.uploaded = False

    def close(self):
        if self._handle:
            lib().jt_plan_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._handle

    def query(self, what):
        out = ctypes.c_int64()
        check(lib().jt_plan_query(self._handle, what, ctypes.byref(out)))
        return out.value

    def node_range(self, node):
        off, n = ctypes.c_int64(), ctypes.c_int64()
        check(lib().jt_plan_node_range(self._handle, node, ctypes.byref(off), ctypes.byref(n)))
        return off.value, n.value

    def message_offsets(self, sep_node):
        up, down = ctypes.c_int64(), ctypes.c_int64()
        check(lib().jt_plan_message_offsets(self._handle, sep_node, ctypes.byref(up), ctypes.byref(down)))
        return up.value, down.value

    def workspace_bytes(self, B, dtype):
        out = ctypes.c_size_t()
        check(lib().jt_workspace_bytes(self._handle, B, dtype_code(dtype), ctypes.byref(out)))
        return out.value

    def workspace_layout(self, B, dtype):
        """Byte offsets ``{fbase, errors, uniform, total}`` inside a workspace."""
        out = (ctypes.c_int64 * 4)()
        check(lib().jt_workspace_layout(self._handle, B, dtype_code(dtype), out))
        return {"fbase": out[0], "errors": out[1], "uniform": out[2], "total": out[3]}

    def upload(self):
        if not self.uploaded:
            check(lib().jt_plan_upload(self._handle))
            self.uploaded = True

    # stage calls: raw device pointers (ints) and a cudaStream_t (int)
    def init(self, factors_ptr, batched, evidence_ptr, B, dtype, ws_ptr, flags, stream):
        check(lib().jt_init(self._handle, factors_ptr, int(batched), evidence_ptr, B, dtype_code(dtype),
                            ws_ptr, flags, stream))

    def collect(self, B, dtype, ws_ptr, flags, stream):
        check(lib().jt_collect(self._handle, B, dtype_code(dtype), ws_ptr, flags, stream))

    def distribute(self, B, dtype, ws_ptr, flags, stream):
        check(lib().jt_distribute(self._handle, B, dtype_code(dtype), ws_ptr, flags, stream))

    def marginal(self, B, dtype, ws_ptr, out_ptr, stream, flags=0):
        check(lib().jt_marginal(self._handle, B, dtype_code(dtype), ws_ptr, out_ptr, flags, stream))

    def propagate(self, factors_ptr, batched, evidence_ptr, B, dtype, ws_ptr, out_ptr, flags, stream):
        check(lib().jt_propagate(self._handle, factors_ptr, int(batched), evidence_ptr, B,
                                 dtype_code(dtype), ws_ptr, out_ptr, flags, stream))

    def normalize(self, B, dtype, out_ptr, logz_ptr, stream, flags=0):
        check(lib().jt_normalize(self._handle, B, dtype_code(dtype), out_ptr, logz_ptr, flags, stream))

    def evidence_errors(self, B, dtype, ws_ptr, stream):
        out = ctypes.c_int64()
        check(lib().jt_evidence_errors(self._handle, B, dtype_code(dtype), ws_ptr, stream, ctypes.byref(out)))
        return out.value


def contract(op_ptrs, tables, maps, n_s, n_r, n_slo, n_rlo, B, dtype, out_ptr, stream, flags=0):
    """``jt_contract``: out[s] = sum_r prod_j op_j[A_j(s) + B_j(r)] (``flags``: JT_SR_* semiring)."""
    n = len(op_ptrs)
    ops = (ctypes.c_void_p * n)(*op_ptrs)
    tables = np.ascontiguousarray(tables, np.int32)
    maps = np.ascontiguousarray(maps, np.int32)
    check(lib().jt_contract(ops, n, tables.ctypes.data_as(_i32p), tables.size,
                            maps.ctypes.data_as(_i32p), n_s, n_r, n_slo, n_rlo, B, dtype_code(dtype),
                            out_ptr, flags, stream))


def ratio(new_ptr, old_ptr, out_ptr, n, dtype, stream):
    """``jt_ratio``: out = new / old with x / 0 = 0."""
    check(lib().jt_ratio(new_ptr, old_ptr, out_ptr, n, dtype_code(dtype), stream))


def copy_rows(dst_ptr, dst_pitch, src_ptr, src_pitch, width_bytes, rows, to_host, stream):
    """``jt_copy_rows``: strided row copy device <-> pinned host on ``stream``."""
    check(lib().jt_copy_rows(dst_ptr, dst_pitch, src_ptr, src_pitch, width_bytes, rows, int(to_host), stream))
