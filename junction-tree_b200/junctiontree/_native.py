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
JT_NO_DENSE, JT_LOGZ_ONLY = 64, 128
# semiring bits of the stage flags (include/jt_b200.h JT_SR_*)
JT_SR_SUM_PRODUCT, JT_SR_MAX_PRODUCT, JT_SR_LOG_SUM_EXP, JT_SR_MAX_SUM, JT_SR_MASK = 0x000, 0x100, 0x200, 0x300, 0x300
ABI_VERSION = 9

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
    "jt_plan_dense_count": (ctypes.c_int, [ctypes.c_void_p]),
    "jt_plan_dense_get": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _i64p]),
    "jt_plan_dense_table": (_i32p, [ctypes.c_void_p, _i64p]),
    "jt_workspace_sparse_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _i64p, ctypes.c_int64, _i64p]),
    "jt_workspace_sparse_bytes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "jt_workspace_sparse_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                  _c_void_pp]),
    "jt_workspace_sparse_ptr": (ctypes.c_void_p, [ctypes.c_void_p]),
    "jt_workspace_sparse_mapped": (ctypes.c_size_t, [ctypes.c_void_p]),
    "jt_workspace_sparse_destroy": (None, [ctypes.c_void_p]),
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
    "jt_plan_single_launch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    "jt_propagate_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_int, ctypes.c_void_p]),
    "jt_beliefs_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
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
    # host compile phase (jt_compile.cpp; no CUDA)
    "jt_ibuf_count": (ctypes.c_int, [ctypes.c_void_p]),
    "jt_ibuf_size": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int]),
    "jt_ibuf_data": (_i32p, [ctypes.c_void_p, ctypes.c_int]),
    "jt_ibuf_destroy": (None, [ctypes.c_void_p]),
    "jt_free": (None, [ctypes.c_void_p]),
    "jt_triangulate": (ctypes.c_int, [ctypes.c_int32, _i64p, ctypes.c_int32, _i32p, _i32p, _i32p, ctypes.c_int32,
                                      _c_void_pp]),
    "jt_junction_tree": (ctypes.c_int, [ctypes.c_int32, _i64p, ctypes.c_int32, _i32p, _i32p, ctypes.c_int32,
                                        _c_void_pp]),
    "jt_plan_build": (ctypes.c_int, [ctypes.c_int32, _i64p, _i64p, ctypes.c_int32, ctypes.c_int32, _i32p, _i32p,
                                     ctypes.c_int32, _i32p, _i32p, _i32p, ctypes.c_int32, _i32p, _i32p, _i32p,
                                     ctypes.c_int32, _i32p, ctypes.c_int32, _i32p, _i32p, ctypes.c_int32, _i32p,
                                     _c_void_pp, ctypes.POINTER(ctypes.c_size_t)]),
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


    DENSE_FIELDS = ("task", "msg", "n_g", "n_i", "K", "n_q", "MT", "n_it", "n_k4", "s_of", "mg", "mk", "r_of",
                    "w_off", "w_size", "launch")

    def dense_tasks(self):
        """The dense contractions derived from the plan (``jt_plan_dense_*``): a list of dicts
        (``DENSE_FIELDS``) and the int32 table their offsets index."""
        n = lib().jt_plan_dense_count(self._handle)
        out = []
        for k in range(n):
            buf = (ctypes.c_int64 * 16)()
            check(lib().jt_plan_dense_get(self._handle, k, buf))
            out.append(dict(zip(self.DENSE_FIELDS, list(buf))))
        count = ctypes.c_int64()
        ptr = lib().jt_plan_dense_table(self._handle, ctypes.byref(count))
        table = np.ctypeslib.as_array(ptr, shape=(count.value,)).copy() if count.value else np.zeros(0, np.int32)
        return out, table

    def sparse_bytes(self, B, dtype, flags):
        """``(mapped, dense)`` bytes of a sparse workspace for the stages run with ``flags``."""
        mapped, dense = ctypes.c_size_t(), ctypes.c_size_t()
        check(lib().jt_workspace_sparse_bytes(self._handle, B, dtype_code(dtype), flags, ctypes.byref(mapped),
                                              ctypes.byref(dense)))
        return mapped.value, dense.value

    def sparse_rows(self, flags):
        """Merged ``[begin, end)`` entry intervals the stages touch per instance with ``flags``."""
        count = ctypes.c_int64()
        check(lib().jt_workspace_sparse_rows(self._handle, flags, None, 0, ctypes.byref(count)))
        buf = (ctypes.c_int64 * (2 * max(count.value, 1)))()
        check(lib().jt_workspace_sparse_rows(self._handle, flags, buf, count.value, ctypes.byref(count)))
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(count.value)]

    def sparse_workspace(self, B, dtype, flags):
        return SparseWorkspace(self, B, dtype, flags)

    def upload(self):
        """Upload the descriptors to the current device.  Always goes through the library: the
        call returns at once when the plan already lives on the current device and fails cleanly
        when it lives on another one (a plan belongs to one device; engines are cached per
        device)."""
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

    def single_launch(self, B, flags):
        """True when ``propagate`` of ``B`` instances is one launch of the whole-propagation kernel."""
        return bool(lib().jt_plan_single_launch(self._handle, B, flags))

    def propagate_host(self, host_factors_ptr, factor_bytes, host_evidence_ptr, B, dtype, dev_factors_ptr,
                       dev_evidence_ptr, ws_ptr, dev_out_ptr, host_out_ptr, out_bytes, flags, stream):
        check(lib().jt_propagate_host(self._handle, host_factors_ptr, factor_bytes, host_evidence_ptr, B,
                                      dtype_code(dtype), dev_factors_ptr, dev_evidence_ptr, ws_ptr, dev_out_ptr,
                                      host_out_ptr, out_bytes, flags, stream))

    def beliefs_host(self, host_potentials_ptr, dtype, ws_ptr, host_beliefs_ptr, flags, stream):
        check(lib().jt_beliefs_host(self._handle, host_potentials_ptr, dtype_code(dtype), ws_ptr, host_beliefs_ptr,
                                    flags, stream))

    def normalize(self, B, dtype, out_ptr, logz_ptr, stream, flags=0):
        check(lib().jt_normalize(self._handle, B, dtype_code(dtype), out_ptr, logz_ptr, flags, stream))

    def evidence_errors(self, B, dtype, ws_ptr, stream):
        out = ctypes.c_int64()
        check(lib().jt_evidence_errors(self._handle, B, dtype_code(dtype), ws_ptr, stream, ctypes.byref(out)))
        return out.value


class SparseWorkspace:
    """A workspace whose untouched rows have no memory behind them (``jt_workspace_sparse_*``).
    ``data_ptr()`` is used like the pointer of a dense workspace tensor."""

    def __init__(self, plan, B, dtype, flags):
        self._handle = ctypes.c_void_p()
        check(lib().jt_workspace_sparse_create(plan.handle, B, dtype_code(dtype), flags, ctypes.byref(self._handle)))
        self._ptr = lib().jt_workspace_sparse_ptr(self._handle)
        self.mapped_bytes = lib().jt_workspace_sparse_mapped(self._handle)

    def data_ptr(self):
        return self._ptr

    def close(self):
        if self._handle:
            lib().jt_workspace_sparse_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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


# ---------------------------------------------------------------------------------------------
# host compile phase (jt_compile.cpp): CSR helpers and the three calls


def _csr(lists):
    """(ptr, data) int32 arrays of a list of int lists."""
    ptr = np.zeros(len(lists) + 1, np.int32)
    if lists:
        np.cumsum([len(x) for x in lists], out=ptr[1:])
    data = np.fromiter((v for x in lists for v in x), np.int32, count=int(ptr[-1]))
    return ptr, data


def _p32(arr):
    return arr.ctypes.data_as(_i32p) if arr is not None else None


def _p64(arr):
    return arr.ctypes.data_as(_i64p) if arr is not None else None


def _take_ibuf(handle):
    """Copy the int32 arrays of a ``jt_ibuf`` out and destroy it."""
    try:
        out = []
        for k in range(lib().jt_ibuf_count(handle)):
            n = lib().jt_ibuf_size(handle, k)
            ptr = lib().jt_ibuf_data(handle, k)
            out.append(np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n > 0 else np.zeros(0, np.int32))
        return out
    finally:
        lib().jt_ibuf_destroy(handle)


def _split(ptr, data):
    return [data[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]


def triangulate(var_sizes, factors, order=None):
    """``jt_triangulate`` on integer variables 0..n-1.  Returns ``(cliques, factor_to_clique,
    fill_edges, elimination_order)`` as Python lists."""
    sizes = np.ascontiguousarray(var_sizes, np.int64)
    fptr, fdata = _csr(factors)
    order_arr = np.ascontiguousarray(order, np.int32) if order is not None else None
    handle = ctypes.c_void_p()
    check(lib().jt_triangulate(len(sizes), _p64(sizes), len(factors), _p32(fptr), _p32(fdata), _p32(order_arr),
                               len(order_arr) if order_arr is not None else 0, ctypes.byref(handle)))
    cptr, cvars, f2c, fill, elim = _take_ibuf(handle)
    return _split(cptr, cvars), f2c.tolist(), fill.reshape(-1, 2).tolist(), elim.tolist()


def junction_tree(var_sizes, cliques, root=None):
    """``jt_junction_tree``.  Returns ``(separators, parent, parent_sep, order)``."""
    sizes = np.ascontiguousarray(var_sizes, np.int64)
    cptr, cdata = _csr(cliques)
    handle = ctypes.c_void_p()
    check(lib().jt_junction_tree(len(sizes), _p64(sizes), len(cliques), _p32(cptr), _p32(cdata),
                                 -1 if root is None else int(root), ctypes.byref(handle)))
    sptr, svars, parent, parent_sep, order = _take_ibuf(handle)
    return _split(sptr, svars), parent.tolist(), parent_sep.tolist(), order.tolist()


def plan_build(sizes, full_sizes, n_cliques, node_vars, tree, factors, factor_to_clique, evidence_vars, outputs,
               likelihood_vars=()):
    """``jt_plan_build``: the plan blob (bytes) for integer variables.  ``tree`` is ``None`` or
    ``(order, parent, parent_sep)``; ``factors`` / ``outputs`` may be ``None``."""
    lik = np.ascontiguousarray(likelihood_vars, np.int32)
    sizes = np.ascontiguousarray(sizes, np.int64)
    full = np.ascontiguousarray(full_sizes, np.int64)
    nptr, ndata = _csr(node_vars)
    order = parent = parent_sep = None
    if tree is not None:
        order, parent, parent_sep = (np.ascontiguousarray(x, np.int32) for x in tree)
    fptr = fdata = f2c = None
    n_factors = -1
    if factors is not None:
        n_factors = len(factors)
        fptr, fdata = _csr(factors)
        f2c = np.ascontiguousarray(factor_to_clique, np.int32)
    ev = np.ascontiguousarray(evidence_vars, np.int32)
    optr = odata = None
    n_out = -1
    if outputs is not None:
        n_out = len(outputs)
        optr, odata = _csr(outputs)
    blob, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
    check(lib().jt_plan_build(len(sizes), _p64(sizes), _p64(full), n_cliques, len(node_vars) - n_cliques,
                              _p32(nptr), _p32(ndata), 1 if tree is not None else 0, _p32(order), _p32(parent),
                              _p32(parent_sep), n_factors, _p32(fptr), _p32(fdata), _p32(f2c), len(ev), _p32(ev),
                              n_out, _p32(optr), _p32(odata), len(lik), _p32(lik), ctypes.byref(blob),
                              ctypes.byref(nbytes)))
    try:
        return ctypes.string_at(blob, nbytes.value)
    finally:
        lib().jt_free(blob)
