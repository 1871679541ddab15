"""Sum-product distributive law: the operator plugin of the propagation path.

Mirror of the reference's ``SumProduct`` (``/root/reference/junctiontree/sum_product.py:2-35``):
``SumProduct(einsum, *args, **kwargs).einsum(op0, vars0, op1, vars1, ..., out_vars)`` remaps
arbitrary hashable variable labels to small integers and forwards to the injected einsum
function.  The difference is the default: without an injected function the contraction runs on
the GPU through ``jt_contract`` (``include/jt_b200.h``) -- there is no NumPy path in this
package.  ``project`` and ``absorb`` (the Hugin operator names of BASELINE.json's north_star;
removed from the reference in 0.2.0, ``CHANGELOG.md:15-20``) are additive API on top of the same
device contraction.
"""

import numpy as np

from . import _native
from . import engine as eng
from .schedule import _Space, _row_major_strides


def device_einsum(*args, semiring=_native.JT_SR_SUM_PRODUCT):
    """GPU einsum in NumPy's interleaved form: ``device_einsum(op0, [i, j], op1, [j], [i])``.

    Computes ``out[out_labels] = sum over the other labels of prod_k op_k`` for any number of
    operands; ``semiring`` (a ``JT_SR_*`` flag) replaces sum and product by another
    distributive law (max / product, logaddexp / plus, max / plus).  Labels are hashable; size-1 axes broadcast (as in ``np.einsum``); a label
    repeated inside one operand takes its diagonal.  Every output label must occur in an
    operand.  float32 inputs give a float32 result only if all operands are float32 (the
    reference promotes to float64 through its float64 separators, ``junctiontree.py:311-315``).

    Returns a NumPy array (operands are copied to the device and the result copied back); CUDA
    tensors as operands give a CUDA tensor result.
    """
    t = eng.require_cuda()
    args = list(args)
    if len(args) % 2 == 0:
        # implicit-output form: only meaningful for label-free operands (reference D16)
        args = args + [[]]
    operands, labels, out_labels = args[0:-1:2], [list(l) for l in args[1:-1:2]], list(args[-1])
    if len(set(out_labels)) != len(out_labels):
        raise ValueError("repeated output label")
    any_tensor = any(t.is_tensor(op) for op in operands)
    arrays = [op if t.is_tensor(op) else np.asarray(op) for op in operands]
    if all((a.dtype == t.float32) if t.is_tensor(a) else (a.dtype == np.float32) for a in arrays):
        dtype = np.dtype(np.float32)
    else:
        dtype = np.dtype(np.float64)
    tdt = eng.torch_dtype(dtype)

    sizes = {}
    for a, ls in zip(arrays, labels):
        if a.ndim != len(ls):
            raise ValueError("operand has %d axes but %d labels" % (a.ndim, len(ls)))
        for n, l in zip(a.shape, ls):
            n = int(n)
            if sizes.get(l, 1) == 1:
                sizes[l] = n
            elif n != 1 and n != sizes[l]:
                raise ValueError("operands could not be broadcast together on label %r" % (l,))
    for l in out_labels:
        if l not in sizes:
            raise ValueError("output label %r does not occur in any operand" % (l,))
    in_out = set(out_labels)
    rest = []
    for ls in labels:
        for l in ls:
            if l not in in_out and l not in rest:
                rest.append(l)
    s_space, r_space = _Space(out_labels, sizes), _Space(rest, sizes)

    tables, maps, dev_ops = [], [], []
    cursor = 0
    for a, ls in zip(arrays, labels):
        dev = (a if t.is_tensor(a) else t.from_numpy(np.ascontiguousarray(a)))
        dev = dev.to(device="cuda", dtype=tdt).contiguous()
        dev_ops.append(dev)
        stride_of = {}
        for l, n, st in zip(ls, dev.shape, _row_major_strides(dev.shape)):
            if int(n) != 1:
                stride_of[l] = stride_of.get(l, 0) + st      # repeated label -> diagonal
        quad = []
        for tab in s_space.tables(stride_of) + r_space.tables(stride_of):
            quad.append(cursor)
            tables.append(np.asarray(tab, np.int64))
            cursor += tab.size
        maps.append(quad)
    out = t.empty(tuple(sizes[l] for l in out_labels), dtype=tdt, device="cuda")
    _native.contract([d.data_ptr() for d in dev_ops], np.concatenate(tables), np.asarray(maps),
                     s_space.n, r_space.n, s_space.n_lo, r_space.n_lo, 1, dtype, out.data_ptr(),
                     t.cuda.current_stream().cuda_stream, semiring)
    if any_tensor:
        return out
    return out.cpu().numpy()


class SumProduct():
    ''' Sum-product distributive law '''

    #: semiring of the device kernels (``JT_SR_*``); subclasses in ``semirings.py`` override it
    semiring_flag = _native.JT_SR_SUM_PRODUCT
    name = "sum-product"

    def __init__(self, einsum=None, *args, **kwargs):
        # `einsum` is the plugin hook of the reference (sum_product.py:6-12): any function with
        # np.einsum's interleaved calling convention.  None selects the sm_100a contraction.
        self.on_device = einsum is None
        self.func = einsum if einsum is not None else self._device_einsum
        self.args = args
        self.kwargs = kwargs

    def _device_einsum(self, *args):
        return device_einsum(*args, semiring=self.semiring_flag)

    def einsum(self, *args, **kwargs):
        '''Einstein summation ``einsum(op0, vars0, op1, vars1, ..., out_vars)`` with arbitrary
        hashable variable labels (reference ``sum_product.py:14-35``).'''
        args_list = list(args)
        explicit = len(args_list) % 2 == 1
        label_lists = args_list[1::2] + ([args_list[-1]] if explicit else [])
        var_map = {}
        for labels in (label_lists if explicit else []):
            for var in labels:
                if var not in var_map:
                    var_map[var] = len(var_map)
        # like the reference, the implicit-output form only works for label-free operands
        args_list[1::2] = [[var_map[var] for var in labels] for labels in args_list[1::2]]
        if explicit:
            args_list[-1] = [var_map[var] for var in args_list[-1]]
        return self.func(*args_list, *self.args, **kwargs, **self.kwargs)

    # ---- Hugin operator names (additive API) ----

    def project(self, potential, variables, sep_variables):
        '''Separator projection: sum ``potential`` over the variables not in ``sep_variables``.'''
        return self.einsum(potential, list(variables), list(sep_variables))

    def absorb(self, potential, variables, message, msg_variables, old=None):
        '''Absorption: multiply ``message`` (scope ``msg_variables``) into ``potential``.  With
        ``old`` given the factor is the ratio message/old with 0/0 = 0 (Hugin update).'''
        if old is not None:
            message = self.ratio(message, old)
        return self.einsum(potential, list(variables), message, list(msg_variables), list(variables))

    def ratio(self, new, old):
        '''Elementwise new/old with x/0 = 0 (the Hugin separator ratio), on the device.'''
        t = eng.require_cuda()
        arrays = [x if t.is_tensor(x) else np.asarray(x) for x in (new, old)]
        f32 = all((a.dtype == t.float32) if t.is_tensor(a) else (a.dtype == np.float32) for a in arrays)
        dtype = np.dtype(np.float32 if f32 else np.float64)
        tdt = eng.torch_dtype(dtype)
        dev = [(a if t.is_tensor(a) else t.from_numpy(np.ascontiguousarray(a))).to(device="cuda", dtype=tdt)
               for a in arrays]
        a, b = t.broadcast_tensors(*dev)
        a, b = a.contiguous(), b.contiguous()
        out = t.empty_like(a)
        _native.ratio(a.data_ptr(), b.data_ptr(), out.data_ptr(), out.numel(), dtype,
                      t.cuda.current_stream().cuda_stream)
        if any(t.is_tensor(x) for x in (new, old)):
            return out
        return out.cpu().numpy()
