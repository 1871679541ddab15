"""Sum-product message passing on a junction tree -- the hot path, on the GPU.

Mirror of the reference module ``/root/reference/junctiontree/computation.py``:

* ``sum_product``       the module-level distributive law (reference ``:9``) -- here the device
                        contraction instead of ``SumProduct(np.einsum)``;
* ``apply_evidence``    evidence slicing (reference ``:11-34``), same quirks;
* ``compute_beliefs``   collect + distribute (reference ``:37-246``).

With the default ``dl`` the whole propagation is compiled once per tree into a level-ordered
schedule (``schedule.py``) and executed by the sm_100a kernels of ``libjt_b200.so``; nothing is
computed with NumPy.  The division trick of the reference (``remove_message``, ``:99-136``) is
replaced by a division-free exclude-one product, which matches it wherever the reference is
correct and is exact where it is not.
"""

import numpy as np

from . import construction as cons
from . import engine as eng
from . import schedule as sch
from .sum_product import SumProduct
from .semirings import max_product, log_sum_exp, max_sum  # noqa: F401  (other laws, same surface)

# The distributive law used by default.  The reference binds np.einsum here
# (computation.py:4-9); this build binds the device contraction.
sum_product = SumProduct()


def apply_evidence(potentials, variables, evidence):
    ''' Shrink potentials based on given evidence

    :param potentials: list of numpy arrays subject to evidence
    :param variables: list of variables in corresponding to potentials
    :param evidence: dictionary with variables as keys and assigned value as value
    :return: a new list of potentials after evidence applied

    Pure indexing (views, bit-exact).  As in the reference (``computation.py:20-34``) every
    element of the result is wrapped in a one-element list and scalars pass through.
    The batched, on-device form of the same slicing is the ``evidence`` argument of
    ``JunctionTree.propagate_batch``.
    '''
    out = []
    for pot, pot_vars in zip(potentials, variables):
        if np.isscalar(pot):
            out.append([pot])
            continue
        index = tuple(
            slice(evidence.get(var, 0), evidence.get(var, pot.shape[i]) + 1)
            for i, var in enumerate(pot_vars)
        )
        out.append([pot[index]])
    return out


_plan_cache = {}
_PLAN_CACHE_MAX = 64


def _tree_key(tree):
    """Hashable form of a nested tree, built iteratively."""
    parts = []
    stack = [tree]
    while stack:
        node = stack.pop()
        parts.append((node[0], tuple(child[0] for child in node[1:]), tuple(child[1][0] for child in node[1:])))
        stack.extend(child[1] for child in node[1:])
    return tuple(parts)


def _engine_for(tree, clique_vars, shapes):
    """Engine for (tree, node variable lists, node shapes), cached."""
    sizes = {}
    for node_vars, shape in zip(clique_vars, shapes):
        if len(node_vars) != len(shape):
            raise ValueError("a potential with shape %s cannot have variables %r" % (shape, node_vars))
        for var, n in zip(node_vars, shape):
            sizes[var] = max(sizes.get(var, 1), int(n))
    try:
        key = (_tree_key(tree), tuple(tuple(v) for v in clique_vars), tuple(sorted(sizes.items(), key=repr)),
               eng.current_device())
        hash(key)
    except TypeError:
        key = None
    if key is not None and key in _plan_cache:
        return _plan_cache[key], sizes
    engine = eng.Engine(sch.Plan(tree, clique_vars, sizes))
    if key is not None:
        if len(_plan_cache) >= _PLAN_CACHE_MAX:
            _plan_cache.pop(next(iter(_plan_cache)))
        _plan_cache[key] = engine
    return engine, sizes


def _compute_beliefs_device(tree, potentials, clique_vars, semiring=0):
    t = eng.require_cuda()
    arrays = [np.asarray(p) for p in potentials]
    for a in arrays:
        if a.dtype.kind not in "f":
            raise TypeError("potentials must be floating point (got %s)" % a.dtype)
    # the reference always ends up in float64 unless everything is float32
    dtype = np.dtype(np.float32) if all(a.dtype == np.float32 for a in arrays) else np.dtype(np.float64)
    engine, sizes = _engine_for(tree, clique_vars, [a.shape for a in arrays])
    plan = engine.plan
    # one library call (jt_beliefs_host): potentials -> the clique rows of the workspace, collect +
    # distribute (a single launch for small trees), clique and separator beliefs -> host.
    # Separator inputs are overwritten by the collect pass, as in the reference (computation.py:92)
    key = ("beliefs", np.dtype(dtype).str)
    stage = engine._workspaces.get(key)
    if stage is None:
        engine.dev.upload()
        tdt = eng.torch_dtype(dtype)
        n_nodes = plan.clique_entries + plan.sep_entries
        host_in = t.zeros(max(plan.clique_entries, 1), dtype=tdt).pin_memory()
        host_out = t.zeros(max(n_nodes, 1), dtype=tdt).pin_memory()
        stage = (host_in, host_out, host_in.numpy(), host_out.numpy(), engine.new_workspace(1, dtype), t.cuda.Stream())
        t.cuda.synchronize()
        engine._workspaces[key] = stage
    host_in, host_out, in_np, out_np, ws, stream = stage
    for c in range(plan.n_cliques):
        full = np.broadcast_to(arrays[c], tuple(plan.node_shape[c]))   # size-1 axes (reference D7)
        in_np[plan.node_off[c]:plan.node_off[c] + plan.node_size[c]] = full.reshape(-1)
    engine.dev.beliefs_host(host_in.data_ptr(), dtype, ws.data_ptr(), host_out.data_ptr(), semiring, stream.cuda_stream)
    flat = out_np.copy()
    return [
        flat[plan.node_off[k]:plan.node_off[k] + plan.node_size[k]].reshape(tuple(plan.node_shape[k]))
        for k in range(len(clique_vars))
    ]


def _compute_beliefs_plugin(tree, potentials, clique_vars, dl):
    """Message passing through a user-supplied distributive law, one ``dl.einsum`` per op.

    Same operator sequence as the reference (``get_message`` E1+E2, ``send_message`` E3-E5) but
    level-ordered, iterative and with the exclude-one product formed directly.  Every number is
    produced by ``dl.einsum``; this function only orchestrates.
    """
    beliefs = list(potentials)
    order, parent, parent_sep, _, children = cons.tree_edges(tree)
    for c in reversed(order):                       # collect: children before parents
        if parent[c] < 0:
            continue
        args = []
        for sep, _ in children[c]:
            args += [beliefs[sep], clique_vars[sep]]
        args += [beliefs[c], clique_vars[c], clique_vars[parent_sep[c]]]
        beliefs[parent_sep[c]] = dl.einsum(*args)
    down = {}
    for c in order:                                 # distribute: parents before children
        incoming = [(beliefs[sep], clique_vars[sep]) for sep, _ in children[c]]
        if parent[c] >= 0:
            incoming.append((down[parent_sep[c]], clique_vars[parent_sep[c]]))
        for i, (sep, _) in enumerate(children[c]):
            args = []
            for j, (msg, msg_vars) in enumerate(incoming):
                if j != i:
                    args += [msg, msg_vars]
            args += [beliefs[c], clique_vars[c], clique_vars[sep]]
            down[sep] = dl.einsum(*args)
            beliefs[sep] = dl.einsum(beliefs[sep], clique_vars[sep], down[sep], clique_vars[sep],
                                     clique_vars[sep])
        args = [beliefs[c], clique_vars[c]]
        for msg, msg_vars in incoming:
            args += [msg, msg_vars]
        beliefs[c] = dl.einsum(*(args + [clique_vars[c]]))
    return beliefs


def compute_beliefs(tree, potentials, clique_vars, dl=sum_product):
    '''Computes beliefs for clique potentials in a junction tree
    using Shafer-Shenoy updates.

    :param tree: list representing the structure of the junction tree
    :param potentials: list of numpy arrays for cliques in junction tree
    :param clique_vars: list of variables included in each clique in potentials list
    :param dl: distributive law; the default runs the compiled schedule on the GPU, a
               ``SumProduct`` built around a user einsum function is called once per operator;
               ``max_product`` / ``log_sum_exp`` / ``max_sum`` run the same schedule in another
               semiring (``semirings.py``)
    :return: list of numpy arrays defining computed beliefs of each clique

    Inputs are never modified (reference ``computation.py:245``); the result lists the clique
    beliefs followed by the separator beliefs in ``clique_vars`` order.
    '''
    if getattr(dl, "on_device", False):
        return _compute_beliefs_device(tree, potentials, clique_vars, getattr(dl, "semiring_flag", 0))
    return _compute_beliefs_plugin(tree, potentials, clique_vars, dl)
