"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.

Oracle tier T2: a float64 NumPy restatement of the reference's sum-product propagation
(`/root/reference/junctiontree/{junctiontree,computation}.py`), op for op, with exactly three
deviations that make it correct outside the reference's narrow valid domain:

1. clique potentials are full-size (no size-1 axes for clique variables that no assigned factor
   covers -- reference defect D7, ``junctiontree.py:52-61``);
2. the exclude-one product of ``send_message`` is formed from the other messages directly
   instead of dividing the full product (reference defects D2/D3, ``computation.py:99-136``);
3. it is run on valid junction trees (the reference construction drops maximal cliques, D8).

Parity pinning: on the reference-valid domain this module is checked against the *unmodified*
reference (tier T1; golden vectors under ``tests/golden/`` were produced by
``tests/golden/make_golden.py`` importing ``/root/reference``) and against brute force
(``oracle/brute.py``, the reference tests' own oracle).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""

import numpy as np

#: distributive laws: name -> (product, reduction over axes, multiplicative identity).
#: "sum_product" is the reference's (np.einsum); the others restate the same contraction with
#: another (+, x) pair by explicit broadcasting.  The log-domain laws are
#: pinned through exact identities: log_sum_exp(log x) = log(sum_product(x)) and
#: max_sum(log x) = log(max_product(x)); max_product is pinned by brute force over the joint.
SEMIRINGS = ("sum_product", "max_product", "log_sum_exp", "max_sum")


def _logsumexp(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore"):
        out = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
    return np.squeeze(out, axis=axis)


def semiring_ops(semiring):
    """(product, reduce(x, axes), one) of a distributive law."""
    if semiring == "sum_product":
        return np.multiply, lambda x, ax: np.sum(x, axis=ax), 1.0
    if semiring == "max_product":
        return np.multiply, lambda x, ax: np.max(x, axis=ax), 1.0
    if semiring == "log_sum_exp":
        return np.add, _logsumexp, 0.0
    if semiring == "max_sum":
        return np.add, lambda x, ax: np.max(x, axis=ax), 0.0
    raise ValueError("unknown semiring %r" % (semiring,))


def _semiring_einsum(semiring, args):
    """The contraction of ``_einsum`` in another semiring: operands are aligned on the union of
    their labels by broadcasting, combined with the product and reduced over the labels that
    are not in the output."""
    mul, reduce_, one = semiring_ops(semiring)
    operands, label_lists, out_labels = args[0:-1:2], args[1:-1:2], list(args[-1])
    order = list(out_labels)
    for labels in label_lists:
        for v in labels:
            if v not in order:
                order.append(v)
    acc = np.asarray(one, dtype=np.float64)
    for op, labels in zip(operands, label_lists):
        op = np.asarray(op, dtype=np.float64)
        perm = sorted(range(len(labels)), key=lambda i: order.index(labels[i]))
        op = np.transpose(op, perm)
        shape = [1] * len(order)
        for axis_len, i in zip(op.shape, perm):
            shape[order.index(labels[i])] = axis_len
        acc = mul(acc, op.reshape(shape))
    acc = np.broadcast_to(np.asarray(acc, dtype=np.float64), np.shape(acc) if np.ndim(acc) == len(order)
                          else [1] * len(order))
    rest = tuple(range(len(out_labels), len(order)))
    return reduce_(acc, rest) if rest else np.array(acc)


def _einsum(*args, semiring="sum_product"):
    """``np.einsum`` in interleaved form with arbitrary hashable labels
    (what ``SumProduct.einsum`` does, reference ``sum_product.py:14-35``)."""
    args = list(args)
    if semiring != "sum_product":
        return _semiring_einsum(semiring, args)
    label_lists = args[1::2] + [args[-1]]
    var_map = {}
    for labels in label_lists:
        for v in labels:
            if v not in var_map:
                var_map[v] = len(var_map)
    args[1::2] = [[var_map[v] for v in labels] for labels in args[1::2]]
    args[-1] = [var_map[v] for v in args[-1]]
    return np.einsum(*args)


def evaluate(factors, values, maxcliques, factor_to_maxclique, sizes, semiring="sum_product"):
    """psi_C = product of the assigned factors, broadcast to the full clique shape.

    Follows ``CliqueGraph.evaluate`` (reference ``junctiontree.py:203-226``) and its einsum
    helper (``:34-80``); deviation 1: uncovered variables get their full size, and a clique
    without factors is all ones.
    """
    out = []
    for c, cvars in enumerate(maxcliques):
        fs = [f for f, home in enumerate(factor_to_maxclique) if home == c]
        shape = tuple(int(sizes[v]) for v in cvars)
        args = []
        for f in fs:
            args += [np.asarray(values[f], dtype=np.float64), list(factors[f])]
        args += [np.full(shape, semiring_ops(semiring)[2]), list(cvars), list(cvars)]
        out.append(_einsum(*args, semiring=semiring))
    return out


def compute_beliefs(tree, potentials, clique_vars, semiring="sum_product"):
    """Shafer-Shenoy collect + distribute, returns beliefs in node order.

    Follows ``compute_beliefs`` (reference ``computation.py:37-246``): ``get_message``
    (``:47-96``) for the collect pass, ``send_message`` (``:140-224``) for the distribute pass,
    separator slots hold the up-message after collect (``:92``) and up*down after distribute
    (``:210``).  Iterative instead of recursive (defect D15).
    """
    beliefs = [np.array(p, dtype=np.float64) for p in potentials]
    mul = semiring_ops(semiring)[0]

    def _einsum(*args):
        return globals()["_einsum"](*args, semiring=semiring)

    # flatten: pre-order list of (clique, parent_sep, [(sep, child) ...])
    nodes = []
    stack = [(tree, None)]
    while stack:
        sub, psep = stack.pop()
        kids = [(s, t[0]) for s, t in sub[1:]]
        nodes.append((sub[0], psep, kids))
        for s, t in reversed(sub[1:]):
            stack.append((t, s))

    # collect: children before parents  (get_message, computation.py:47-96)
    for c, psep, kids in reversed(nodes):
        if psep is None:
            continue  # the root's collect (Z) is discarded by the reference (:90-96)
        args = []
        for s, _ in kids:
            args += [beliefs[s], clique_vars[s]]
        args += [beliefs[c], clique_vars[c], clique_vars[psep]]
        beliefs[psep] = _einsum(*args)                       # E1 + E2

    # distribute: parents before children  (send_message, computation.py:140-224)
    down = {}
    for c, psep, kids in nodes:
        incoming = [(beliefs[s], clique_vars[s]) for s, _ in kids]
        if psep is not None:
            incoming.append((down[psep], clique_vars[psep]))
        for i, (s, _) in enumerate(kids):
            args = []
            for j, (m, mv) in enumerate(incoming):
                if j != i:
                    args += [m, mv]                           # exclude-one product (deviation 2)
            args += [beliefs[c], clique_vars[c], clique_vars[s]]
            message = _einsum(*args)                          # E4
            down[s] = message
            beliefs[s] = mul(beliefs[s], message)             # M1 (:210)
        args = [beliefs[c], clique_vars[c]]
        for m, mv in incoming:
            args += [m, mv]
        args += [clique_vars[c]]
        beliefs[c] = _einsum(*args)                           # E5 (:216-224)
    return beliefs


def marginalize(factors, maxcliques, factor_to_maxclique, ys, semiring="sum_product"):
    """Per-factor output = clique belief summed to the factor scope, axes in factor order
    (``CliqueGraph.marginalize``, reference ``junctiontree.py:229-274``)."""
    return [
        _einsum(ys[home], list(maxcliques[home]), list(fv), semiring=semiring)
        for fv, home in zip(factors, factor_to_maxclique)
    ]


def propagate(tree, separators, maxcliques, factor_to_maxclique, factors, sizes, values,
              semiring="sum_product"):
    """``JunctionTree.propagate`` (reference ``junctiontree.py:297-331``)."""
    psi = evaluate(factors, values, maxcliques, factor_to_maxclique, sizes, semiring)
    one = semiring_ops(semiring)[2]
    seps = [np.full(tuple(int(sizes[v]) for v in s), one) for s in separators]      # :311-315
    ys = compute_beliefs(tree, psi + seps, list(maxcliques) + list(separators), semiring)
    return marginalize(factors, maxcliques, factor_to_maxclique, ys, semiring), ys


def slice_evidence(values, factors, evidence):
    """Evidence slicing ``pot[e:e+1]`` on observed axes
    (``apply_evidence``, reference ``computation.py:11-34``, without the list wrapping D4)."""
    out = []
    for pot, fv in zip(values, factors):
        ix = tuple(slice(evidence[v], evidence[v] + 1) if v in evidence else slice(None) for v in fv)
        out.append(np.asarray(pot)[ix])
    return out


def with_likelihood_factors(factors, factor_to_maxclique, maxcliques, values, likelihoods, b):
    """Soft evidence in the reference's own terms: a likelihood vector on variable v is one more
    factor ``[v]`` of the factor graph (assigned to any clique containing v).  Returns the
    extended ``(factors, factor_to_maxclique, values)`` for instance ``b``."""
    factors, f2c, values = list(factors), list(factor_to_maxclique), list(values)
    for v, lam in likelihoods.items():
        factors.append([v])
        f2c.append(next(c for c, cv in enumerate(maxcliques) if v in cv))
        values.append(np.asarray(lam[b], np.float64))
    return factors, f2c, values


def propagate_batch(tree, separators, maxcliques, factor_to_maxclique, factors, sizes, values,
                    evidence_vars=(), evidence=None, n=None, semiring="sum_product"):
    """Loop of independent propagations, one per evidence row.

    :param sizes: full sizes; observed variables are sliced to size 1 per instance
    :param evidence: int array ``[B, len(evidence_vars)]``
    :return: (factor outputs ``[B, *shape]`` per factor, node beliefs ``[B, *shape]`` per node)
    """
    evidence_vars = list(evidence_vars)
    B = n if n is not None else (len(evidence) if evidence is not None else 1)
    eff = dict(sizes)
    for v in evidence_vars:
        eff[v] = 1
    outs, nodes = None, None
    for b in range(B):
        ev = {v: int(evidence[b][i]) for i, v in enumerate(evidence_vars)}
        vals = slice_evidence(values, factors, ev)
        fo, ys = propagate(tree, separators, maxcliques, factor_to_maxclique, factors, eff, vals, semiring)
        if outs is None:
            outs = [np.empty((B,) + o.shape) for o in fo]
            nodes = [np.empty((B,) + y.shape) for y in ys]
        for k, o in enumerate(fo):
            outs[k][b] = o
        for k, y in enumerate(ys):
            nodes[k][b] = y
    return outs, nodes
