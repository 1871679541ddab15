"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.

Oracle tier T0: brute-force sum-product.  The joint product of all potentials is formed with one
einsum and marginalised to each requested scope -- the oracle of the reference's own unit tests
(``brute_force_sum_product``, reference ``tests/test_computation.py:19-32``).  Feasible while the
joint state space stays below ~2**26.
"""

import numpy as np

from .ref_fixed import _einsum, slice_evidence


def joint_marginals(arrays, var_lists, scopes, semiring="sum_product"):
    """Marginals of prod_k arrays[k] (axes ``var_lists[k]``) onto each scope in ``scopes``
    (``semiring``: the (+, x) pair, see ``ref_fixed.SEMIRINGS``)."""
    ops = []
    for a, vs in zip(arrays, var_lists):
        ops += [np.asarray(a, dtype=np.float64), list(vs)]
    return [_einsum(*(ops + [list(scope)]), semiring=semiring) for scope in scopes]


def tree_beliefs(tree, node_list, potentials, semiring="sum_product"):
    """Beliefs of every node of a tree given *node* potentials (clique and separator arrays),
    in node-list order.  Same quantity as ``brute_force_sum_product`` of the reference tests,
    which multiplies every node potential (separator potentials are ones there)."""
    ids = []
    stack = [tree]
    while stack:
        sub = stack.pop()
        ids.append(sub[0])
        for s, t in sub[1:]:
            ids.append(s)
            stack.append(t)
    arrays = [potentials[i] for i in ids]
    var_lists = [node_list[i] for i in ids]
    return joint_marginals(arrays, var_lists, node_list, semiring)


def factor_graph_marginals(factors, values, scopes, evidence=None, semiring="sum_product"):
    """Marginals of the factor-graph joint onto ``scopes``; observed variables keep a size-1
    axis (evidence slicing semantics of the reference, ``computation.py:11-34``)."""
    if evidence:
        values = slice_evidence(values, factors, evidence)
    return joint_marginals(values, factors, scopes, semiring)
