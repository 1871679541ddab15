"""Host-side compilation: triangulation, maximal cliques, junction tree.

This is the compile phase that stays on the host (BASELINE.json north_star).  It keeps the
call surface of the reference module (``/root/reference/junctiontree/construction.py``):

* ``find_triangulation(factors, var_sizes)``      reference ``construction.py:176-353``
* ``construct_junction_tree(cliques, var_sizes)`` reference ``construction.py:522-578``
* traversal helpers ``bf_traverse`` / ``df_traverse`` / ``get_clique`` /
  ``generate_potential_pairs`` / ``find_subtree`` (reference ``:6-36, :431-519, :604-640``)

and the same data formats (``README.md:52-70`` of the reference): a tree is
``[clique_ix, (sep_ix, subtree), ...]`` and ``node_list = maxcliques + separators``.

It is *not* a transcription.  The reference's construction is only valid on a narrow input
domain (D1 heap ties with int labels, D8 dropped maximal cliques, D9 stale
heap scores, D5 ``raise StopIteration`` in generators).  Here:

* elimination is min-fill on the *current* graph, ties broken by cluster weight and then by
  label rank, with an incrementally maintained lazy heap (milliseconds at 500 variables);
* every elimination cluster that is not contained in an earlier one is a maximal clique, so the
  running-intersection property always holds;
* separators carry an explicit, deterministic axis order (label rank), independent of
  ``PYTHONHASHSEED`` (the reference uses ``tuple(set(..))``, ``construction.py:538``);
* the tree is rooted at its centre so the level-ordered schedule is as shallow as possible;
* an explicit elimination ``order`` may be supplied (needed for grid models).

Two implementations of the same algorithms: C++ (``csrc/jt_compile.cpp`` behind
``jt_triangulate`` / ``jt_junction_tree`` of the C ABI; the default -- 500 variables compile in
about a millisecond) and the pure-Python ones below (``impl="python"``, or
``JT_HOST_COMPILE=python`` in the environment), kept as the cross-check: both produce the same
cliques, the same assignment of factors and the same tree (``tests/test_native_compile.py``).
"""

import heapq
import os
from collections import deque


def _use_native(impl):
    impl = impl or os.environ.get("JT_HOST_COMPILE", "native")
    if impl not in ("native", "python"):
        raise ValueError("impl must be 'native' or 'python'")
    return impl == "native"


# ---------------------------------------------------------------------------------------------
# label utilities


def _label_ranks(labels):
    """Deterministic total order over arbitrary hashable labels.

    Sorted order when the labels are mutually comparable, first-appearance order otherwise.
    """
    labels = list(labels)
    try:
        ordered = sorted(labels)
    except TypeError:
        ordered = labels
    return {v: i for i, v in enumerate(ordered)}


def _used_variables(factors):
    seen = {}
    for factor in factors:
        for var in factor:
            if var not in seen:
                seen[var] = len(seen)
    return list(seen)


def factors_to_undirected_graph(factors):
    """Edges of the moral graph, ``{frozenset((a, b)): {factor ids}}``.

    Same output format as the reference (``construction.py:39-55``).
    """
    edges = {}
    for fix, factor in enumerate(factors):
        for i, a in enumerate(factor):
            for b in factor[i + 1:]:
                edges.setdefault(frozenset((a, b)), set()).add(fix)
    return edges


# ---------------------------------------------------------------------------------------------
# triangulation


def _fill_and_weight(var, adj, sizes):
    nbrs = list(adj[var])
    fill = 0
    for i, a in enumerate(nbrs):
        adj_a = adj[a]
        for b in nbrs[i + 1:]:
            if b not in adj_a:
                fill += 1
    weight = sizes[var]
    for n in nbrs:
        weight *= sizes[n]
    return fill, weight


def elimination_clusters(factors, var_sizes, order=None):
    """Eliminate every variable and return ``(order, clusters, fill_edges)``.

    ``clusters[i]`` is the set {order[i]} | remaining neighbours at the time of elimination.
    With ``order=None`` the order is min-fill, ties by induced cluster weight (product of the
    sizes of the cluster, the reference's intended key ``construction.py:106-108``), then by
    label rank.  Scores are always those of the up-to-date graph.
    """
    variables = _used_variables(factors)
    rank = _label_ranks(variables)
    sizes = {v: int(var_sizes[v]) for v in variables}
    adj = {v: set() for v in variables}
    for factor in factors:
        for i, a in enumerate(factor):
            for b in factor[i + 1:]:
                if a != b:
                    adj[a].add(b)
                    adj[b].add(a)

    out_order, clusters, fill_edges = [], [], []

    def eliminate(var):
        nbrs = sorted(adj[var], key=rank.__getitem__)
        for i, a in enumerate(nbrs):
            for b in nbrs[i + 1:]:
                if b not in adj[a]:
                    adj[a].add(b)
                    adj[b].add(a)
                    fill_edges.append((a, b))
        for n in nbrs:
            adj[n].discard(var)
        del adj[var]
        out_order.append(var)
        clusters.append([var] + nbrs)
        return nbrs

    if order is not None:
        order = list(order)
        if set(order) != set(variables) or len(order) != len(variables):
            raise ValueError("order must be a permutation of the variables used by the factors")
        for var in order:
            eliminate(var)
        return out_order, clusters, fill_edges

    version = {v: 0 for v in variables}
    heap = []
    for v in variables:
        fill, weight = _fill_and_weight(v, adj, sizes)
        heap.append((fill, weight, rank[v], 0, v))
    heapq.heapify(heap)

    while adj:
        fill, weight, _, ver, var = heapq.heappop(heap)
        if var not in adj or ver != version[var]:
            continue  # stale entry
        nbrs = eliminate(var)
        # scores can only change for the neighbours and for their neighbours
        touched = set(nbrs)
        for n in nbrs:
            touched |= adj[n]
        for t in touched:
            version[t] += 1
            f, w = _fill_and_weight(t, adj, sizes)
            heapq.heappush(heap, (f, w, rank[t], version[t], t))

    return out_order, clusters, fill_edges


def find_triangulation(factors, var_sizes, order=None, impl=None):
    """Triangulate the factor graph.

    :param factors: list of factors, each a list of variable labels
    :param var_sizes: ``{label: size}``; entries for unused variables are ignored
    :param order: optional elimination order (permutation of the used variables)
    :param impl: ``"native"`` (C++, default) or ``"python"``
    :return: ``(tri, max_cliques, factor_to_maxclique)`` as the reference
             (``construction.py:176-197``): fill-in edges, maximal cliques (each a list of
             labels sorted by label rank, cf. ``construction.py:347``) and, per factor, the index
             of a maximal clique containing it.
    """
    if _use_native(impl):
        return _find_triangulation_native(factors, var_sizes, order)
    variables = _used_variables(factors)
    rank = _label_ranks(variables)
    order_out, clusters, tri = elimination_clusters(factors, var_sizes, order)

    max_cliques = []      # list of sorted label lists
    clique_sets = []
    cliques_of = {v: [] for v in variables}   # var -> indices of kept cliques containing it
    for var, cluster in zip(order_out, clusters):
        cset = frozenset(cluster)
        # only an earlier cluster that contains `var` can contain this one
        if any(cset <= clique_sets[ix] for ix in cliques_of[var]):
            continue
        ix = len(max_cliques)
        max_cliques.append(sorted(cluster, key=rank.__getitem__))
        clique_sets.append(cset)
        for v in cluster:
            cliques_of[v].append(ix)

    if not max_cliques:
        # only scalar factors (or no factors): one empty clique holds them all
        max_cliques = [[]]
        clique_sets = [frozenset()]

    factor_to_maxclique = []
    for factor in factors:
        fset = set(factor)
        if not fset:
            factor_to_maxclique.append(0)
            continue
        candidates = cliques_of[factor[0]]
        home = next(ix for ix in candidates if fset <= clique_sets[ix])
        factor_to_maxclique.append(home)

    return tri, max_cliques, factor_to_maxclique


def _find_triangulation_native(factors, var_sizes, order):
    """``jt_triangulate``: labels are numbered by rank, the C++ side works on the integers."""
    from . import _native
    variables = _used_variables(factors)
    rank = _label_ranks(variables)
    labels = [None] * len(variables)
    for v, r in rank.items():
        labels[r] = v
    order_ids = None
    if order is not None:
        order = list(order)
        if len(order) != len(variables) or any(v not in rank for v in order) or len(set(order)) != len(order):
            raise ValueError("order must be a permutation of the variables used by the factors")
        order_ids = [rank[v] for v in order]
    cliques, f2c, fill, _ = _native.triangulate([int(var_sizes[v]) for v in labels],
                                                [[rank[v] for v in f] for f in factors], order_ids)
    tri = [(labels[a], labels[b]) for a, b in fill]
    return tri, [[labels[v] for v in c] for c in cliques], f2c


# ---------------------------------------------------------------------------------------------
# junction tree


def _clique_weight(clique, var_sizes):
    w = 1
    for v in clique:
        w *= int(var_sizes[v])
    return w


def construct_junction_tree(cliques, var_sizes, root=None, impl=None):
    """Maximum-weight spanning tree over the clique graph, in the reference's nested format.

    :param cliques: list of maximal cliques (lists of labels)
    :param var_sizes: ``{label: size}``
    :param root: optional clique index to root the tree at (default: the tree centre)
    :param impl: ``"native"`` (C++, default) or ``"python"``
    :return: ``(tree, separators)`` -- cf. reference ``construction.py:522-578``.  Separator ``k``
             is node ``len(cliques) + k``.  Empty separators join unconnected components
             (reference ``construction.py:530``).
    """
    n = len(cliques)
    if n == 0:
        return [], []
    if _use_native(impl):
        return _construct_junction_tree_native(cliques, var_sizes, root)
    all_vars = []
    seen = set()
    for c in cliques:
        for v in c:
            if v not in seen:
                seen.add(v)
                all_vars.append(v)
    rank = _label_ranks(all_vars)
    csets = [frozenset(c) for c in cliques]
    weights = [_clique_weight(c, var_sizes) for c in cliques]

    # candidate edges: larger separators first, then lighter clique pairs, then pair index --
    # the reference's heap key is [1/(|S|+.001), w1+w2, i]  (construction.py:594-598)
    candidates = []
    members = {}
    for ix, c in enumerate(cliques):
        for v in c:
            members.setdefault(v, []).append(ix)
    shared = {}
    for v, ixs in members.items():
        for a_pos, a in enumerate(ixs):
            for b in ixs[a_pos + 1:]:
                shared[(a, b)] = shared.get((a, b), 0) + 1
    for (a, b), k in shared.items():
        candidates.append((-k, weights[a] + weights[b], a, b))
    candidates.sort()

    parent_uf = list(range(n))

    def find(x):
        while parent_uf[x] != x:
            parent_uf[x] = parent_uf[parent_uf[x]]
            x = parent_uf[x]
        return x

    nbrs = [[] for _ in range(n)]
    n_edges = 0
    for _, _, a, b in candidates:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent_uf[ra] = rb
            nbrs[a].append(b)
            nbrs[b].append(a)
            n_edges += 1
    if n_edges < n - 1:
        # unconnected components: join them through empty separators
        comps = {}
        for ix in range(n):
            comps.setdefault(find(ix), ix)
        reps = sorted(comps.values())
        for a, b in zip(reps, reps[1:]):
            parent_uf[find(a)] = find(b)
            nbrs[a].append(b)
            nbrs[b].append(a)

    def bfs(src):
        dist = {src: 0}
        prev = {src: None}
        q = deque([src])
        last = src
        while q:
            u = q.popleft()
            last = u
            for w in sorted(nbrs[u]):
                if w not in dist:
                    dist[w] = dist[u] + 1
                    prev[w] = u
                    q.append(w)
        return last, dist, prev

    if root is None:
        # centre of the tree: middle of a longest path (minimises the number of levels);
        # of the (at most two) centres take the heavier clique
        end_a, _, _ = bfs(0)
        end_b, dist, prev = bfs(end_a)
        path = [end_b]
        while prev[path[-1]] is not None:
            path.append(prev[path[-1]])
        mid = (len(path) - 1) // 2
        centres = {path[mid], path[len(path) - 1 - mid]}
        root = max(sorted(centres), key=lambda c: weights[c])

    # orient away from the root, number the separators in breadth-first order
    order = []
    children = [[] for _ in range(n)]
    separators = []
    visited = [False] * n
    visited[root] = True
    q = deque([root])
    while q:
        u = q.popleft()
        order.append(u)
        for w in sorted(nbrs[u]):
            if not visited[w]:
                visited[w] = True
                sep = sorted(csets[u] & csets[w], key=rank.__getitem__)
                children[u].append((n + len(separators), w))
                separators.append(sep)
                q.append(w)

    subtree = [None] * n
    for u in reversed(order):
        subtree[u] = [u] + [(s_ix, subtree[w]) for s_ix, w in children[u]]
    return subtree[root], separators


def _construct_junction_tree_native(cliques, var_sizes, root):
    """``jt_junction_tree`` and conversion of its arrays to the nested tree format."""
    from . import _native
    all_vars = _used_variables(cliques)
    rank = _label_ranks(all_vars)
    labels = [None] * len(all_vars)
    for v, r in rank.items():
        labels[r] = v
    n = len(cliques)
    if root is not None and not 0 <= int(root) < n:
        raise ValueError("root must be a clique index")
    seps, parent, parent_sep, order = _native.junction_tree([int(var_sizes[v]) for v in labels],
                                                            [[rank[v] for v in c] for c in cliques], root)
    subtree = [None] * n
    for u in order:
        subtree[u] = [u]
    for w in order[1:]:                      # parents precede children: append in order of appearance
        subtree[parent[w]].append((parent_sep[w], subtree[w]))
    return subtree[order[0]], [[labels[v] for v in s] for s in seps]


# ---------------------------------------------------------------------------------------------
# tree utilities (iterative; the reference's recursive versions are construction.py:6-36,
# 431-519, 604-640)


def yield_id(tree):
    yield tree[0]


def yield_clique_pairs(tree):
    for child in tree[1:]:
        yield (tree[0], child[0])


def bf_traverse(tree, clique_ix=None, func=yield_id):
    """Breadth-first traversal over the nodes (cliques and separators) of ``tree``.

    Stops after the node ``clique_ix`` when given (reference ``construction.py:459-477``; the
    reference raises ``StopIteration`` inside the generator there, which is a RuntimeError
    since PEP 479 -- here the generator simply returns).
    """
    queue = deque([tree])
    while queue:
        node = queue.popleft()
        yield from func(node)
        if node[0] == clique_ix:
            return
        queue.extend(node[1:])


def df_traverse(tree, clique_ix=None, func=yield_id):
    """Depth-first (pre-order) traversal, cf. reference ``construction.py:501-519``."""
    stack = [tree]
    while stack:
        node = stack.pop()
        yield from func(node)
        if node[0] == clique_ix:
            return
        stack.extend(reversed(node[1:]))


def generate_potential_pairs(tree):
    """``[(clique_id, child_separator_id), ...]`` in breadth-first order
    (reference ``construction.py:624-640``)."""
    return list(bf_traverse(tree, func=yield_clique_pairs))


def get_clique_vars(clique_vars, clique_ix):
    return clique_vars[clique_ix] if len(clique_vars) > clique_ix else None


def get_clique(tree, node_list, var_label):
    """First node (clique or separator) in depth-first order containing ``var_label``
    (reference ``construction.py:6-36``)."""
    for ix in df_traverse(tree):
        if var_label in node_list[ix]:
            return ix, node_list[ix]
    return None


def find_subtree(tree, clique_ix):
    """True when a clique with id ``clique_ix`` is in ``tree`` (reference ``:604-621``)."""
    stack = [tree]
    while stack:
        node = stack.pop()
        if node[0] == clique_ix:
            return True
        stack.extend(child[1] for child in node[1:])
    return False


def tree_edges(tree):
    """Flatten a nested tree into ``(order, parent, parent_sep, depth, children)``.

    ``order`` lists clique ids breadth-first; ``parent[c]`` / ``parent_sep[c]`` are ``-1`` for
    the root; ``children[c]`` is a list of ``(sep_ix, child_clique)``.  Iterative, so chains
    deeper than Python's recursion limit are fine (reference defect D15).
    """
    order, parent, parent_sep, depth, children = [], {}, {}, {}, {}
    if not tree:
        return order, parent, parent_sep, depth, children
    root = tree[0]
    parent[root], parent_sep[root], depth[root] = -1, -1, 0
    queue = deque([tree])
    while queue:
        node = queue.popleft()
        c = node[0]
        order.append(c)
        children[c] = []
        for sep_ix, sub in node[1:]:
            child = sub[0]
            children[c].append((sep_ix, child))
            parent[child], parent_sep[child], depth[child] = c, sep_ix, depth[c] + 1
            queue.append(sub)
    return order, parent, parent_sep, depth, children


def check_running_intersection(tree, node_list):
    """True when for every variable the cliques containing it form a connected subtree and
    every separator equals the intersection of its two cliques."""
    order, parent, parent_sep, _, _ = tree_edges(tree)
    for c in order:
        p = parent[c]
        if p >= 0 and set(node_list[parent_sep[c]]) != set(node_list[c]) & set(node_list[p]):
            return False
    variables = set(v for c in order for v in node_list[c])
    for v in variables:
        holders = [c for c in order if v in node_list[c]]
        # connected iff exactly one holder has no holder parent
        tops = sum(1 for c in holders if parent[c] < 0 or v not in node_list[parent[c]])
        if tops != 1:
            return False
    return True
