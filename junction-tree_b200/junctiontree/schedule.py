"""Static, level-ordered message schedule ("plan") for the sm_100a propagation kernels.

The plan is everything the device needs that does not depend on the numbers: for every
junction-tree edge the axis maps between clique and separator index spaces, the order in which
messages are computed (collect bottom-up by level, distribute top-down by level, independent
subtrees of one level fused into one launch), the evidence maps for clique initialisation and
the maps that marginalise clique beliefs back to factor scopes.

What it replaces in the reference (all recomputed there on every call, in Python):

* ``junctiontree.py:203-226``  clique <- factors map and the per-clique einsum subscripts (E0)
* ``computation.py:47-96``     collect recursion, per-edge einsum subscripts (E1, E2)
* ``computation.py:140-224``   distribute recursion, exclude-one product, subscripts (E3-E5, D1, M1)
* ``junctiontree.py:229-274``  marginalisation subscripts (E6)
* ``computation.py:11-34``     evidence slicing (V1) -> integer strides per (factor, observed variable)

Everything is expressed through one primitive, the *projection task*::

    term(s, r) = src[ S(s) + R(r) ] * prod_j msg_j[ A_j(s) + B_j(r) ]
    acc(s)     = sum_r term(s, r)
    sm(s)      = prod_j smsg_j[ A_j(s) ]                (messages that do not depend on r)
    out[s]     = acc(s) * sm(s)                         (optional)
    bel[s]     = out[s] * own[s]                        (optional: separator belief = up * down)
    beta[S(s)+R(r)] = term(s, r) * sm(s) * own[s]       (optional: clique belief, in place)

``s`` enumerates the output (separator / factor scope) index space in row-major order, ``r``
the remaining clique axes.  Every map is additive over axes, so it is stored as two small int32
tables (leading axes / trailing axes of the index space): ``S(s) = hi[s // n_lo] + lo[s % n_lo]``.
Identical tables are stored once.

All device buffers use the batch-innermost layout ``[entry, B]``: the entry offsets in the plan
are multiplied by the batch size on the device.
"""

import os

import numpy as np

from . import construction as cons

MAGIC = 0x324E4C5042544A  # "JTBPLN2"
VERSION = 9

# header word indices (int64 words); mirrored in include/jt_b200.h
H_MAGIC, H_VERSION, H_NCLIQUES, H_NSEPS, H_NFACTORS, H_NEVID, H_CLIQUE_ENTRIES, H_SEP_ENTRIES, \
    H_FIN_ENTRIES, H_FOUT_ENTRIES, H_NTAB, H_NTASKS, H_NMSGS, H_NLAUNCHES, H_MAXDEPTH, \
    H_NEVF, H_ROOT_ENTRIES, H_UNI_ENTRIES, H_NOUT, H_LIK_ENTRIES, H_WORDS = range(21)

TASK_WORDS = 24
(T_KIND, T_SRC, T_OUT, T_BETA, T_BEL, T_OWN, T_NS, T_NR, T_NSLO, T_NRLO, T_SRC_SHI, T_SRC_SLO,
 T_SRC_RHI, T_SRC_RLO, T_RMSG_BEGIN, T_RMSG_END, T_SMSG_BEGIN, T_SMSG_END, T_OUT_SPACE, T_NODE,
 T_AUX, T_FLAGS) = range(22)

MSG_WORDS = 8
M_OFF, M_AHI, M_ALO, M_BHI, M_BLO, M_FID, M_UNI = range(7)

# task flags (honoured only when the run time enables uniform mode)
TF_SRC_UNIFORM = 1     # src is the potential of a clique no evidence touches: read from the uniform workspace
TF_OWN_UNIFORM = 2     # the task's own up-message is uniform
TF_TASK_UNIFORM = 4    # every input is uniform: the task runs once, in the uniform workspace

LAUNCH_WORDS = 4
L_PHASE, L_BEGIN, L_END, L_LEVEL = range(4)

KIND_PROJECT, KIND_INIT = 0, 1
(PHASE_INIT, PHASE_COLLECT, PHASE_DIST_PRE, PHASE_DIST_MAIN, PHASE_MARGINAL,
 PHASE_INIT_UNIFORM, PHASE_INIT_INSTANCE, PHASE_COLLECT_UNIFORM, PHASE_COLLECT_INSTANCE,
 PHASE_DIST_MAIN_MESSAGES, PHASE_MARGINAL_DIRECT, PHASE_DIST_UNIFORM, PHASE_DIST_PRE_INSTANCE) = range(13)
SPACE_WORK, SPACE_FOUT = 0, 1

#: trailing-axes table is grown while its length stays within this bound
LO_TABLE_MAX = 1024


def _prod(xs):
    p = 1
    for x in xs:
        p *= int(x)
    return p


def _row_major_strides(shape):
    strides = [0] * len(shape)
    acc = 1
    for i in range(len(shape) - 1, -1, -1):
        strides[i] = acc
        acc *= int(shape[i])
    return strides


def _split_point(shape):
    """Number of trailing axes that go into the ``lo`` table."""
    k, acc = 0, 1
    for n in reversed(shape):
        if k > 0 and acc * int(n) > LO_TABLE_MAX:
            break
        acc *= int(n)
        k += 1
    return k


class _TableArena:
    """int32 tables, stored once per distinct content."""

    def __init__(self):
        self.chunks = []
        self.size = 0
        self.index = {}

    def add(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int64)
        if arr.size and (arr.min() < 0 or arr.max() >= 2 ** 31):
            raise OverflowError("index table entry does not fit int32")
        arr = arr.astype(np.int32)
        key = arr.tobytes()
        hit = self.index.get(key)
        if hit is not None:
            return hit
        off = self.size
        self.index[key] = off
        self.chunks.append(arr)
        self.size += arr.size
        return off

    def array(self):
        if not self.chunks:
            return np.zeros(0, np.int32)
        return np.concatenate(self.chunks)


class _Space:
    """An index space: ordered variables with sizes, split into leading / trailing axes."""

    def __init__(self, variables, sizes):
        self.vars = list(variables)
        self.shape = [int(sizes[v]) for v in self.vars]
        self.n = _prod(self.shape)
        k = _split_point(self.shape)
        self.hi_vars, self.lo_vars = self.vars[:len(self.vars) - k], self.vars[len(self.vars) - k:]
        self.hi_shape, self.lo_shape = self.shape[:len(self.vars) - k], self.shape[len(self.vars) - k:]
        self.n_lo = _prod(self.lo_shape)
        self.n_hi = _prod(self.hi_shape)

    @staticmethod
    def _table(variables, shape, stride_of):
        n = _prod(shape)
        if not variables:
            return np.zeros(1, np.int64)
        tab = np.zeros(shape, np.int64)
        for ax, (v, sz) in enumerate(zip(variables, shape)):
            st = stride_of.get(v, 0)
            if st:
                idx = np.arange(sz, dtype=np.int64) * st
                view = [1] * len(shape)
                view[ax] = sz
                tab = tab + idx.reshape(view)
        return np.broadcast_to(tab, shape).reshape(n)

    def tables(self, stride_of):
        """(hi, lo) tables of the additive map x -> sum_v digit_v(x) * stride_of[v]."""
        return (self._table(self.hi_vars, self.hi_shape, stride_of),
                self._table(self.lo_vars, self.lo_shape, stride_of))

    def touches(self, variables):
        return any(v in variables for v in self.vars)


class Plan:
    """Compiled schedule.  See the module docstring for the task semantics.

    :param tree: nested tree ``[clique, (sep, subtree), ...]`` (reference ``README.md:52-70``),
                 or ``None`` for a bare clique graph (init and marginal stages only)
    :param node_vars: ``maxcliques + separators`` variable lists (reference
                      ``junctiontree.py:317-323``); axis order of every node is taken from here
    :param sizes: ``{var: size}`` effective sizes (1 for observed variables)
    :param factors: optional list of factor variable lists (enables init / marginal stages)
    :param factor_to_clique: clique index per factor
    :param evidence_vars: ordered list of variables observed per instance; their effective
                          size must be 1 and ``full_sizes`` gives the size of the factor axis
    :param full_sizes: ``{var: size}`` of the factor tables as stored (defaults to ``sizes``)
    :param outputs: scopes (variable lists) the marginal stage sums the clique beliefs down to;
                    default: the factor scopes, as ``CliqueGraph.marginalize`` of the reference
    :param likelihood_vars: variables that receive soft evidence: a per-instance likelihood
                    vector ``lambda_v[x_v]`` multiplied into the potential of a clique containing
                    ``v`` (the same as one more single-variable factor ``[v]`` per instance).  The
                    vectors live in the workspace, region ``likelihoods`` ``[lik_entries][B]`` after
                    the down-messages, slot ``k`` at entry ``lik_base + lik_off[k]``; the caller
                    fills them before ``jt_init``
    :param emitter: ``"native"`` (default): tasks and index tables are emitted by the C++ host
                    compile (``jt_plan_build``, ``csrc/jt_compile.cpp``), tens of times faster on
                    large trees; ``"python"``: by the ``_build_*`` methods below, the cross-check
                    (both give byte-identical blobs).  ``JT_PLAN_EMITTER`` overrides the default.
    """

    def __init__(self, tree, node_vars, sizes, factors=None, factor_to_clique=None,
                 evidence_vars=(), full_sizes=None, outputs=None, emitter=None, likelihood_vars=()):
        self.emitter = emitter or os.environ.get("JT_PLAN_EMITTER", "native")
        if self.emitter not in ("native", "python"):
            raise ValueError("emitter must be 'native' or 'python'")
        self.tree = tree
        self.node_vars = [list(v) for v in node_vars]
        self.sizes = dict(sizes)
        self.factors = None if factors is None else [list(f) for f in factors]
        self.factor_to_clique = None if factor_to_clique is None else list(factor_to_clique)
        self.evidence_vars = list(evidence_vars)
        self.full_sizes = dict(full_sizes) if full_sizes is not None else dict(sizes)
        self.outputs = None if outputs is None else [list(o) for o in outputs]
        self.likelihood_vars = list(likelihood_vars)
        if self.likelihood_vars and (self.factors is None or tree is None):
            raise ValueError("soft evidence needs the factor graph and the tree")
        if len(set(self.likelihood_vars)) != len(self.likelihood_vars):
            raise ValueError("duplicate variable in likelihood_vars")
        for v in self.likelihood_vars:
            if v in self.evidence_vars:
                raise ValueError("variable %r is observed: it cannot also carry a likelihood" % (v,))
        for v in self.evidence_vars:
            if int(self.sizes[v]) != 1:
                raise ValueError("observed variable %r must have effective size 1" % (v,))
        self._compile()

    # -----------------------------------------------------------------------------------------

    def _compile(self):
        n_nodes = len(self.node_vars)
        if self.tree is None:
            # clique graph without a tree: only the init and marginal stages are scheduled
            order = list(range(n_nodes))
            parent = {c: -1 for c in order}
            parent_sep = {c: -1 for c in order}
            depth = {c: 0 for c in order}
            children = {c: [] for c in order}
        else:
            order, parent, parent_sep, depth, children = cons.tree_edges(self.tree)
        cliques = sorted(order)
        self.n_cliques = len(cliques)
        if cliques != list(range(self.n_cliques)):
            raise ValueError("clique ids in the tree must be 0..N-1 (node_list = cliques + separators)")
        self.n_seps = n_nodes - self.n_cliques
        sep_ids = sorted(s for c in order for s, _ in children[c])
        if sep_ids != list(range(self.n_cliques, n_nodes)):
            raise ValueError("separator ids in the tree must be N..N+S-1, each used once")
        self.order, self.parent, self.parent_sep, self.depth, self.children = \
            order, parent, parent_sep, depth, children
        self.root = order[0] if self.tree is not None else -1
        self.max_depth = max(depth.values()) if depth else 0

        for c in order:
            cv = set(self.node_vars[c])
            if len(cv) != len(self.node_vars[c]):
                raise ValueError("duplicate variable in clique %d" % c)
            for s, ch in children[c]:
                sv = self.node_vars[s]
                if len(set(sv)) != len(sv):
                    raise ValueError("duplicate variable in separator %d" % s)
                if not set(sv) <= cv or not set(sv) <= set(self.node_vars[ch]):
                    raise ValueError("separator %d is not contained in both of its cliques" % s)

        self.node_shape = [[int(self.sizes[v]) for v in vs] for vs in self.node_vars]
        self.node_size = [_prod(sh) for sh in self.node_shape]
        self.node_off = [0] * n_nodes
        acc = 0
        for k in range(n_nodes):
            self.node_off[k] = acc
            acc += self.node_size[k]
        self.clique_entries = sum(self.node_size[:self.n_cliques])
        self.sep_entries = sum(self.node_size[self.n_cliques:])
        # workspace regions (entries): [cliques | separator beliefs | up messages | down messages]
        self.up_base = self.clique_entries + self.sep_entries
        self.down_base = self.up_base + self.sep_entries
        # soft evidence: one [size_v][B] likelihood table per variable after the down-messages,
        # multiplied into the smallest clique containing the variable
        self.lik_base = self.down_base + self.sep_entries
        self.lik_off, self.lik_size, self.lik_clique = [], [], []
        self.lik_entries = 0
        for v in self.likelihood_vars:
            holders = [c for c in range(self.n_cliques) if v in self.node_vars[c]]
            if not holders:
                raise ValueError("no clique contains the likelihood variable %r" % (v,))
            self.lik_clique.append(min(holders, key=lambda c: (self.node_size[c], c)))
            self.lik_off.append(self.lik_entries)
            self.lik_size.append(int(self.sizes[v]))
            self.lik_entries += int(self.sizes[v])
        self.work_entries = self.lik_base + self.lik_entries

        self.tab = _TableArena()
        self.tasks, self.msgs, self.launches = [], [], []

        self._build_factor_tables()
        self._find_uniform_cliques()
        self.by_depth = {}
        for c in self.order:
            self.by_depth.setdefault(self.depth[c], []).append(c)
        if self.factors is not None:
            self.clique_factors = [[] for _ in range(self.n_cliques)]
            for f, home in enumerate(self.factor_to_clique):
                self.clique_factors[home].append(f)
        self._blob = None
        if self.emitter == "native":
            self._emit_native()
            return
        self._build_init()
        self._build_collect()
        self._build_distribute()
        self._build_marginal()

        self.tables = self.tab.array()
        self.tasks_arr = np.asarray(self.tasks, np.int64).reshape(-1, TASK_WORDS)
        self.msgs_arr = np.asarray(self.msgs, np.int64).reshape(-1, MSG_WORDS)
        self.launches_arr = np.asarray(self.launches, np.int64).reshape(-1, LAUNCH_WORDS)

    def _emit_native(self):
        """Tasks, messages, launches and index tables from ``jt_plan_build`` (C++).  Variables
        are numbered in order of first appearance; the blob does not depend on the numbering."""
        from . import _native
        ids = {}

        def vid(v):
            if v not in ids:
                ids[v] = len(ids)
            return ids[v]

        node_vars = [[vid(v) for v in vs] for vs in self.node_vars]
        factors = None if self.factors is None else [[vid(v) for v in fv] for fv in self.factors]
        outputs = None if self.outputs is None else [[vid(v) for v in o] for o in self.outputs]
        evidence = [vid(v) for v in self.evidence_vars]
        sizes = [int(self.sizes[v]) for v in ids]
        full = [int(self.full_sizes.get(v, self.sizes[v])) for v in ids]
        tree = None
        if self.tree is not None:
            tree = (self.order, [self.parent[c] for c in range(self.n_cliques)],
                    [self.parent_sep[c] for c in range(self.n_cliques)])
        self._blob = _native.plan_build(sizes, full, self.n_cliques, node_vars, tree, factors,
                                        self.factor_to_clique, evidence, outputs,
                                        [vid(v) for v in self.likelihood_vars])
        words = np.frombuffer(self._blob, np.int64)
        h = words[:H_WORDS]
        n_nodes, F, n_out = int(h[H_NCLIQUES] + h[H_NSEPS]), int(h[H_NFACTORS]), int(h[H_NOUT])
        pos = H_WORDS + 2 * n_nodes + 2 * F + 2 * n_out + int(h[H_NEVID]) + (F + 1 if self.factors is not None else 0) \
            + 2 * int(h[H_NEVF])
        # the emitter's own bookkeeping must agree with the metadata computed above
        expect = (self.n_cliques, self.n_seps, self.clique_entries, self.sep_entries, self.fin_entries,
                  self.fout_entries, self.uni_entries, self.max_depth, self.lik_entries)
        got = tuple(int(h[k]) for k in (H_NCLIQUES, H_NSEPS, H_CLIQUE_ENTRIES, H_SEP_ENTRIES, H_FIN_ENTRIES,
                                        H_FOUT_ENTRIES, H_UNI_ENTRIES, H_MAXDEPTH, H_LIK_ENTRIES))
        if expect != got:
            raise RuntimeError("native plan emitter disagrees with the host metadata: %r vs %r" % (got, expect))

        def take(rows, width):
            nonlocal pos
            arr = words[pos:pos + rows * width].reshape(rows, width)
            pos += rows * width
            return arr

        self.tasks_arr = take(int(h[H_NTASKS]), TASK_WORDS)
        self.msgs_arr = take(int(h[H_NMSGS]), MSG_WORDS)
        self.launches_arr = take(int(h[H_NLAUNCHES]), LAUNCH_WORDS)
        self.tables = np.frombuffer(self._blob, np.int32, count=int(h[H_NTAB]), offset=pos * 8)

    # offsets of the three per-separator buffers
    def bel_off(self, sep):
        return self.node_off[sep]

    def up_off(self, sep):
        return self.up_base + self.node_off[sep] - self.clique_entries

    def down_off(self, sep):
        return self.down_base + self.node_off[sep] - self.clique_entries

    def _strides(self, node):
        return dict(zip(self.node_vars[node], _row_major_strides(self.node_shape[node])))

    # -----------------------------------------------------------------------------------------

    def _build_factor_tables(self):
        self.fin_off, self.fin_size, self.fin_shape = [], [], []
        self.fout_off, self.fout_size, self.fout_shape = [], [], []
        self.out_scopes, self.out_clique = [], []
        self.evf_ptr, self.evf_var, self.evf_stride = [0], [], []
        self.ev_card = [int(self.full_sizes[v]) for v in self.evidence_vars]
        self.fin_entries = self.fout_entries = 0
        if self.factors is None:
            return
        ev_index = {v: i for i, v in enumerate(self.evidence_vars)}
        for f, fv in enumerate(self.factors):
            if len(set(fv)) != len(fv):
                raise ValueError("duplicate variable in factor %d" % f)
            home = self.factor_to_clique[f]
            if not set(fv) <= set(self.node_vars[home]):
                raise ValueError("factor %d is not contained in its clique %d" % (f, home))
            full = [int(self.full_sizes[v]) if v in ev_index else int(self.sizes[v]) for v in fv]
            self.fin_shape.append(full)
            self.fin_off.append(self.fin_entries)
            self.fin_size.append(_prod(full))
            self.fin_entries += _prod(full)
            for v, st in zip(fv, _row_major_strides(full)):
                if v in ev_index:
                    self.evf_var.append(ev_index[v])
                    self.evf_stride.append(st)
            self.evf_ptr.append(len(self.evf_var))
        # output scopes of the marginal stage
        if self.outputs is None:
            scopes = [(list(fv), home) for fv, home in zip(self.factors, self.factor_to_clique)]
        else:
            scopes = []
            for scope in self.outputs:
                if len(set(scope)) != len(scope):
                    raise ValueError("duplicate variable in output scope %r" % (scope,))
                holders = [c for c in range(self.n_cliques) if set(scope) <= set(self.node_vars[c])]
                if not holders:
                    raise ValueError("no clique contains the output scope %r" % (scope,))
                scopes.append((list(scope), min(holders, key=lambda c: (self.node_size[c], c))))
        for scope, home in scopes:
            eff = [int(self.sizes[v]) for v in scope]
            self.out_scopes.append(scope)
            self.out_clique.append(home)
            self.fout_shape.append(eff)
            self.fout_off.append(self.fout_entries)
            self.fout_size.append(_prod(eff))
            self.fout_entries += _prod(eff)

    def _find_uniform_cliques(self):
        """Which potentials and up-messages are identical for every instance of a batch.

        With factor tables shared by the batch, psi_C depends on the instance only through the
        observed axes of its assigned factors, and an up-message only through the cliques below
        it.  ``uniform[c]``: no factor of clique c contains an observed variable;
        ``uniform_up[c]``: the same holds for the whole subtree of c.  In *uniform mode* such
        operands are kept once, in a B = 1 copy of the workspace with the same entry offsets
        (the uniform workspace), evidence-free subtrees are collected there once, and the batch
        kernels broadcast them instead of streaming [n][B] rows.  Per-instance factor tables
        switch uniform mode off at run time; the flags are then ignored.
        """
        self.uniform = [False] * self.n_cliques
        self.uniform_up = [False] * self.n_cliques
        self.uniform_down = {}                     # separator node -> its down-message is uniform
        self.uni_entries = 0
        if self.tree is not None:
            self.uniform_down = {s: False for c in self.order for s, _ in self.children[c]}
        if self.factors is None or self.tree is None:
            return
        observed = set(self.evidence_vars)
        touched = set(home for fv, home in zip(self.factors, self.factor_to_clique) if observed & set(fv))
        touched.update(self.lik_clique)               # a likelihood makes the potential per-instance
        if not self.children[self.root]:
            touched.add(self.root)        # a single clique gets no distribute task: keep it per instance
        for c in reversed(self.order):
            self.uniform[c] = c not in touched
            self.uniform_up[c] = self.uniform[c] and all(self.uniform_up[k] for _, k in self.children[c])
        self.uni_entries = sum(self.node_size[c] for c in range(self.n_cliques) if self.uniform[c])
        # a down-message is uniform when everything on its source side is: the clique's
        # potential, the message from above and the up-messages of the other children
        for c in self.order:
            above = self.parent[c] < 0 or self.uniform_down[self.parent_sep[c]]
            for s, _ in self.children[c]:
                self.uniform_down[s] = self.uniform[c] and above and all(
                    self.uniform_up[k2] for s2, k2 in self.children[c] if s2 != s)

    def _add_msg(self, off, s_space, r_space, stride_of, fid=-1, uniform=False):
        a_hi, a_lo = s_space.tables(stride_of)
        row = [0] * MSG_WORDS
        row[M_OFF] = off
        row[M_UNI] = 1 if uniform else 0
        row[M_AHI], row[M_ALO] = self.tab.add(a_hi), self.tab.add(a_lo)
        if r_space is not None:
            b_hi, b_lo = r_space.tables(stride_of)
            row[M_BHI], row[M_BLO] = self.tab.add(b_hi), self.tab.add(b_lo)
        row[M_FID] = fid
        self.msgs.append(row)

    def _new_task(self, kind, s_space, r_space, node, src_node=None, src_is_psi=False):
        row = [0] * TASK_WORDS
        row[T_KIND] = kind
        for w in (T_SRC, T_OUT, T_BETA, T_BEL, T_OWN):
            row[w] = -1
        if src_is_psi and src_node is not None and self.uniform[src_node]:
            row[T_FLAGS] |= TF_SRC_UNIFORM
        row[T_NS], row[T_NSLO] = s_space.n, s_space.n_lo
        row[T_NR], row[T_NRLO] = (r_space.n, r_space.n_lo) if r_space is not None else (1, 1)
        row[T_NODE] = node
        if src_node is not None:
            st = self._strides(src_node)
            row[T_SRC] = self.node_off[src_node]
            hi, lo = s_space.tables(st)
            row[T_SRC_SHI], row[T_SRC_SLO] = self.tab.add(hi), self.tab.add(lo)
            hi, lo = r_space.tables(st)
            row[T_SRC_RHI], row[T_SRC_RLO] = self.tab.add(hi), self.tab.add(lo)
        return row

    def _attach_msgs(self, row, msgs, s_space, r_space):
        """msgs: list of (workspace offset, separator node, uniform).  r-dependent ones first."""
        rdep = [m for m in msgs if r_space.touches(set(self.node_vars[m[1]]))]
        sonly = [m for m in msgs if not r_space.touches(set(self.node_vars[m[1]]))]
        row[T_RMSG_BEGIN] = len(self.msgs)
        for off, sep, uni in rdep:
            self._add_msg(off, s_space, r_space, self._strides(sep), uniform=uni)
        row[T_RMSG_END] = row[T_SMSG_BEGIN] = len(self.msgs)
        for off, sep, uni in sonly:
            self._add_msg(off, s_space, None, self._strides(sep), uniform=uni)
        row[T_SMSG_END] = len(self.msgs)

    def _launch(self, phase, begin, level):
        if len(self.tasks) > begin:
            self.launches.append([phase, begin, len(self.tasks), level])

    # -----------------------------------------------------------------------------------------

    def _build_init(self):
        """E0 + V1: psi_C = prod of assigned factors, observed axes gathered per instance.

        Uniform cliques come first, so one task list serves three launches: all cliques per
        instance (general mode), the uniform ones once into the uniform workspace, the others
        per instance (uniform mode)."""
        if self.factors is None:
            return
        by_clique = self.clique_factors
        begin = len(self.tasks)
        ordered = [c for c in range(self.n_cliques) if self.uniform[c]] + \
                  [c for c in range(self.n_cliques) if not self.uniform[c]]
        n_uniform = sum(self.uniform)
        for c in ordered:
            s_space = _Space(self.node_vars[c], self.sizes)
            row = self._new_task(KIND_INIT, s_space, None, c)
            row[T_OUT] = self.node_off[c]
            if self.uniform[c]:
                row[T_FLAGS] |= TF_TASK_UNIFORM
            row[T_SMSG_BEGIN] = row[T_RMSG_BEGIN] = row[T_RMSG_END] = len(self.msgs)
            for f in by_clique[c]:
                st = dict(zip(self.factors[f], _row_major_strides(self.fin_shape[f])))
                # observed axes contribute through the per-instance base offset only
                for v in self.evidence_vars:
                    st.pop(v, None)
                self._add_msg(self.fin_off[f], s_space, None, st, fid=f)
            for k, v in enumerate(self.likelihood_vars):
                if self.lik_clique[k] == c:           # fid -2: operand read from the workspace
                    self._add_msg(self.lik_base + self.lik_off[k], s_space, None, {v: 1}, fid=-2)
            row[T_SMSG_END] = len(self.msgs)
            self.tasks.append(row)
        self._launch_split(PHASE_INIT, PHASE_INIT_UNIFORM, PHASE_INIT_INSTANCE, begin, begin + n_uniform, 0)

    def _launch_split(self, phase_all, phase_uniform, phase_instance, begin, middle, level):
        end = len(self.tasks)
        if end > begin:
            self.launches.append([phase_all, begin, end, level])
        if middle > begin:
            self.launches.append([phase_uniform, begin, middle, level])
        if end > middle:
            self.launches.append([phase_instance, middle, end, level])

    def _build_collect(self):
        """E1 + E2: up-messages, deepest level first (uniform subtrees first within a level)."""
        by_depth = self.by_depth
        for d in range(self.max_depth, 0, -1):
            begin = len(self.tasks)
            ordered = [c for c in by_depth[d] if self.uniform_up[c]] + \
                      [c for c in by_depth[d] if not self.uniform_up[c]]
            n_uniform = sum(1 for c in by_depth[d] if self.uniform_up[c])
            for c in ordered:
                psep = self.parent_sep[c]
                s_space = _Space(self.node_vars[psep], self.sizes)
                in_sep = set(self.node_vars[psep])
                r_space = _Space([v for v in self.node_vars[c] if v not in in_sep], self.sizes)
                row = self._new_task(KIND_PROJECT, s_space, r_space, c, src_node=c, src_is_psi=True)
                row[T_OUT] = self.up_off(psep)
                if self.uniform_up[c]:
                    row[T_FLAGS] |= TF_TASK_UNIFORM
                self._attach_msgs(row, [(self.up_off(s), s, self.uniform_up[k]) for s, k in self.children[c]],
                                  s_space, r_space)
                self.tasks.append(row)
            self._launch_split(PHASE_COLLECT, PHASE_COLLECT_UNIFORM, PHASE_COLLECT_INSTANCE,
                               begin, begin + n_uniform, d)

    def _down_msg(self, c):
        """The message clique c receives from its parent, as an incoming-message triple."""
        psep = self.parent_sep[c]
        return (self.down_off(psep), psep, self.uniform_down[psep])

    def _build_distribute(self):
        """E3 + E4 + M1 + E5 with a division-free exclude-one product, top level first.

        For a clique with k children there are k projection tasks (one per child separator);
        the last one also writes the clique belief in place, so it runs in a second launch
        after the others of its level have read psi_C.  A non-root leaf gets one elementwise
        task.  psi_C is read max(k, 1) times and written once (the reference reads it k+2
        times, ``computation.py:169-224``).

        Uniform mode: a down-message whose whole source side is evidence-free is the same for
        every instance.  It is computed once, in the uniform workspace (``PHASE_DIST_UNIFORM``
        tasks, one per such separator), and every consumer reads that copy as a broadcast
        scalar instead of streaming an [n][B] row per item.  A non-writer task whose message
        is uniform shrinks to an elementwise task in the instance launch (separator belief =
        uniform down x up, plus the per-instance copy of the down-message); task order of the
        first launch of a level: ``[full form, uniform down | full form, others | elementwise
        forms]`` -- ``PHASE_DIST_PRE`` is the first two groups, ``PHASE_DIST_PRE_INSTANCE``
        the last two.
        """
        for d in range(0, self.max_depth + 1):
            pre_ud, pre_other, main, inst, unis = [], [], [], [], []
            for c in self.by_depth.get(d, []):
                kids = self.children[c]
                incoming = []
                if self.parent[c] >= 0:
                    incoming.append(self._down_msg(c))
                incoming += [(self.up_off(s), s, self.uniform_up[k]) for s, k in kids]
                if not kids:
                    if self.parent[c] < 0:
                        continue  # single-clique tree: belief = potential
                    s_space = _Space(self.node_vars[c], self.sizes)
                    r_space = _Space([], self.sizes)
                    row = self._new_task(KIND_PROJECT, s_space, r_space, c, src_node=c, src_is_psi=True)
                    row[T_BETA] = self.node_off[c]
                    self._attach_msgs(row, incoming, s_space, r_space)
                    main.append(row)
                    continue
                for i, (sep, kid) in enumerate(kids):
                    s_space = _Space(self.node_vars[sep], self.sizes)
                    in_sep = set(self.node_vars[sep])
                    r_space = _Space([v for v in self.node_vars[c] if v not in in_sep], self.sizes)
                    row = self._new_task(KIND_PROJECT, s_space, r_space, c, src_node=c, src_is_psi=True)
                    row[T_OUT] = self.down_off(sep)
                    row[T_BEL] = self.bel_off(sep)
                    row[T_OWN] = self.up_off(sep)
                    if self.uniform_up[kid]:
                        row[T_FLAGS] |= TF_OWN_UNIFORM
                    others = [m for m in incoming if m[1] != sep]
                    writer = i == len(kids) - 1
                    if writer:
                        row[T_BETA] = self.node_off[c]
                    # messages are appended when the task is placed (keeps ranges contiguous)
                    item = (row, others, s_space, r_space)
                    (main if writer else pre_ud if self.uniform_down[sep] else pre_other).append(item)
                    if self.uniform_down[sep]:
                        u_row = self._new_task(KIND_PROJECT, s_space, r_space, c, src_node=c, src_is_psi=True)
                        u_row[T_OUT] = self.down_off(sep)
                        u_row[T_FLAGS] |= TF_TASK_UNIFORM
                        unis.append((u_row, others, s_space, r_space))
                        if not writer:
                            none = _Space([], self.sizes)
                            i_row = self._new_task(KIND_PROJECT, s_space, none, c, src_node=sep)
                            i_row[T_SRC] = self.down_off(sep)
                            i_row[T_FLAGS] |= TF_SRC_UNIFORM
                            i_row[T_OUT] = self.down_off(sep)
                            i_row[T_BEL] = self.bel_off(sep)
                            i_row[T_OWN] = self.up_off(sep)
                            if self.uniform_up[kid]:
                                i_row[T_FLAGS] |= TF_OWN_UNIFORM
                            inst.append((i_row, [], s_space, none))
            # tasks that send a message first, the belief-only ones (leaves) last: when clique
            # beliefs are not wanted the launch stops before them
            main.sort(key=lambda item: 0 if isinstance(item, tuple) else 1)
            begin = len(self.tasks)
            for row, others, s_space, r_space in pre_ud + pre_other + inst:
                self._attach_msgs(row, others, s_space, r_space)
                self.tasks.append(row)
            n_ud, n_full = len(pre_ud), len(pre_ud) + len(pre_other)
            if n_full:
                self.launches.append([PHASE_DIST_PRE, begin, begin + n_full, d])
            if len(self.tasks) > begin + n_ud:
                self.launches.append([PHASE_DIST_PRE_INSTANCE, begin + n_ud, len(self.tasks), d])
            begin = len(self.tasks)
            n_sending = 0
            for item in main:
                if isinstance(item, tuple):
                    row, others, s_space, r_space = item
                    self._attach_msgs(row, others, s_space, r_space)
                    self.tasks.append(row)
                    n_sending += 1
                else:
                    self.tasks.append(item)
            self._launch(PHASE_DIST_MAIN, begin, d)
            if n_sending:
                self.launches.append([PHASE_DIST_MAIN_MESSAGES, begin, begin + n_sending, d])
            begin = len(self.tasks)
            for row, others, s_space, r_space in unis:
                self._attach_msgs(row, others, s_space, r_space)
                self.tasks.append(row)
            self._launch(PHASE_DIST_UNIFORM, begin, d)

    def _build_marginal(self):
        """E6: per output scope (by default per factor) = clique belief summed down to the scope.

        Two forms: from the stored belief beta_C (``PHASE_MARGINAL``), or *direct* from psi_C
        times every incoming message (``PHASE_MARGINAL_DIRECT``), which does not need the clique
        beliefs to be written at all -- used when only the outputs are wanted."""
        if self.factors is None:
            return
        begin = len(self.tasks)
        for k, (scope, c) in enumerate(zip(self.out_scopes, self.out_clique)):
            s_space = _Space(scope, self.sizes)
            in_f = set(scope)
            r_space = _Space([v for v in self.node_vars[c] if v not in in_f], self.sizes)
            row = self._new_task(KIND_PROJECT, s_space, r_space, c, src_node=c)
            row[T_OUT] = self.fout_off[k]
            row[T_OUT_SPACE] = SPACE_FOUT
            row[T_AUX] = k
            row[T_RMSG_BEGIN] = row[T_RMSG_END] = row[T_SMSG_BEGIN] = row[T_SMSG_END] = len(self.msgs)
            self.tasks.append(row)
        self._launch(PHASE_MARGINAL, begin, 0)
        if self.tree is None:
            return
        begin = len(self.tasks)
        for k, (scope, c) in enumerate(zip(self.out_scopes, self.out_clique)):
            s_space = _Space(scope, self.sizes)
            in_f = set(scope)
            r_space = _Space([v for v in self.node_vars[c] if v not in in_f], self.sizes)
            row = self._new_task(KIND_PROJECT, s_space, r_space, c, src_node=c, src_is_psi=True)
            row[T_OUT] = self.fout_off[k]
            row[T_OUT_SPACE] = SPACE_FOUT
            row[T_AUX] = k
            incoming = []
            if self.parent[c] >= 0:
                incoming.append(self._down_msg(c))
            incoming += [(self.up_off(sp), sp, self.uniform_up[kid]) for sp, kid in self.children[c]]
            self._attach_msgs(row, incoming, s_space, r_space)
            self.tasks.append(row)
        self._launch(PHASE_MARGINAL_DIRECT, begin, 0)

    # -----------------------------------------------------------------------------------------

    def algorithmic_entries(self, with_init=True):
        """Entries moved per instance by the minimal two-pass schedule:
        4*sum(n_C) - n_root + 6*sum(n_S)  (3*sum(n_C) without the init write)."""
        n_root = self.node_size[self.root] if self.root >= 0 else 0
        return (4 if with_init else 3) * self.clique_entries - n_root + 6 * self.sep_entries

    def scheduled_entries(self, uniform=True, sep_beliefs=True, beliefs=True):
        """Entries per instance this schedule reads + writes in HBM for init + collect +
        distribute, every buffer counted once per consuming task (the accounting of SURVEY.md 8d
        applied to the schedule as built).  In uniform mode the potential of a uniform clique is
        neither written nor read per instance (only its belief is written), a uniform
        up-message is not written or read per instance (only down-message and belief are), and a
        uniform down-message is written per instance (for callers that read it) while the child
        reads the uniform copy.  ``beliefs=False`` (JT_NO_BELIEFS, the pipelines): no clique or
        separator belief is written; instead every output scope is projected directly from the
        potential and the incoming messages of its clique, and the outputs are written."""
        total = 0
        scopes = [0] * self.n_cliques
        if not beliefs:
            for c in self.out_clique:
                scopes[c] += 1
            total += self.fout_entries
            sep_beliefs = False
        for c in self.order:
            n = self.node_size[c]
            kids = self.children[c]
            if beliefs:
                reads = (0 if c == self.root else 1) + max(len(kids), 1 if c != self.root else 0)
                if uniform and self.uniform[c]:
                    total += n if (kids or c != self.root) else 0       # belief write
                else:
                    total += n * (1 + reads) + (n if (kids or c != self.root) else 0)
            elif not (uniform and self.uniform[c]):
                # init write, collect read, one read per message sent down, one per output scope
                total += n * (1 + (0 if c == self.root else 1) + len(kids) + scopes[c])
            for s, k in kids:
                ns = self.node_size[s]
                down_read = 0 if (uniform and self.uniform_down[s]) else 1 + (0 if beliefs else scopes[k])
                if uniform and self.uniform_up[k]:
                    total += ns * (1 + down_read + (1 if sep_beliefs else 0))   # down write (+ read), belief write
                else:
                    # up: write + parent's collect read + parent's distribute reads + own read
                    total += ns * (4 + down_read + (1 if sep_beliefs else 0) + (0 if beliefs else scopes[c]))
        return total

    def header(self):
        h = np.zeros(H_WORDS, np.int64)
        h[H_MAGIC], h[H_VERSION] = MAGIC, VERSION
        h[H_NCLIQUES], h[H_NSEPS] = self.n_cliques, self.n_seps
        h[H_NFACTORS] = 0 if self.factors is None else len(self.factors)
        h[H_NEVID] = len(self.evidence_vars)
        h[H_CLIQUE_ENTRIES], h[H_SEP_ENTRIES] = self.clique_entries, self.sep_entries
        h[H_FIN_ENTRIES], h[H_FOUT_ENTRIES] = self.fin_entries, self.fout_entries
        h[H_NTAB] = self.tables.size
        h[H_NTASKS], h[H_NMSGS], h[H_NLAUNCHES] = len(self.tasks_arr), len(self.msgs_arr), len(self.launches_arr)
        h[H_MAXDEPTH] = self.max_depth
        h[H_NEVF] = len(self.evf_var)
        h[H_ROOT_ENTRIES] = self.node_size[self.root] if self.root >= 0 else 0
        h[H_UNI_ENTRIES] = self.uni_entries
        h[H_NOUT] = len(self.fout_off)
        h[H_LIK_ENTRIES] = self.lik_entries
        return h

    def to_blob(self):
        """Serialise to the byte blob ``jt_plan_create`` parses (layout: include/jt_b200.h)."""
        if self._blob is not None:
            return self._blob
        i64 = lambda xs: np.asarray(xs, np.int64).reshape(-1)
        tables = self.tables
        if tables.size % 2:
            tables = np.concatenate([tables, np.zeros(1, np.int32)])
        parts = [
            self.header(),
            i64(self.node_off), i64(self.node_size),
            i64(self.fin_off), i64(self.fin_size), i64(self.fout_off), i64(self.fout_size),
            i64(self.ev_card), i64(self.evf_ptr if self.factors is not None else []),
            i64(self.evf_var), i64(self.evf_stride),
            i64(self.tasks_arr), i64(self.msgs_arr), i64(self.launches_arr),
        ]
        return b"".join(p.tobytes() for p in parts) + tables.tobytes()

    def summary(self):
        return {
            "cliques": self.n_cliques, "separators": self.n_seps,
            "clique_entries": self.clique_entries, "sep_entries": self.sep_entries,
            "max_clique": max(self.node_size[:self.n_cliques]) if self.n_cliques else 0,
            "root_entries": int(self.node_size[self.root]) if self.root >= 0 else 0,
            "depth": self.max_depth, "tasks": len(self.tasks_arr), "launches": len(self.launches_arr),
            "table_entries": int(self.tables.size),
            "algorithmic_entries": self.algorithmic_entries(),
            "uniform_clique_entries": self.uni_entries,
            "uniform_up_entries": sum(self.node_size[self.parent_sep[c]] for c in self.order
                                      if self.parent[c] >= 0 and self.uniform_up[c]),
        }
