"""
User interface of the junction tree library (B200 build).

Mirror of ``/root/reference/junctiontree/junctiontree.py``: ``create_junction_tree`` (``:12-16``),
``FactorGraph`` (``:83-117``), ``CliqueGraph`` (``:120-274``) and ``JunctionTree`` (``:277-331``)
keep their names, attributes and call signatures.  Compilation stays on the host
(``construction.py``); ``evaluate`` / ``propagate`` / ``marginalize`` run on the GPU through the
compiled schedule (``schedule.py`` -> ``libjt_b200.so``).  ``propagate_batch`` is new: many
independent propagations over the same tree per launch, with per-instance evidence.
"""

import weakref

import numpy as np
import attr

from . import computation as comp
from . import construction as cons
from . import engine as eng
from . import schedule as sch


#: batches larger than this with host output are streamed through the chunked pipeline
_STREAM_THRESHOLD = 4096


def create_junction_tree(factors, sizes, order=None):
    """Create a Junction tree for a given factor graph.

    ``order`` (optional, new) is an elimination order over the variables; the default is
    min-fill.  Grid-like models need a sweep order to reach their treewidth."""
    assert all(type(l) == list for l in factors), "Provided factor is not a list"
    fg = FactorGraph(factors=factors, sizes=sizes)
    return fg.triangulate(order=order).create_junction_tree()


def einsum(xs, xs_keys, y_keys):
    """Product of ``xs`` (axes ``xs_keys``) summed to ``y_keys``, arbitrary keys; keys that only
    occur in the output become size-1 axes (reference ``junctiontree.py:34-80``).  Runs on the
    device through the sum-product plugin."""
    xs = [np.asarray(x) for x in xs]
    xs_keys = [list(k) for k in xs_keys]
    present = set(k for keys in xs_keys for k in keys)
    missing = [k for k in y_keys if k not in present]
    if xs:
        xs[0] = np.reshape(xs[0], len(missing) * (1,) + np.shape(xs[0]))
        xs_keys[0] = missing + xs_keys[0]
    args = [arg for pair in zip(xs, xs_keys) for arg in pair] + [list(y_keys)]
    return comp.sum_product.einsum(*args)


@attr.s(frozen=True)
class FactorGraph():
    """A graph containing a set of nodes that each contain a set of variables.

    Each variable has a corresponding size associated to it.
    """

    # Axis variables in each factor
    factors = attr.ib()

    # Size of each axis
    sizes = attr.ib()

    def triangulate(self, order=None):
        """Create a triangulated clique tree from a factor graph."""
        (_, maxcliques, factor_to_maxclique) = cons.find_triangulation(
            self.factors,
            self.sizes,
            order
        )
        return CliqueGraph(
            maxcliques=maxcliques,
            factor_to_maxclique=factor_to_maxclique,
            factor_graph=self,
        )


def _effective_sizes(factors, xs, batched=False):
    """Variable sizes as found in the factor arrays (the reference derives all clique shapes
    from the arrays and only the throw-away separator ones from ``sizes``,
    ``junctiontree.py:311-315``)."""
    sizes = {}
    for f, (fv, x) in enumerate(zip(factors, xs)):
        shape = tuple(np.shape(x))[1:] if batched else tuple(np.shape(x))
        if len(shape) != len(fv):
            raise ValueError("factor %d has variables %r but its array has shape %s" % (f, fv, shape))
        for var, n in zip(fv, shape):
            if sizes.setdefault(var, int(n)) != int(n):
                raise ValueError("variable %r has size %d in one factor array and %d in another"
                                 % (var, sizes[var], n))
    return sizes


# Compiled state (engines, single-call runners) lives outside the attrs objects, so that
# ``CliqueGraph`` and ``JunctionTree`` stay frozen value classes as in the reference
# (``junctiontree.py:120,277``): owner id -> {cache name -> dict}, dropped with the owner.
_caches = {}


#: compiled engines kept per clique graph (see ``CliqueGraph._engine``)
_MAX_ENGINES = 64


def _cache_of(owner, name):
    slot = _caches.get(id(owner))
    if slot is None:
        slot = _caches[id(owner)] = {}
        weakref.finalize(owner, _caches.pop, id(owner), None)
    return slot.setdefault(name, {})


@attr.s(frozen=True)
class CliqueGraph():
    """
    Clique graph for an underlying factor graph.
    """

    # Axis variables in each maximal clique
    maxcliques = attr.ib()

    # Maximal clique for each factor (multiple factors can belong to the same
    # maximal clique)
    factor_to_maxclique = attr.ib()

    # The underlying factor graph
    factor_graph = attr.ib()

    @property
    def _engines(self):
        """Compiled engines of this clique graph (one per tree, sizes, evidence set, outputs and
        device; see ``_engine``)."""
        return _cache_of(self, "engines")

    def create_junction_tree(self):
        """Create a Junction tree from a triangulated clique tree."""
        (tree, separators) = cons.construct_junction_tree(
            self.maxcliques,
            self.factor_graph.sizes
        )
        return JunctionTree(
            tree=tree,
            separators=separators,
            clique_tree=self
        )

    def _f2c(self):
        f2c = self.factor_to_maxclique
        return [f2c[i] for i in range(len(self.factor_graph.factors))]   # list or dict (reference D13)

    def _engine(self, sizes, tree=None, separators=(), evidence_vars=(), full_sizes=None, outputs=None,
                likelihood_vars=()):
        # everything the plan is compiled from, and the device its descriptors live on: a second
        # JunctionTree over this clique graph (re-rooted, other separator axis order) or another
        # current device gets its own engine
        key = (tuple(sizes.get(v) for c in self.maxcliques for v in c),
               None if tree is None else comp._tree_key(tree),
               None if tree is None else tuple(tuple(sep) for sep in separators),
               tuple(evidence_vars), tuple(tuple(c) for c in self.maxcliques), tuple(self._f2c()),
               None if outputs is None else tuple(tuple(o) for o in outputs), tuple(likelihood_vars),
               None if full_sizes is None else tuple(full_sizes.get(v) for v in evidence_vars),
               eng.current_device())
        engines = self._engines
        hit = engines.pop(key, None)
        if hit is None:
            node_vars = list(self.maxcliques) + ([list(s) for s in separators] if tree is not None else [])
            plan = sch.Plan(tree, node_vars, sizes, self.factor_graph.factors, self._f2c(),
                            evidence_vars, full_sizes, outputs, likelihood_vars=likelihood_vars)
            hit = eng.Engine(plan)
            # least recently used first out: every distinct evidence pattern of propagate_evidence
            # compiles its own plan (descriptors on the device, cached workspaces); an engine still
            # held by a caller (a session) simply stays alive outside the cache
            while len(engines) >= _MAX_ENGINES:
                engines.pop(next(iter(engines)))
        engines[key] = hit                      # (re)inserted last: dicts keep insertion order
        return hit

    def evaluate(self, xs, dl=None, reference_shapes=True):
        """Compute maximum clique values based on factor values.

        GPU stage ``jt_init``.  As in the reference (``junctiontree.py:52-61``, pinned by its
        ``tests/test_junctiontree.py:88-109``) a clique variable that none of the assigned
        factors covers comes back as a size-1 axis; the device computes the full-size potential
        (constant along such an axis) and the result is sliced.  ``reference_shapes=False``
        returns the full-size arrays the propagation stages work on."""
        t = eng.require_cuda()
        sizes = dict(self.factor_graph.sizes)
        sizes.update(_effective_sizes(self.factor_graph.factors, xs))
        engine = self._engine(sizes)
        plan = engine.plan
        dtype = _result_dtype(xs)
        fdev, _ = engine.factors_to_device(xs, dtype)
        ws = engine.workspace(1, dtype)
        engine.dev.upload()
        engine.dev.init(fdev.data_ptr(), False, None, 1, dtype, ws.data_ptr(), _semiring(dl), engine._stream())
        flat = engine.work_view(ws, 1, dtype)[:plan.clique_entries, 0].cpu().numpy()
        out = [
            flat[plan.node_off[c]:plan.node_off[c] + plan.node_size[c]].reshape(tuple(plan.node_shape[c])).copy()
            for c in range(plan.n_cliques)
        ]
        if reference_shapes:
            covered = [set() for _ in self.maxcliques]
            for fv, home in zip(self.factor_graph.factors, self._f2c()):
                covered[home].update(fv)
            out = [y[tuple(slice(None) if v in covered[c] else slice(0, 1) for v in cv)]
                   for c, (cv, y) in enumerate(zip(self.maxcliques, out))]
        return out

    def marginalize(self, ys, dl=None):
        """Marginalize results for maxcliques to results for factors

        For each factor, take the maxclique it belongs to and sum out the axes that don't belong
        to the factor; axes come out in the factor's own order (reference
        ``junctiontree.py:229-274``).  GPU stage ``jt_marginal``.

        Every clique array is taken with the shape it has -- as in the reference, where each
        factor is one einsum over its own clique -- so the size-1 axes ``evaluate`` returns for
        uncovered clique variables are summed as axes of length 1.
        """
        t = eng.require_cuda()
        ys = [np.asarray(y) for y in ys]
        if len(ys) != len(self.maxcliques):
            raise ValueError("expected %d clique arrays, got %d" % (len(self.maxcliques), len(ys)))
        for cv, y in zip(self.maxcliques, ys):
            if y.ndim != len(cv):
                raise ValueError("clique %r cannot have an array of shape %s" % (cv, y.shape))
        cache = _cache_of(self, "marginalize")
        key = (tuple(y.shape for y in ys), tuple(self._f2c()), eng.current_device())
        engine = cache.get(key)
        if engine is None:
            # cliques are independent here, so each gets its own copy (c, var) of its variables:
            # a variable may then have length 1 in one clique array and its full size in another
            f2c = self._f2c()
            sizes = {(c, v): int(n) for c, (cv, y) in enumerate(zip(self.maxcliques, ys)) for v, n in zip(cv, y.shape)}
            node_vars = [[(c, v) for v in cv] for c, cv in enumerate(self.maxcliques)]
            factors = [[(f2c[f], v) for v in fv] for f, fv in enumerate(self.factor_graph.factors)]
            engine = eng.Engine(sch.Plan(None, node_vars, sizes, factors, f2c))
            if len(cache) >= 16:
                cache.pop(next(iter(cache)))
            cache[key] = engine
        plan = engine.plan
        dtype = _result_dtype(ys)
        ws = engine.workspace(1, dtype)
        work = engine.work_view(ws, 1, dtype)
        host = np.empty(plan.clique_entries, dtype)
        for c, y in enumerate(ys):
            host[plan.node_off[c]:plan.node_off[c] + plan.node_size[c]] = y.reshape(-1)
        work[:plan.clique_entries, 0].copy_(t.from_numpy(host))
        fout = t.empty((plan.fout_entries, 1), dtype=eng.torch_dtype(dtype), device="cuda")
        engine.dev.upload()
        engine.dev.marginal(1, dtype, ws.data_ptr(), fout.data_ptr(), engine._stream(), _semiring(dl))
        flat = fout[:, 0].cpu().numpy()
        return [
            flat[plan.fout_off[f]:plan.fout_off[f] + plan.fout_size[f]].reshape(tuple(plan.fout_shape[f])).copy()
            for f in range(len(plan.fout_off))
        ]


def _likelihood_vars(factor_graph, likelihoods):
    """Soft-evidence variables in a canonical order (first appearance in the factors)."""
    if not likelihoods:
        return []
    seen = []
    for fv in factor_graph.factors:
        for v in fv:
            if v in likelihoods and v not in seen:
                seen.append(v)
    unknown = [v for v in likelihoods if v not in seen]
    if unknown:
        raise ValueError("likelihoods on unknown variables %r" % (unknown,))
    return seen


def _semiring(dl):
    """``JT_SR_*`` flag of a distributive law (``None``: sum-product).  Only laws bound to the
    device kernels are accepted here; a law built around a user einsum function goes through
    ``computation.compute_beliefs(..., dl)``."""
    if dl is None:
        return 0
    if not getattr(dl, "on_device", False):
        raise ValueError("this entry point runs on the device kernels; a distributive law wrapping a user "
                         "einsum function is only accepted by computation.compute_beliefs")
    return int(dl.semiring_flag)


def _result_dtype(arrays):
    """float32 only when every input is float32; otherwise float64 (the reference's float64
    separators promote everything, ``junctiontree.py:311-315``)."""
    kinds = [np.asarray(a).dtype for a in arrays]
    if kinds and all(k == np.float32 for k in kinds):
        return np.dtype(np.float32)
    return np.dtype(np.float64)


@attr.s(frozen=True)
class JunctionTree():
    """
    Junction tree for an underlying factor graph.
    """

    # Tree data structure
    #
    # (cliqueID, (separatorID, subtree), (separatorID, subtree), ...)
    tree = attr.ib()

    # Tuple of axis vars in each separator
    #
    # ( (var3, var1), (var2, var1), (var2) )
    separators = attr.ib()

    # The underlying triangulated clique graph
    clique_tree = attr.ib()

    def _engine(self, sizes, evidence_vars=(), full_sizes=None, outputs=None, likelihood_vars=()):
        return self.clique_tree._engine(sizes, self.tree, self.separators, evidence_vars, full_sizes, outputs,
                                        likelihood_vars)

    def plan(self, evidence_vars=(), sizes=None):
        """The compiled schedule for the current variable sizes (``schedule.Plan``)."""
        full = dict(self.clique_tree.factor_graph.sizes) if sizes is None else dict(sizes)
        eff = dict(full)
        for v in evidence_vars:
            eff[v] = 1
        return self._engine(eff, evidence_vars, full).plan

    def propagate(self, xs, dtype=None, dl=None):
        """Run belief propagation on the Junction tree.

        :param xs: one array per factor (shape = sizes of the factor's variables; an observed
                   variable is conditioned on by slicing its axis to length 1, reference
                   ``README.md:148-166``)
        :param dl: distributive law (default sum-product, the only one of the reference,
                   ``junctiontree.py:300-305``); ``semirings.max_product`` returns max-marginals,
                   ``log_sum_exp`` / ``max_sum`` work on log potentials
        :return: one array per factor with the same shape: the consistent (unnormalised)
                 clique belief summed down to the factor's variables (reference
                 ``junctiontree.py:297-331``)
        """
        ct = self.clique_tree
        xs = [x if isinstance(x, np.ndarray) else np.asarray(x) for x in xs]
        # (shapes, dtypes, semiring, device) of a single propagate() call -> (runner, output
        # slots), per JunctionTree: another tree over the same clique graph compiles its own
        single = _cache_of(self, "single")
        key = (tuple(x.shape for x in xs), tuple(x.dtype.char for x in xs), dtype, id(dl), eng.current_device())
        hit = single.get(key)
        if hit is None:
            fg = ct.factor_graph
            sizes = dict(fg.sizes)
            sizes.update(_effective_sizes(fg.factors, xs))
            engine = self._engine(sizes)
            plan = engine.plan
            # single instance: latency-bound.  Small trees: one library call (tables in, one
            # kernel, beliefs out); larger ones: one CUDA-graph replay over static buffers
            runner = engine.host_runner(1, np.dtype(dtype) if dtype is not None else _result_dtype(xs),
                                        semiring=_semiring(dl))
            slots = [(plan.fout_off[f], plan.fout_off[f] + plan.fout_size[f], tuple(plan.fout_shape[f]))
                     for f in range(len(plan.fout_off))]
            hit = (runner, slots, dl)                    # dl is kept alive so that its id stays unique
            if len(single) >= 16:
                single.pop(next(iter(single)))
            single[key] = hit
        runner, slots, _ = hit
        runner.set_factors(xs)
        flat = runner.run().numpy()[:, 0].copy()         # one fresh array; the outputs are views of it
        return [flat[lo:hi].reshape(shape) for lo, hi, shape in slots]

    def marginals_batch(self, xs, variables=None, evidence_vars=(), evidence=None, batch=None, dtype=None,
                        normalize=True, dl=None, likelihoods=None):
        """Posterior marginals of single variables for a batch of evidence (output stage on the
        device: the step after the propagation path).

        :param variables: variables to report (default: every unobserved variable)
        :param normalize: divide each marginal by its sum (the default) or return the
                          unnormalised beliefs, as ``propagate`` does
        :param likelihoods: soft evidence ``{variable: array [B, size]}`` (see ``propagate_batch``)
        :param dl: distributive law; with ``max_product`` / ``max_sum`` the marginals are
                   max-marginals (their argmax is the MAP state when it is unique) and ``log_z`` is
                   the log-probability of the best joint state; ``log_sum_exp`` / ``max_sum`` take
                   log potentials and return log marginals
        :return: ``(marginals, log_z)``: ``{variable: array [B, size]}`` and ``log_z [B]``, the log
                 of the partition function P(evidence) of every instance -- the quantity the
                 reference computes at the root and discards (``computation.py:90-96``)
        """
        t = eng.require_cuda()
        fg = self.clique_tree.factor_graph
        evidence_vars = list(evidence_vars)
        full = dict(fg.sizes)
        full.update(_effective_sizes(fg.factors, xs))
        eff = dict(full)
        for v in evidence_vars:
            eff[v] = 1
        if variables is None:
            seen = []
            for fv in fg.factors:
                for v in fv:
                    if v not in seen and v not in evidence_vars:
                        seen.append(v)
            variables = seen
        variables = list(variables)
        lik_vars = _likelihood_vars(fg, likelihoods)
        engine = self._engine(eff, evidence_vars, full, outputs=[[v] for v in variables], likelihood_vars=lik_vars)
        plan = engine.plan
        dtype = np.dtype(dtype) if dtype is not None else _result_dtype(xs)
        B = int(evidence.shape[0]) if evidence is not None else int(batch) if batch is not None else \
            int(np.shape(likelihoods[lik_vars[0]])[0]) if lik_vars else None
        if B is None:
            raise ValueError("batch size unknown: give evidence, likelihoods or batch=")
        fdev, batched = engine.factors_to_device(xs, dtype)
        ev_host = None
        if plan.evidence_vars:
            if evidence is None:
                raise ValueError("the plan has evidence variables %r but no evidence was given" % (plan.evidence_vars,))
            ev_host = t.from_numpy(np.ascontiguousarray(evidence, dtype=np.int32)).pin_memory()
            if tuple(ev_host.shape) != (B, len(plan.evidence_vars)):
                raise ValueError("evidence must have shape [%d, %d]" % (B, len(plan.evidence_vars)))
        pipe = engine.pipeline(B, dtype, chunk=self._chunk_for(engine, B, dtype), normalize=normalize, log_z=True,
                               semiring=_semiring(dl))
        out_host = pipe.host_output()
        lik_host = engine.likelihoods_host(likelihoods, B, dtype, pin=True) if lik_vars else None
        pipe.run(fdev, False, ev_host, out_host, sync=True, lik_host=lik_host)
        if ev_host is not None and pipe.evidence_errors():
            raise ValueError("evidence states outside the range of their variable")
        views = pipe.factor_views(out_host)
        return dict(zip(variables, views)), pipe.host_logz.numpy()

    def marginals_session(self, xs, batch, variables=None, evidence_vars=(), dtype=None, normalize=True, dl=None,
                          likelihood_vars=(), chunk=None):
        """A reusable :class:`MarginalsSession` for serving: everything ``marginals_batch`` sets up
        per call -- plan, factor tables on the device, chunk workspaces (sparse when that saves
        memory), streams, pinned staging buffers -- is built once for batches of ``batch``
        instances; ``session.run(evidence, likelihoods)`` then only streams the batch."""
        return MarginalsSession(self, xs, batch, variables, evidence_vars, dtype, normalize, dl, likelihood_vars,
                                chunk=chunk)

    def propagate_session(self, xs, batch, evidence_vars=(), dtype=None, dl=None, likelihood_vars=(), chunk=None):
        """The same for ``propagate_batch``: ``session.run(evidence)`` returns the per-factor
        beliefs ``[B, *factor_shape]`` of a batch (host evidence in, host arrays out)."""
        return MarginalsSession(self, xs, batch, None, evidence_vars, dtype, False, dl, likelihood_vars,
                                factor_scopes=True, chunk=chunk)

    @staticmethod
    def _chunk_for(engine, B, dtype):
        """Instances per pipeline chunk: at most 8192, and a third of the free device memory."""
        t = eng.torch()
        per_instance = engine.pipeline_bytes_per_instance(dtype)
        free, _ = t.cuda.mem_get_info()
        chunk = int(max(1, min(8192, B, (free // 3) // max(per_instance, 1))))
        return max(2, chunk - chunk % 2) if chunk > 1 else 1

    def _propagate_streamed(self, engine, fdev, evidence, B, dtype, semiring=0, likelihoods=None):
        """Host-in / host-out propagation of a large batch through ``engine.BatchPipeline``."""
        t = eng.require_cuda()
        plan = engine.plan
        pipe = engine.pipeline(B, dtype, chunk=self._chunk_for(engine, B, dtype), semiring=semiring)
        ev_host = None
        if plan.evidence_vars:
            if evidence is None:
                raise ValueError("the plan has evidence variables %r but no evidence was given"
                                 % (plan.evidence_vars,))
            ev_host = t.from_numpy(np.ascontiguousarray(evidence, dtype=np.int32))
            if tuple(ev_host.shape) != (B, len(plan.evidence_vars)):
                raise ValueError("evidence must have shape [%d, %d], got %s"
                                 % (B, len(plan.evidence_vars), tuple(ev_host.shape)))
            ev_host = ev_host.pin_memory()
        out_host = pipe.host_output()
        lik_host = engine.likelihoods_host(likelihoods, B, dtype, pin=True) if plan.likelihood_vars else None
        pipe.run(fdev, False, ev_host, out_host, sync=True, lik_host=lik_host)
        if ev_host is not None:
            bad = pipe.evidence_errors()
            if bad:
                raise ValueError("%d evidence states are outside the range of their variable" % bad)
        return pipe.factor_views(out_host)

    def propagate_evidence(self, xs, evidence, variables=None, dtype=None, dl=None):
        """Batched ``apply_evidence`` + ``propagate`` with a different evidence *pattern* per
        instance (the batched evidence front-end).

        :param xs: factor tables shared by the batch (stored shapes)
        :param evidence: one ``{variable: state}`` dict per instance -- the argument of the
                         reference's ``apply_evidence`` (``computation.py:11-34``) -- or, with
                         ``variables`` given, an int array ``[B, len(variables)]`` in which a
                         negative entry means "not observed in this instance"
        :return: a list with one entry per instance: the list of per-factor arrays that
                 ``tree.propagate([a[0] for a in apply_evidence(xs, factors, evidence[b])])``
                 returns in the reference (observed axes have length 1)

        Instances are grouped by the set of observed variables; each group is one
        ``propagate_batch`` call (its own compiled plan, cached per pattern), so the work per
        instance is that of the sliced network.  The per-instance arrays are views of the
        group results.
        """
        fg = self.clique_tree.factor_graph
        rank = {}
        for fv in fg.factors:
            for v in fv:
                rank.setdefault(v, len(rank))
        groups = {}
        if variables is not None:
            variables = list(variables)
            for v in variables:
                if v not in rank:
                    raise ValueError("evidence on unknown variable %r" % (v,))
            table = np.asarray(evidence)
            if table.ndim != 2 or table.shape[1] != len(variables):
                raise ValueError("evidence must have shape [B, %d]" % len(variables))
            B = table.shape[0]
            masks, inverse = np.unique(table >= 0, axis=0, return_inverse=True)
            inverse = np.asarray(inverse).reshape(-1)
            for g, mask in enumerate(masks):
                rows = np.nonzero(inverse == g)[0]
                cols = [j for j in np.nonzero(mask)[0].tolist()]
                cols.sort(key=lambda j: rank[variables[j]])
                groups[tuple(variables[j] for j in cols)] = (rows, np.ascontiguousarray(table[rows][:, cols], np.int32))
        else:
            evidence = list(evidence)
            B = len(evidence)
            members = {}
            for b, ev in enumerate(evidence):
                for v in ev:
                    if v not in rank:
                        raise ValueError("evidence on unknown variable %r" % (v,))
                members.setdefault(tuple(sorted(ev, key=rank.__getitem__)), []).append(b)
            for key, rows in members.items():
                states = np.asarray([[evidence[b][v] for v in key] for b in rows], np.int32).reshape(len(rows), len(key))
                groups[key] = (np.asarray(rows), states)
        result = [None] * B
        for key, (rows, states) in groups.items():
            outs = self.propagate_batch(xs, list(key), states if key else None, batch=len(rows), dtype=dtype, dl=dl)
            for j, b in enumerate(rows.tolist()):
                result[b] = [o[j] for o in outs]
        return result

    def propagate_batch(self, xs, evidence_vars=(), evidence=None, batch=None, dtype=None,
                        nodes=False, device_output=False, uniform=True, dl=None, likelihoods=None, dense=True):
        """Many independent propagations over this tree in one pass.

        :param xs: factor tables shared by the whole batch (stored shapes, observed axes at full
                   size), or per-instance tables with a leading batch axis ``[B, *shape]``
        :param evidence_vars: variables observed in every instance
        :param evidence: int array ``[B, len(evidence_vars)]`` of observed states; equivalent to
                         ``apply_evidence`` / slicing ``e:e+1`` per instance
        :param batch: batch size when neither ``evidence`` nor batched ``xs`` determine it
        :param nodes: also return the clique and separator beliefs (node order
                      ``maxcliques + separators``)
        :param device_output: return CUDA tensors (views of the batch-innermost buffers) instead
                              of NumPy arrays
        :param uniform: with shared tables, compute potentials and up-messages that no evidence
                        reaches once per batch instead of once per instance (same results)
        :param dense: in uniform mode, contract shared potentials with a single per-instance
                      message on the FP64 tensor pipe (float64 sum-product, B >= 128; same results
                      to rounding); ``False`` keeps every task on the projection kernels
        :param dl: distributive law (``semirings.py``); default sum-product
        :param likelihoods: soft evidence ``{variable: array [B, size]}``: a per-instance likelihood
                            vector multiplied into the model, i.e. one more single-variable factor
                            per instance (log-likelihoods for the log-domain laws).  A one-hot
                            vector is hard evidence without slicing the axis.
        :return: list of ``[B, *factor_shape]`` arrays (observed axes have length 1); with
                 ``nodes=True`` a pair ``(factor_outputs, node_beliefs)``
        """
        t = eng.require_cuda()
        fg = self.clique_tree.factor_graph
        evidence_vars = list(evidence_vars)
        per_instance = bool(xs) and np.ndim(xs[0]) == len(fg.factors[0]) + 1
        full = dict(fg.sizes)
        full.update(_effective_sizes(fg.factors, xs, batched=per_instance))
        eff = dict(full)
        for v in evidence_vars:
            eff[v] = 1
        lik_vars = _likelihood_vars(fg, likelihoods)
        engine = self._engine(eff, evidence_vars, full, likelihood_vars=lik_vars)
        plan = engine.plan
        dtype = np.dtype(dtype) if dtype is not None else _result_dtype(xs)
        if evidence is not None:
            B = int(evidence.shape[0])
        elif lik_vars:
            B = int(np.shape(likelihoods[lik_vars[0]])[0])
        elif per_instance:
            B = int(np.shape(xs[0])[0])
        elif batch is not None:
            B = int(batch)
        else:
            raise ValueError("batch size unknown: give evidence, batched tables or batch=")
        fdev, batched = engine.factors_to_device(xs, dtype, B)
        if not nodes and not device_output and not batched and B > _STREAM_THRESHOLD:
            # large batches with host output: stream chunks over two CUDA streams, sized to the
            # free device memory (config 5 needs ~80 MB of workspace per instance)
            return self._propagate_streamed(engine, fdev, evidence, B, dtype, _semiring(dl), likelihoods)
        edev = engine.evidence_to_device(evidence, B)
        # node beliefs handed out as device tensors are views of the workspace: give such a call
        # a workspace of its own (kept alive by the views) instead of the engine's cached one,
        # which the next call with the same batch size overwrites
        ws = engine.new_workspace(B, dtype) if (nodes and device_output) else engine.workspace(B, dtype)
        engine.load_likelihoods(ws, B, dtype, likelihoods)
        ws, fout = engine.propagate(fdev, batched, edev, B, dtype, ws=ws, sep_beliefs=nodes, uniform=uniform,
                                    beliefs=nodes, semiring=_semiring(dl), dense=dense)
        if edev is not None:
            bad = engine.dev.evidence_errors(B, dtype, ws.data_ptr(), engine._stream())
            if bad:
                engine.clear_evidence_errors(ws, B, dtype)
                raise ValueError("%d evidence states are outside the range of their variable" % bad)
        outs = [engine.factor_tensor(fout, f, B) for f in range(len(plan.fout_off))]
        node_out = None
        if nodes:
            node_out = [engine.node_tensor(ws, k, B, dtype) for k in range(len(plan.node_vars))]
        if not device_output:
            outs = [o.cpu().numpy() for o in outs]
            if nodes:
                node_out = [o.cpu().numpy() for o in node_out]
        return (outs, node_out) if nodes else outs


class MarginalsSession:
    """Posterior marginals for repeated batches of one size over one tree (see
    ``JunctionTree.marginals_session``).  ``run`` returns fresh arrays; ``close`` (or leaving
    the ``with`` block) frees the device and pinned memory."""

    def __init__(self, tree, xs, batch, variables=None, evidence_vars=(), dtype=None, normalize=True, dl=None,
                 likelihood_vars=(), factor_scopes=False, chunk=None):
        t = eng.require_cuda()
        fg = tree.clique_tree.factor_graph
        self.evidence_vars = list(evidence_vars)
        full = dict(fg.sizes)
        full.update(_effective_sizes(fg.factors, xs))
        eff = dict(full)
        for v in self.evidence_vars:
            eff[v] = 1
        self.factor_scopes = bool(factor_scopes)       # outputs: the factor scopes (propagate) or single variables
        if variables is None and not self.factor_scopes:
            variables = []
            for fv in fg.factors:
                for v in fv:
                    if v not in variables and v not in self.evidence_vars:
                        variables.append(v)
        self.variables = None if self.factor_scopes else list(variables)
        self.likelihood_vars = _likelihood_vars(fg, dict.fromkeys(likelihood_vars)) if likelihood_vars else []
        self.engine = tree._engine(eff, self.evidence_vars, full,
                                   outputs=None if self.factor_scopes else [[v] for v in self.variables],
                                   likelihood_vars=self.likelihood_vars)
        self.B = int(batch)
        self.dtype = np.dtype(dtype) if dtype is not None else _result_dtype(xs)
        self.fdev, _ = self.engine.factors_to_device(xs, self.dtype)
        self.pipe = self.engine.pipeline(self.B, self.dtype,
                                         chunk=chunk or tree._chunk_for(self.engine, self.B, self.dtype),
                                         normalize=normalize and not self.factor_scopes,
                                         log_z=not self.factor_scopes, semiring=_semiring(dl))
        self.out_host = self.pipe.host_output()
        self._ran = False
        n_ev = len(self.engine.plan.evidence_vars)
        self.ev_host = t.zeros((self.B, n_ev), dtype=t.int32).pin_memory() if n_ev else None
        self.lik_host = None
        if self.likelihood_vars:
            self.lik_host = t.zeros((self.engine.plan.lik_entries, self.B),
                                    dtype=eng.torch_dtype(self.dtype)).pin_memory()

    def run(self, evidence=None, likelihoods=None, copy=True):
        """``(marginals, log_z)`` as ``JunctionTree.marginals_batch`` returns them, or -- for a
        ``propagate_session`` -- the list of per-factor beliefs of ``propagate_batch``.
        ``copy=False`` returns views of the session's pinned result buffer, valid until the next
        ``run`` (no host copy of the results: the call is then bound by the PCIe transfer)."""
        if self.pipe is None:
            raise RuntimeError("the session is closed")
        if self.ev_host is not None:
            ev = np.asarray(evidence)
            if ev.shape != tuple(self.ev_host.shape):
                raise ValueError("evidence must have shape %s" % (tuple(self.ev_host.shape),))
            self.ev_host.numpy()[...] = ev
        if self.lik_host is not None:
            self.lik_host.copy_(self.engine.likelihoods_host(likelihoods, self.B, self.dtype))
        # the factor tables of a session never change: the uniform workspaces are computed by the first run
        self.pipe.run(self.fdev, False, self.ev_host, self.out_host, sync=True, lik_host=self.lik_host,
                      same_tables=self._ran)
        self._ran = True
        if self.ev_host is not None and self.pipe.evidence_errors():
            self.close()           # the error counters live in the workspaces: start from clean ones next time
            raise ValueError("evidence states outside the range of their variable")
        views = self.pipe.factor_views(self.out_host)
        if copy:
            views = [np.array(a) for a in views]
        if self.factor_scopes:
            return views
        log_z = self.pipe.host_logz.numpy()
        return dict(zip(self.variables, views)), log_z.copy() if copy else log_z

    def close(self):
        self.pipe = self.out_host = self.ev_host = self.lik_host = self.fdev = None
        eng.torch().cuda.empty_cache()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
