"""Shapes of the dense-contraction candidates of a plan (uniform src, one streamed row per item).

    python junction-tree_b200/tools/dense_stats.py dag500 [--no-beliefs]
"""
import bisect
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import numpy as np  # noqa: E402
import jt_workloads as wl  # noqa: E402
import junctiontree as jt  # noqa: E402
from junctiontree import schedule as sch  # noqa: E402


def main():
    name = sys.argv[1]
    no_beliefs = "--no-beliefs" in sys.argv
    net = {"dag37": wl.dag37, "dag500": wl.dag500, "ising16": lambda: wl.ising(16),
           "large_state_tree": wl.large_state_tree}[name]()
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = list(net.get("evidence_vars", []))
    plan = tree.plan(evars)
    T, M, L = plan.tasks_arr, plan.msgs_arr, plan.launches_arr
    seps = list(range(plan.n_cliques, plan.n_cliques + plan.n_seps))
    sep_rel = [plan.node_off[s] - plan.clique_entries for s in seps]

    def sep_of(off):
        rel = off - plan.up_base if off < plan.down_base else off - plan.down_base
        return seps[bisect.bisect_right(sep_rel, rel) - 1]

    main_phase = sch.PHASE_DIST_MAIN_MESSAGES if no_beliefs else sch.PHASE_DIST_MAIN
    phases = [sch.PHASE_COLLECT_INSTANCE, sch.PHASE_DIST_PRE_INSTANCE, main_phase]
    if no_beliefs:
        phases.append(sch.PHASE_MARGINAL_DIRECT)
    rows = []
    for ph, b, e, lvl in L:
        if ph not in phases:
            continue
        for t in range(b, e):
            row = T[t]
            fl = int(row[sch.T_FLAGS])
            if not (fl & sch.TF_SRC_UNIFORM):
                continue
            if row[sch.T_BETA] >= 0 and not no_beliefs:
                continue
            rm = [M[j] for j in range(row[sch.T_RMSG_BEGIN], row[sch.T_RMSG_END]) if not M[j][sch.M_UNI]]
            smm = [M[j] for j in range(row[sch.T_SMSG_BEGIN], row[sch.T_SMSG_END]) if not M[j][sch.M_UNI]]
            if len(rm) != 1:
                continue
            c = int(row[sch.T_NODE])
            msep = sep_of(int(rm[0][sch.M_OFF]))
            mv = set(plan.node_vars[msep])
            cv = plan.node_vars[c]
            n_s, n_r = int(row[sch.T_NS]), int(row[sch.T_NR])
            # s vars: recover from the output node: collect -> parent sep; dist -> child sep; marginal -> scope
            if ph == sch.PHASE_MARGINAL_DIRECT:
                sv = plan.out_scopes[int(row[sch.T_AUX])]
            else:
                sv = plan.node_vars[sep_of(int(row[sch.T_OUT]))]
            rv = [v for v in cv if v not in sv]
            n_g = int(np.prod([plan.sizes[v] for v in sv if v in mv] or [1]))
            n_i = n_s // n_g
            K = int(np.prod([plan.sizes[v] for v in rv if v in mv] or [1]))
            n_q = n_r // K
            n_m = plan.node_size[msep]
            rows.append((n_s * n_r, n_g, n_i, K, n_q, n_m, n_s, len(smm)))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    moved = sum(r[5] + r[6] for r in rows)
    print("%s: %d dense candidates, %d items -> %d rows moved (x%.1f less)" % (name, len(rows), tot, moved, tot / max(moved, 1)))
    print("%12s %8s %8s %6s %6s %10s %10s %5s" % ("items", "groups", "i", "K", "q", "M rows", "out rows", "srows"))
    for r in rows[:25]:
        print("%12d %8d %8d %6d %6d %10d %10d %5d" % r)
    # distribution of flops by (i, K) bucket
    import collections
    bucket = collections.Counter()
    for it, g, i, K, q, nm, ns, _ in rows:
        bucket[("i>=16" if i >= 16 else "i>=4" if i >= 4 else "i<4", "K>=8" if K >= 8 else "K>=4" if K >= 4 else "K<4")] += it
    for k, v in sorted(bucket.items(), key=lambda kv: -kv[1]):
        print(k, "%.1f%%" % (100.0 * v / tot))


if __name__ == "__main__":
    main()
