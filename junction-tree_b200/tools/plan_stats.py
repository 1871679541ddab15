"""Where the (s, r) items of a plan go in uniform mode, by task category (host only, no GPU).

    python junction-tree_b200/tools/plan_stats.py dag500 [--no-beliefs]

Category = (src uniform?, streamed per-item rows, writes beta?).  The TMA projection kernel
streams `rows` batch rows per item from L2/HBM, so items x rows is its cost; a task with a
uniform src and exactly one streamed row is a dense contraction in disguise.
"""
import collections
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import jt_workloads as wl  # noqa: E402
import junctiontree as jt  # noqa: E402
from junctiontree import schedule as sch  # noqa: E402


def main():
    name = sys.argv[1]
    no_beliefs = "--no-beliefs" in sys.argv
    net = {"dag37": wl.dag37, "dag500": wl.dag500, "ising16": lambda: wl.ising(16),
           "large_state_tree": wl.large_state_tree}[name]()
    tree = jt.create_junction_tree(net["factors"], net["sizes"], order=net.get("order"))
    evars = [] if "--no-evidence" in sys.argv else list(net.get("evidence_vars", []))
    plan = tree.plan(evars)
    T, M, L = plan.tasks_arr, plan.msgs_arr, plan.launches_arr
    main_phase = sch.PHASE_DIST_MAIN_MESSAGES if no_beliefs else sch.PHASE_DIST_MAIN
    phases = {sch.PHASE_COLLECT_INSTANCE: "collect", sch.PHASE_DIST_PRE_INSTANCE: "dist_pre", main_phase: "dist_main"}
    if no_beliefs:
        phases[sch.PHASE_MARGINAL_DIRECT] = "marginal"
    cat = collections.Counter()
    rows_cat = collections.Counter()
    n_tasks = collections.Counter()
    per_launch = []
    for ph, b, e, lvl in L:
        if ph not in phases:
            continue
        items_l = 0
        for t in range(b, e):
            row = T[t]
            fl = int(row[sch.T_FLAGS])
            src_uni = bool(fl & sch.TF_SRC_UNIFORM) or row[sch.T_SRC] < 0
            n_s, n_r = int(row[sch.T_NS]), int(row[sch.T_NR])
            rm = [M[j] for j in range(row[sch.T_RMSG_BEGIN], row[sch.T_RMSG_END])]
            smm = [M[j] for j in range(row[sch.T_SMSG_BEGIN], row[sch.T_SMSG_END])]
            r_rows = sum(1 for m in rm if not m[sch.M_UNI]) + (0 if src_uni else 1)
            s_rows = sum(1 for m in smm if not m[sch.M_UNI])
            writer = row[sch.T_BETA] >= 0 and not no_beliefs
            key = (phases[ph], "src_uni" if src_uni else "src_row", r_rows, s_rows, "beta" if writer else "-")
            cat[key] += n_s * n_r
            rows_cat[key] += n_s * n_r * (r_rows + (1 if writer else 0)) + n_s * (s_rows + 1)
            n_tasks[key] += 1
            items_l += n_s * n_r
        per_launch.append((phases[ph], lvl, e - b, items_l))
    total = sum(cat.values())
    print("%s: %d cliques, clique entries %d (uniform %d), sep entries %d, items in instance launches %d" % (
        name, plan.n_cliques, plan.clique_entries, plan.uni_entries, plan.sep_entries, total))
    print("%-52s %8s %14s %7s %16s" % ("category (phase, src, item rows, s rows, beta)", "tasks", "items", "share", "rows moved"))
    for key, n in sorted(cat.items(), key=lambda kv: -kv[1]):
        print("%-52s %8d %14d %6.1f%% %16d" % (str(key), n_tasks[key], n, 100.0 * n / total, rows_cat[key]))
    print("launches: %d; items per launch min/median/max: %s" % (
        len(per_launch), [sorted(x[3] for x in per_launch)[i] for i in (0, len(per_launch) // 2, -1)]))
    print("scheduled entries/instance (uniform): %d" % plan.scheduled_entries(uniform=True))


if __name__ == "__main__":
    main()
