"""Summarise an ncu launch list (the `--metrics gpu__time_duration.sum,dram__bytes_*` pass of
B200_PROFILING.md over `bench.py --steps 2 --warmup 1 --skip-cpu`) per launch and per step.

    python junction-tree_b200/tools/summarize_launches.py profiles/r01c_launches_uniform.csv [--step 3]
        [--update profiles/r01_traffic.json --key dag37:65536:f64:uniform]

A step is delimited by two consecutive launches of jt_evidence_kernel.  `--update` rewrites one
entry of the traffic table that bench.py reads for `roofline.traffic` (DRAM bytes per launch of
the message-passing kernels -- jt_project_tma_kernel, jt_dense_kernel, jt_beta_kernel, jt_scalar_kernel --
averaged over those launches of that step).
"""
import argparse
import collections
import csv
import io
import json


def load(path):
    lines = open(path).read().splitlines()
    start = next(i for i, line in enumerate(lines) if line.startswith('"ID"'))
    launches = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        d = launches.setdefault(int(row["ID"]), {"name": row["Kernel Name"], "grid": row["Grid Size"],
                                                 "block": row["Block Size"]})
        d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    return launches


def short(name):
    name = name.split("::")[1] if "::" in name else name
    return name.split("(")[0].split("<")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--step", type=int, default=3, help="which step of the run (0 = first warm-up step)")
    ap.add_argument("--update")
    ap.add_argument("--key")
    args = ap.parse_args()
    launches = load(args.csv)
    ids = list(launches)
    marks = [i for i in ids if "jt_evidence_kernel" in launches[i]["name"]]
    if not marks:       # no evidence variables: a step starts with the first init launch after a non-init one
        prev_init = False
        for i in ids:
            is_init = "jt_init" in launches[i]["name"]
            if is_init and not prev_init:
                marks.append(i)
            prev_init = is_init
    lo = marks[args.step]
    hi = marks[args.step + 1] if args.step + 1 < len(marks) else ids[-1] + 1
    print("| id | kernel | grid | time (us) | dram read (GB) | dram write (GB) | dram GB/s | issue active |")
    print("|---|---|---|---|---|---|---|---|")
    total_t = total_b = 0.0
    proj_t = proj_b = 0.0
    proj_n = 0
    for i in ids:
        if not lo <= i < hi:
            continue
        d = launches[i]
        t = d.get("gpu__time_duration.sum", 0.0) / 1e3
        rb, wb = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        issue = d.get("smsp__issue_active.avg.pct_of_peak_sustained_active")
        total_t += t
        total_b += rb + wb
        # the batch launches of collect + distribute (the B = 1 launches of the uniform workspace use
        # jt_project_kernel / jt_project_splitr_kernel / jt_dense_prep_kernel)
        if any(k in d["name"] for k in ("jt_project_tma_kernel", "jt_dense_kernel", "jt_beta_kernel", "jt_scalar_kernel")):
            proj_t += t
            proj_b += rb + wb
            proj_n += 1
        print("| %d | %s | %s | %.1f | %.3f | %.3f | %.0f | %s |" % (
            i, short(d["name"]), d["grid"], t, rb / 1e9, wb / 1e9, (rb + wb) / 1e3 / t if t else 0.0,
            "%.0f %%" % issue if issue is not None else "-"))
    print()
    print("step: %.1f us under ncu, %.2f GB of DRAM traffic" % (total_t, total_b / 1e9))
    print("batch projection launches: %d, %.1f us (%.1f %% of the step), %.2f GB, %.0f GB/s" % (
        proj_n, proj_t, 100.0 * proj_t / total_t, proj_b / 1e9, proj_b / 1e3 / proj_t if proj_t else 0.0))
    if args.update:
        try:
            table = json.load(open(args.update))
        except Exception:
            table = {}
        table[args.key] = {"launches_per_step": proj_n, "dram_bytes_per_step": int(proj_b),
                           "traffic_per_launch": int(proj_b / max(proj_n, 1)),
                           "ncu_time_us_per_step": round(proj_t, 1), "source": args.csv}
        json.dump(table, open(args.update, "w"), indent=1)


if __name__ == "__main__":
    main()
