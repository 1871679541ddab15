O=gpurun_out/r02
mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/gpu_tests.txt 2>&1
cat $O/gpu_tests.txt
python bench.py --steps 10 --warmup 3 > $O/r02_bench_uniform.json 2> $O/bench_uniform.err
python bench.py --steps 10 --warmup 3 --no-uniform --configs none > $O/r02_bench_perinstance.json 2> $O/bench_perinstance.err
tail -2 $O/bench_uniform.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/r02_bench_uniform.json").read())
print("value %.0f ms %.3f frac %.3f traffic %s e2e %.0f e2e_marg %.0f"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["traffic"],d["e2e"]["value"],d["e2e_marginals"]["value"]))
for e in d["configs"]:
    print("  ", e["config"], e.get("batch_per_gpu"), e.get("dtype"), e.get("mode"), e.get("beliefs_stored"), round(e.get("ms_per_step",0),3), round(e["value"]), e.get("frac") and round(e["frac"],3))
PY
P="python junction-tree_b200/tools/prof_step.py"
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv"
JT_BENCH_SHORT_WARMUP=1 timeout 1200 $NCU --log-file $O/r02_launches_dag500_uniform.csv $P --config dag500 --batch 2048 --steps 2 --warmup 1 > /dev/null 2>&1
JT_BENCH_SHORT_WARMUP=1 timeout 600 $NCU --log-file $O/r02_launches_large_state_tree_f64_uniform.csv $P --config large_state_tree --batch 512 --steps 2 --warmup 1 > /dev/null 2>&1
ls -la $O | tail -8
