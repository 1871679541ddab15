O=gpurun_out/r2n
mkdir -p $O
(time timeout 1500 python -m pytest tests/test_gpu_dense.py tests/test_gpu_parity.py tests/test_gpu_reference_behaviour.py -m gpu -x -q 2>&1 | tail -8) > $O/gpu_tests.txt 2>&1
cat $O/gpu_tests.txt
P="python junction-tree_b200/tools/prof_step.py"
for cfg in "large_state_tree 512 f64" "dag500 1024 f64" "dag500 2048 f64" "dag37 65536 f64" "ising16 256 f64"; do
  set -- $cfg
  timeout 300 $P --config $1 --batch $2 --dtype $3 --compare >> $O/steps.jsonl 2>> $O/steps.err
done
JT_LEVEL_STREAMS=0 timeout 300 $P --config dag500 --batch 1024 --compare >> $O/steps_nofork.jsonl 2>> $O/steps.err
timeout 300 $P --config dag500 --batch 8192 --no-beliefs --compare >> $O/steps.jsonl 2>> $O/steps.err
python - <<'PY'
import json
for f in ("steps","steps_nofork"):
    print(f)
    for line in open("gpurun_out/r2n/%s.jsonl"%f):
        d=json.loads(line)
        print("  %-18s B=%-6d beliefs=%-5s ms=%.3f again=%.3f no_dense=%.3f  frac=%.3f launches=%.0f"%(d["config"],d["batch"],d["beliefs"],d["ms_per_step"],d["ms_per_step_again"] or 0,d["ms_per_step_no_dense"] or 0,d["scheduled_frac"],d["launches_per_step"]))
PY
tail -5 $O/steps.err
