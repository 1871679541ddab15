O=gpurun_out/r2p
mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/gpu_tests.txt 2>&1
cat $O/gpu_tests.txt
P="python junction-tree_b200/tools/prof_step.py"
timeout 300 $P --config large_state_tree --batch 512 --dtype f32 --compare >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config large_state_tree --batch 512 --dtype f64 --compare >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config dag37 --batch 65536 --dtype f32 --compare >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config dag500 --batch 2048 --dtype f32 --compare >> $O/steps.jsonl 2>> $O/steps.err
python - <<'PY'
import json
for line in open("gpurun_out/r2p/steps.jsonl"):
    d=json.loads(line)
    print("  %-18s %s B=%-6d beliefs=%-5s ms=%.3f again=%.3f no_dense=%.3f  frac=%.3f launches=%.0f"%(d["config"],d["dtype"],d["batch"],d["beliefs"],d["ms_per_step"],d["ms_per_step_again"] or 0,d["ms_per_step_no_dense"] or 0,d["scheduled_frac"],d["launches_per_step"]))
PY
tail -5 $O/steps.err
