O=gpurun_out/r2v
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
timeout 300 $P --config dag37 --batch 65536 --compare >> $O/steps.jsonl 2>> $O/steps.err
JT_BETA_MAX_B=1000000 timeout 300 $P --config dag37 --batch 65536 --compare >> $O/steps_maxb.jsonl 2>> $O/steps.err
JT_DENSE_MIN_GAIN=2 timeout 300 $P --config dag37 --batch 65536 --compare >> $O/steps_gain2.jsonl 2>> $O/steps.err
JT_BETA_MAX_B=1000000 JT_DENSE_MIN_GAIN=2 timeout 300 $P --config dag37 --batch 65536 --compare >> $O/steps_both.jsonl 2>> $O/steps.err
JT_DENSE_MIN_GAIN=2 timeout 300 $P --config dag500 --batch 2048 >> $O/steps_gain2.jsonl 2>> $O/steps.err
JT_DENSE_MIN_GAIN=4 timeout 300 $P --config dag500 --batch 2048 >> $O/steps_gain4.jsonl 2>> $O/steps.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2v/steps*.jsonl")):
    print(f)
    for line in open(f):
        d=json.loads(line)
        print("  %-18s %s B=%-6d ms=%.3f no_dense=%s frac=%.3f"%(d["config"],d["dtype"],d["batch"],d["ms_per_step"],d.get("ms_per_step_no_dense"),d["scheduled_frac"]))
PY
tail -3 $O/steps.err
