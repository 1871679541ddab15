O=gpurun_out/r2h2
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
for w in 4 2 1 8; do
JT_DENSE_WAVES=$w timeout 300 $P --config dag500 --batch 2048 >> $O/steps_w$w.jsonl 2>> $O/steps.err
JT_DENSE_WAVES=$w timeout 300 $P --config dag500 --batch 4096 --no-beliefs >> $O/steps_w$w.jsonl 2>> $O/steps.err
JT_DENSE_WAVES=$w timeout 300 $P --config dag37 --batch 65536 >> $O/steps_w$w.jsonl 2>> $O/steps.err
JT_DENSE_WAVES=$w timeout 300 $P --config large_state_tree --batch 512 >> $O/steps_w$w.jsonl 2>> $O/steps.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h2/steps*.jsonl")):
    print(f)
    for line in open(f):
        d=json.loads(line)
        print("  %-18s %s B=%-6d bel=%d ms=%.3f frac=%.3f"%(d["config"],d["dtype"],d["batch"],d["beliefs"],d["ms_per_step"],d["scheduled_frac"]))
PY
tail -3 $O/steps.err
