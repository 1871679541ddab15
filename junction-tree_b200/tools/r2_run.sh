O=gpurun_out/r2a2
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
for g in 2 1.5 1.25; do
for c in "dag37 65536" "dag500 2048" "ising16 256" "large_state_tree 512"; do set -- $c
JT_DENSE_MIN_GAIN=$g timeout 300 $P --config $1 --batch $2 >> $O/steps_gain$g.jsonl 2>> $O/steps.err
done
JT_DENSE_MIN_GAIN=$g timeout 300 $P --config dag500 --batch 4096 --no-beliefs >> $O/steps_gain$g.jsonl 2>> $O/steps.err
JT_DENSE_MIN_GAIN=$g timeout 300 $P --config dag37 --batch 65536 --no-beliefs >> $O/steps_gain$g.jsonl 2>> $O/steps.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2a2/steps*.jsonl")):
    print(f)
    for line in open(f):
        d=json.loads(line)
        print("  %-18s %s B=%-6d bel=%d ms=%.3f frac=%.3f"%(d["config"],d["dtype"],d["batch"],d["beliefs"],d["ms_per_step"],d["scheduled_frac"]))
PY
tail -3 $O/steps.err
