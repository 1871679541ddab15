O=gpurun_out/r2d2
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
for m in 8 16 32 64; do
for c in "dag37 65536" "dag500 2048" "ising16 256" "large_state_tree 512"; do set -- $c
JT_TMA_CTAS_PER_SM=$m timeout 300 $P --config $1 --batch $2 >> $O/steps_cps$m.jsonl 2>> $O/steps.err
done
JT_TMA_CTAS_PER_SM=$m timeout 300 $P --config dag37 --batch 65536 --no-uniform >> $O/steps_cps$m.jsonl 2>> $O/steps.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d2/steps*.jsonl")):
    print(f)
    for line in open(f):
        d=json.loads(line)
        print("  %-18s %s B=%-6d uni=%d ms=%.3f frac=%.3f"%(d["config"],d["dtype"],d["batch"],d["uniform"],d["ms_per_step"],d["scheduled_frac"]))
PY
tail -3 $O/steps.err
