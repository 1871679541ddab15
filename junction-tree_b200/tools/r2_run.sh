O=gpurun_out/r2t
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
for blk in 8 4 16 32 64; do
for cfg in "large_state_tree 512 f64" "dag500 2048 f64"; do
  set -- $cfg
  JT_BETA_BLOCK=$blk timeout 300 $P --config $1 --batch $2 --dtype $3 >> $O/steps_$blk.jsonl 2>> $O/steps.err
done
done
JT_BETA_MIN_MB=256 timeout 300 $P --config dag500 --batch 2048 >> $O/steps_min256.jsonl 2>> $O/steps.err
JT_BETA_MIN_MB=256 timeout 300 $P --config ising16 --batch 256 --compare >> $O/steps_min256.jsonl 2>> $O/steps.err
JT_BETA_MIN_MB=64 timeout 300 $P --config ising16 --batch 256 --compare >> $O/steps_min64.jsonl 2>> $O/steps.err
JT_BETA_MIN_MB=64 timeout 300 $P --config dag500 --batch 2048 >> $O/steps_min64.jsonl 2>> $O/steps.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t/steps_*.jsonl")):
    print(f)
    for line in open(f):
        d=json.loads(line)
        print("  %-18s %s B=%-6d ms=%.3f no_dense=%s frac=%.3f"%(d["config"],d["dtype"],d["batch"],d["ms_per_step"],d.get("ms_per_step_no_dense"),d["scheduled_frac"]))
PY
tail -5 $O/steps.err
