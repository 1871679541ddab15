O=gpurun_out/r2q
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
for cfg in "ising16 256 f64 --compare" "ising16 256 f64 --no-uniform" "large_state_tree 512 f64 --no-uniform" "large_state_tree 512 f32 --no-uniform" "large_state_tree 512 f32 --compare" "dag500 1024 f64 --compare"; do
  set -- $cfg
  timeout 300 $P --config $1 --batch $2 --dtype $3 $4 >> $O/steps.jsonl 2>> $O/steps.err
  JT_TMA_SMALL_VPT=1 timeout 300 $P --config $1 --batch $2 --dtype $3 $4 >> $O/steps_vpt.jsonl 2>> $O/steps.err
done
for cfg in "dag500 2048" "ising16 256" "large_state_tree 512"; do
  set -- $cfg
  timeout 300 $P --config $1 --batch $2 --uniform-valid >> $O/steps_uv.jsonl 2>> $O/steps.err
done
python - <<'PY'
import json
for f in ("steps","steps_vpt","steps_uv"):
    print(f)
    for line in open("gpurun_out/r2q/%s.jsonl"%f):
        d=json.loads(line)
        print("  %-18s %s B=%-6d uniform=%-5s ms=%.3f no_dense=%s uniform_valid=%s frac=%.3f"%(d["config"],d["dtype"],d["batch"],d["uniform"],d["ms_per_step"],d.get("ms_per_step_no_dense"),d.get("ms_uniform_valid"),d["scheduled_frac"]))
PY
tail -5 $O/steps.err
