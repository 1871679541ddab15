O=gpurun_out/r2g2
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
P="python junction-tree_b200/tools/prof_step.py"
timeout 300 $P --config ising16 --batch 256 --no-uniform >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config dag37 --batch 65536 --no-uniform >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config dag37 --batch 65536 >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config dag500 --batch 1024 --no-uniform >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config large_state_tree --batch 512 --dtype f32 --no-uniform >> $O/steps.jsonl 2>> $O/steps.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g2/steps*.jsonl")):
    for line in open(f):
        d=json.loads(line)
        print("  %-18s %s B=%-6d uni=%d ms=%.3f init=%.3f frac=%.3f"%(d["config"],d["dtype"],d["batch"],d["uniform"],d["ms_per_step"],d["init_ms"],d["scheduled_frac"]))
PY
tail -3 $O/steps.err
