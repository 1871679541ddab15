mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q 2>&1 | tail -15 > gpurun_out/r2b/dense_tests.txt
cat gpurun_out/r2b/dense_tests.txt
P="python junction-tree_b200/tools/prof_step.py"
for cfg in "large_state_tree 512" "dag500 1024" "dag37 65536"; do
  set -- $cfg
  for mode in "" "--no-dense"; do
    timeout 300 $P --config $1 --batch $2 $mode >> gpurun_out/r2b/steps.jsonl 2>> gpurun_out/r2b/steps.err
  done
done
for mode in "" "--no-dense"; do
  timeout 300 $P --config dag500 --batch 4096 --no-beliefs $mode >> gpurun_out/r2b/steps.jsonl 2>> gpurun_out/r2b/steps.err
  timeout 300 $P --config dag37 --batch 65536 --no-beliefs $mode >> gpurun_out/r2b/steps.jsonl 2>> gpurun_out/r2b/steps.err
done
cat gpurun_out/r2b/steps.jsonl
tail -5 gpurun_out/r2b/steps.err
