O=gpurun_out/r2i
mkdir -p $O
(time timeout 1500 python -m pytest tests/test_gpu_dense.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15) > $O/gpu_tests.txt 2>&1
cat $O/gpu_tests.txt
P="python junction-tree_b200/tools/prof_step.py"
for cfg in "large_state_tree 512 f64" "dag500 1024 f64" "dag37 65536 f64" "ising16 128 f64"; do
  set -- $cfg
  timeout 300 $P --config $1 --batch $2 --dtype $3 >> $O/steps.jsonl 2>> $O/steps.err
  JT_LEVEL_STREAMS=0 timeout 300 $P --config $1 --batch $2 --dtype $3 >> $O/steps_nofork.jsonl 2>> $O/steps.err
  JT_DISABLE_BETA=1 timeout 300 $P --config $1 --batch $2 --dtype $3 >> $O/steps_nobeta.jsonl 2>> $O/steps.err
done
for mode in "" "--no-dense"; do
timeout 300 $P --config dag500 --batch 4096 --no-beliefs $mode >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config dag37 --batch 65536 --no-beliefs $mode >> $O/steps.jsonl 2>> $O/steps.err
timeout 300 $P --config ising16 --batch 256 --no-beliefs $mode >> $O/steps.jsonl 2>> $O/steps.err
done
echo default; cut -c1-300 $O/steps.jsonl
echo nofork; cut -c1-300 $O/steps_nofork.jsonl
echo nobeta; cut -c1-300 $O/steps_nobeta.jsonl
tail -5 $O/steps.err
