set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2a
P="python junction-tree_b200/tools/prof_step.py"
O=gpurun_out/r2a
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv"
for cfg in "large_state_tree 512 f64" "large_state_tree 512 f32" "dag500 1024 f64" "ising16 128 f64" "dag37 65536 f64"; do
  set -- $cfg
  for mode in "" "--no-uniform"; do
    $P --config $1 --batch $2 --dtype $3 $mode >> $O/steps.jsonl 2>> $O/steps.err
  done
done
$P --config dag37 --batch 65536 --no-evidence >> $O/steps.jsonl 2>> $O/steps.err
$P --config dag500 --batch 4096 --no-beliefs >> $O/steps.jsonl 2>> $O/steps.err
$P --config ising16 --batch 256 --no-beliefs >> $O/steps.jsonl 2>> $O/steps.err
# launch lists (2 steps + 1 warm-up each)
for cfg in "large_state_tree 512 f64" "large_state_tree 512 f32" "dag500 1024 f64" "ising16 128 f64"; do
  set -- $cfg
  for mode in "" "--no-uniform"; do
    timeout 900 $NCU --log-file $O/launches_$1_$3$mode.csv $P --config $1 --batch $2 --dtype $3 $mode --steps 1 --warmup 1 > /dev/null 2>&1
  done
done
cat $O/steps.jsonl
