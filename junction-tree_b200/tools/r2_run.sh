O=gpurun_out/r2f
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv"
for cfg in "large_state_tree 512" "dag500 1024" "ising16 128"; do
  set -- $cfg
  timeout 900 $NCU --log-file $O/launches_$1.csv $P --config $1 --batch $2 --steps 1 --warmup 1 > /dev/null 2>&1
done
ls -la $O
