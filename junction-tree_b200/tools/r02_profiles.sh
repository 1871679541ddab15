#!/bin/bash
# Round-2 evidence run (one B200, `gpurun -- bash junction-tree_b200/tools/r02_profiles.sh [part]`):
#   part 1: bench lines (default, --no-uniform) + launch lists of the small configs + full captures
#   part 2: launch lists of the deep trees (config 3 and 5: hundreds of launches per step)
#   part 3: compute-sanitizer
# Everything lands in gpurun_out/r02/ and is copied into profiles/ by hand (see profiles/README.md).
O=gpurun_out/r02
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv"
export JT_BENCH_SHORT_WARMUP=1
list() {   # name, then prof_step arguments
  name=$1; shift
  timeout 1200 $NCU --log-file $O/r02_launches_$name.csv $P "$@" --steps 2 --warmup 1 > /dev/null 2>&1
}
part=${1:-1}
if [ "$part" = "1" ]; then
  unset JT_BENCH_SHORT_WARMUP
  (time python bench.py --steps 10 --warmup 3 > $O/r02_bench_uniform.json) 2> $O/bench_uniform.err
  python bench.py --steps 10 --warmup 3 --no-uniform --configs none > $O/r02_bench_perinstance.json 2> $O/bench_perinstance.err
  python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/bench_reference.err
  tail -2 $O/bench_uniform.err
  export JT_BENCH_SHORT_WARMUP=1
  list dag37_uniform --config dag37 --batch 65536
  list dag37_perinstance --config dag37 --batch 65536 --no-uniform
  list dag37_noevidence --config dag37 --batch 65536 --no-evidence
  list large_state_tree_f64_uniform --config large_state_tree --batch 512
  list large_state_tree_f64_perinstance --config large_state_tree --batch 512 --no-uniform
  list large_state_tree_f32_uniform --config large_state_tree --batch 512 --dtype f32
  list large_state_tree_f32_perinstance --config large_state_tree --batch 512 --dtype f32 --no-uniform
  # full captures: the dense contraction and the belief kernel (config 4), the projection kernel (config 2, per instance)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:jt_dense_kernel --launch-skip 2 -c 2 -f -o $O/r02_dense_cfg4 $P --config large_state_tree --batch 512 --steps 1 --warmup 1 > /dev/null 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:jt_beta_kernel --launch-skip 2 -c 2 -f -o $O/r02_beta_cfg4 $P --config large_state_tree --batch 512 --steps 1 --warmup 1 > /dev/null 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:jt_project_tma_kernel --launch-skip 40 -c 6 -f -o $O/r02_tma_dag37_perinstance $P --config dag37 --batch 65536 --no-uniform --steps 1 --warmup 1 > /dev/null 2>&1
  for f in r02_dense_cfg4 r02_beta_cfg4 r02_tma_dag37_perinstance; do
    ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
    rm -f $O/$f.ncu-rep
  done
fi
if [ "$part" = "2" ]; then
  list ising16_uniform --config ising16 --batch 256
  list ising16_perinstance --config ising16 --batch 256 --no-uniform
  list dag500_uniform --config dag500 --batch 2048
  list dag500_perinstance --config dag500 --batch 1024 --no-uniform
  list dag500_pipeline_mode --config dag500 --batch 4096 --no-beliefs
fi
if [ "$part" = "4" ]; then   # lighter lists (time + DRAM bytes, one timed step) of the per-instance deep trees
  NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
  list1() { name=$1; shift; timeout 600 $NCU --log-file $O/r02_launches_$name.csv $P "$@" --steps 1 --warmup 1 > /dev/null 2>&1; }
  list1 ising16_perinstance --config ising16 --batch 256 --no-uniform
  list1 dag500_perinstance --config dag500 --batch 1024 --no-uniform
fi
if [ "$part" = "3" ]; then
  timeout 900 compute-sanitizer --tool memcheck python tests/tools/sanitize_case.py 2>&1 | tail -25 > $O/r02_memcheck.txt
  cat $O/r02_memcheck.txt
fi
ls -la $O | head -40
