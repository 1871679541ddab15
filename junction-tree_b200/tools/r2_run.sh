O=gpurun_out/r2o
mkdir -p $O
P="python junction-tree_b200/tools/prof_step.py"
JT_NVTX=1 JT_BENCH_SHORT_WARMUP=1 timeout 600 ncu --nvtx --nvtx-include "jt collect level 28/" --set full --clock-control none --import-source on -k regex:jt_dense_kernel -c 2 -f -o $O/dense_dag500_l28 $P --config dag500 --batch 1024 --steps 1 --warmup 1 > $O/ncu1.log 2>&1
tail -3 $O/ncu1.log
JT_NVTX=1 JT_BENCH_SHORT_WARMUP=1 timeout 600 ncu --nvtx --nvtx-include "jt collect level 29/" --set full --clock-control none --import-source on -k regex:jt_dense_kernel -c 2 -f -o $O/dense_dag500_l29 $P --config dag500 --batch 1024 --steps 1 --warmup 1 > $O/ncu2.log 2>&1
tail -3 $O/ncu2.log
ls -la $O
