mkdir -p gpurun_out/r2u
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2u/bench_8gpu.json 2> gpurun_out/r2u/bench_8gpu.err
tail -3 gpurun_out/r2u/bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2u/bench_8gpu.json").read())
print(d["value"], d["n_gpus"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["d2h_gbs_per_gpu"], d["e2e"]["host_ceiling"], "marg", d["e2e_marginals"]["value"])
print(d["all_gather"])
print([(e["config"], e.get("mode"), round(e.get("ms_per_step",0),2), round(e["value"])) for e in d["configs"]])
PY
nvidia-smi topo -m 2>/dev/null | head -14
python -c "
import os; print('cpus', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for n in range(4):
    p='/sys/devices/system/node/node%d/cpulist'%n
    if os.path.exists(p): print(n, open(p).read().strip())
"
