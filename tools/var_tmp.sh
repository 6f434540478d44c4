mkdir -p gpurun_out/r2
for n in 8 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2/bench_n$n.err | grep '^{' | tail -1 > gpurun_out/r2/r2_bench_c2_${n}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2/r2_bench_c2_${n}gpu.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['single_job_ms_per_step'], d['workload_details'].get('host_cores_bound_to_gpu_numa_node'))"
grep -i "error\|Traceback" -A5 gpurun_out/r2/bench_n$n.err | head -20
done
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2/r2_bench_c2_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2/r2_bench_c2_1gpu.json')); print(1, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['single_job_ms_per_step'], d['workload_details'].get('host_cores_bound_to_gpu_numa_node'))"
nvidia-smi topo -m | head -14; lscpu | grep -i "numa\|socket\|^CPU(s)"
