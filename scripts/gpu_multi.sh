# usage: gpu_multi.sh NGPUS
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_worker.py > gpurun_out/mgpu$N.log 2>&1; echo "worker rc=$?"; grep "AssertionError\| ok on\|Error" gpurun_out/mgpu$N.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_${N}gpu.json'):
    if line.startswith('{'):
        d=json.loads(line); print({k:d[k] for k in ('n_gpus','value','ms_per_step','log_ml','gpu_launches')}); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step']); print(d['kernel_ms_profile_pass'])
PY
grep -v "^W\|^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_${N}gpu.err | tail -5
