# usage: gpu_multi.sh NGPUS [TAG] -- sharded parity worker + bench on N GPUs
mkdir -p gpurun_out
N=$1; TAG=${2:-r2}
git_sha=$(cat .git_sha 2>/dev/null)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_worker.py > gpurun_out/${TAG}_mgpu$N.log 2>&1; echo "worker rc=$?"; grep "AssertionError\| ok on\|Error" gpurun_out/${TAG}_mgpu$N.log | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
for line in open('gpurun_out/${TAG}_bench_${N}gpu.json'):
    if line.startswith('{'):
        d=json.loads(line); print({k:d[k] for k in ('n_gpus','value','ms_per_step','log_ml','gpu_launches')}); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step']); print(d['kernel_ms_profile_pass'])
PY
grep -v "^W\|^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/${TAG}_bench_${N}gpu.err | tail -5
