# usage: gpu_8b.sh NGPUS TAG -- sharded parity worker, cfg-3 bench and the step/event breakdown on N GPUs (no cfg-5 run)
mkdir -p gpurun_out
N=$1; TAG=${2:-r2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29512 tests/mgpu_worker.py > gpurun_out/${TAG}_mgpu$N.log 2>&1; echo "worker rc=$?"; grep "AssertionError\| ok on\|Error" gpurun_out/${TAG}_mgpu$N.log | head -12
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench cfg3 rc=$?"
python - <<PY
import json
for line in open('gpurun_out/${TAG}_bench_${N}gpu.json'):
    if line.startswith('{'):
        d=json.loads(line); print({k:d[k] for k in ('n_gpus','value','ms_per_step','log_ml','log_ml_e2e','gpu_launches')}); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline_whole_run']['frac']); print(d['kernel_ms_profile_pass']); print(d.get('nvlink'))
PY
timeout 300 $RUN --master-port 29515 scripts/mgpu_breakdown.py 2>&1 | grep "^R=" | tee gpurun_out/${TAG}_breakdown_${N}gpu.txt
