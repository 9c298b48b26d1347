# usage: gpu_iter2.sh TAG -- configs 4 and 5 (one shard) + ncu --set full of propagate at HEAD
TAG=${1:-it}
mkdir -p gpurun_out
timeout 600 python scripts/run_configs.py cfg4 cfg5 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"; cat gpurun_out/${TAG}_configs.jsonl | cut -c1-600
timeout 300 python scripts/resample_cost.py 2>&1 | tail -6
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"propagate_kernel" -s 8 -c 4 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
