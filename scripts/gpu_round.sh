# usage: gpu_round.sh TAG -- end-of-milestone measurement set: tests, smoke, bench (ours + reference arm), configs, ncu launch list, ncu --set full of every hot kernel
TAG=${1:-r2}
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 600 python scripts/run_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_short2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"propagate_kernel|finalize|weights_kernel|partition|search_sorted" -s 60 -c 14 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
