# end-of-milestone measurement set: tests, smoke, bench (ours + reference arm), ncu launch list, ncu --set full of every hot kernel
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err; echo "ref rc=$?"
timeout 600 python scripts/run_configs.py > gpurun_out/configs_r1.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_short2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"propagate_kernel|finalize|weights_kernel|partition|search_sorted" -s 50 -c 10 -f -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
