mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hmm or run_steps or multinomial" 2>&1 | tail -3
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-12} -c ${3:-6} -o gpurun_out/prof_r1c $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
