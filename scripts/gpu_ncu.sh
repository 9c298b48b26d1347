mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_short2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-12} -c ${3:-6} -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
