mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cur.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
