# one --set full capture covering every kernel of a filter step + resample (a window of launches of the 4th run)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_short2.log; exit 1; }
timeout ${3:-900} ncu --set full --clock-control none --import-source on -s ${1:-700} -c ${2:-100} -f -o gpurun_out/prof_all $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
