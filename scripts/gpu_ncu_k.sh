# ncu --set full of a few launches of the kernels matching $1 (regex), skipping $2 matches, capturing $3
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_short2.log; exit 1; }
timeout ${4:-600} ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-40} -c ${3:-8} -f -o gpurun_out/${5:-prof_k} $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
