# usage: gpu_ncu_kernel.sh TAG KERNEL_REGEX [SKIP] [COUNT] -- ncu --set full of one kernel family of the short bench run
TAG=$1; RX=$2; SKIP=${3:-8}; CNT=${4:-6}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
