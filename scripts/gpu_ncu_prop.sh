# ncu --set full of the hot kernels (after the plain run exits 0), few launches each
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 12"
$CMD > gpurun_out/plain_short2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-propagate_kernel}" -s ${SKIP:-20} -c ${COUNT:-3} -f -o gpurun_out/prof_cur $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
