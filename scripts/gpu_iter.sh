# usage: gpu_iter.sh TAG [KERNEL_REGEX] -- one iteration on the GPU: parity tests, smoke, bench, ncu --set full of one kernel family
TAG=${1:-it}; RX=${2:-propagate_kernel}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline_whole_run']['frac'],'prop frac',d['roofline']['frac'],'resample frac',d['roofline_resample']['frac'])
print(d['kernel_ms_profile_pass'], d['log_ml'])
PY
tail -3 gpurun_out/${TAG}_bench.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 20"
$CMD > gpurun_out/plain_short.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 8 -c 6 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
