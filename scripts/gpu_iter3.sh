# usage: gpu_iter3.sh TAG -- parity tests, smoke, bench and the 1-GPU step/event cost breakdown
TAG=${1:-it}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline_whole_run']['frac'],'prop frac',d['roofline']['frac'],'resample frac',d['roofline_resample']['frac'])
print(d['kernel_ms_profile_pass'], d['log_ml'])
PY
timeout 300 python scripts/mgpu_breakdown.py 2>&1 | grep "^R=" | tee gpurun_out/${TAG}_breakdown_1gpu.txt
