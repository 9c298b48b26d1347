# quick GPU check: parity tests, smoke, short bench (no cpu baseline)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline_whole_run']['frac'],'prop frac',d['roofline']['frac'])
print(d['kernel_ms_profile_pass'], d['log_ml'])
PY
tail -3 gpurun_out/bench_quick.err
GSMC_NO_PDL=1 python bench.py --no-cpu-baseline > gpurun_out/bench_nopdl.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_nopdl.json')); print('NO_PDL ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
