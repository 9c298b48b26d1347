mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_quick.json'))
    print({k:d[k] for k in ('value','ms_per_step','log_ml','gpu_launches')})
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['launches'])
    print('whole', d['roofline_whole_run'])
    print('kernel ms', d['kernel_ms_profile_pass'])
    print('clocks', d['clocks'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
