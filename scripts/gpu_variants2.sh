# kernel-parameter sweep over variants/*.so: cfg 3 bench + cfg 4 run
mkdir -p gpurun_out
cp gen_b200/libgensmc.so /tmp/head.so
for v in variants/*.so; do
  cp $v gen_b200/libgensmc.so
  python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/var.json 2>/dev/null
  python scripts/run_configs.py cfg4 > gpurun_out/var4.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/var.json')); k=d['kernel_ms_profile_pass']
print("$v", round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'search', round(k['search'],2), 'scan', round(k['scan'],2), 'prop', round(k['propagate'],2), round(k['propagate_gather'],2), 'lml', d['log_ml'])
d=json.loads(open('gpurun_out/var4.json').readline()); print("   cfg4 ms_per_run", round(d['ms_per_run'],2), d['log_ml'])
PY
done
cp /tmp/head.so gen_b200/libgensmc.so
