mkdir -p gpurun_out
for M in 0x3e 0x3f 0x00 0x3e 0x3f; do
  GSMC_PDL_MASK=$M python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_mask.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_mask.json')); print('mask $M ms_per_step %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
