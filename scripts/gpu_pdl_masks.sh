# PDL experiment: which kernel classes gain from starting early (GSMC_PDL_MASK bits: 0 propagate, 1 finalize, 2 weights, 3 partition, 4 search, 5 other)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for M in ${MASKS:-0x3f 0x00 0x3e 0x3f 0x00 0x3e 0x2e 0x3a 0x36}; do
  GSMC_PDL_MASK=$M python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_mask.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_mask.json')); print('mask $M ms_per_step %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k: round(v,2) for k,v in d['kernel_ms_profile_pass'].items()})"
done
