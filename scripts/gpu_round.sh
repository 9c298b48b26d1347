mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1_a.json 2>&1; cat gpurun_out/bench_ref_r1_a.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --T 20 > gpurun_out/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --T 20 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 12 > gpurun_out/plain_short2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 20 -c 4 -o gpurun_out/prof_propagate_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --T 12 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
