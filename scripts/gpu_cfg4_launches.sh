# ncu launch list of cfg 4 (stochastic volatility, residual resampling, N=2^22): which kernels the residual path spends its time in
mkdir -p gpurun_out
TAG=${1:-r2}
GSMC_NO_GRAPH=1 python scripts/run_configs.py cfg4 > gpurun_out/cfg4_plain.log 2>&1 && \
GSMC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 6000 --csv --log-file gpurun_out/${TAG}_cfg4_launches.csv python scripts/run_configs.py cfg4 > gpurun_out/ncu_cfg4.log 2>&1
echo "rc=$?"
