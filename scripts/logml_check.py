"""Bias check of the log-ML estimate after the change of the normal generator: the cfg-3 filter (LG-SSM, T = 100,
multinomial at ESS < N/2) with several seeds and particle counts against the Kalman filter's exact log p(y).
The estimator is unbiased for p(y) (not for its log): E[log-ML] = log p(y) - var/2; its standard deviation falls like 1/sqrt(N).
  python scripts/logml_check.py"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gen_b200 as g  # noqa: E402
from bench import kalman  # noqa: E402

params = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]
T = 100
rng = np.random.default_rng(0)
x, ys = params[0] + params[1] * rng.standard_normal(), []
for t in range(T):
    if t:
        x = params[2] * x + params[3] + params[4] * rng.standard_normal()
    ys.append(params[5] * x + params[6] * rng.standard_normal())
ys = np.array(ys)
exact = kalman(ys)          # bench.py's LG parameters are the ones above
print("Kalman log p(y) = %.6f" % exact)
for log2n, seeds in ((18, 24), (20, 24), (22, 12), (24, 4)):
    N = 1 << log2n
    errs = []
    for s in range(seeds):
        st = g.ParticleFilterState(g.LinearGaussianSSM(*params), N, seed=1000 + 17 * s, keep_history=False, device=0)
        st.init([ys[0]])
        st.run_steps(ys[1:], N / 2)
        errs.append(st.log_ml_estimate() - exact)
        st.close()
    e = np.array(errs)
    print("N=2^%d: %d seeds, mean error %+.5f, std %.5f, standard error of the mean %.5f, z = %+.2f" % (
        log2n, seeds, e.mean(), e.std(ddof=1), e.std(ddof=1) / math.sqrt(seeds), e.mean() / (e.std(ddof=1) / math.sqrt(seeds))), flush=True)
