"""Cost of one resampling event and of one plain step at a given size: the same filter with ess_threshold = 0 (never
resamples), N (always) through gsmc_run_steps; prints ms per step for both and the difference.
usage: python scripts/resample_cost.py [log2n=22] [scheme=residual|multinomial] [model=sv|lgssm]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gen_b200 as g  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
scheme = sys.argv[2] if len(sys.argv) > 2 else "residual"
which = sys.argv[3] if len(sys.argv) > 3 else "sv"
N, T = 1 << log2n, 200
model = g.StochasticVolatility(-1.0, 0.97, 0.2) if which == "sv" else g.LinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0)
ys = np.random.default_rng(0).standard_normal(T) * 0.6
out = {}
for name, thr in (("never", 0.0), ("always", float(N)), ("half", N / 2)):
    st = g.ParticleFilterState(model, N, seed=0, resample=scheme, keep_history=False)
    for rep in range(4):
        st.reset()
        st.init([ys[0]])
        if rep == 3:
            st.synchronize()
            st.timer_start()
        st.run_steps(ys[1:], thr)
        if rep == 3:
            out[name] = (st.timer_stop() / (T - 1), st.stats()["num_resamples"], st.stats()["graph_replays"])
    st.close()
print("N=2^%d %s %s: never %.1f us/step, always %.1f us/step (%d resamples, graph replays %d) -> one event = %.1f us; ESS<N/2: %.1f us/step (%d resamples)" % (
    log2n, scheme, which, out["never"][0] * 1e3, out["always"][0] * 1e3, out["always"][1], out["always"][2],
    (out["always"][0] - out["never"][0]) * 1e3, out["half"][0] * 1e3, out["half"][1]))
