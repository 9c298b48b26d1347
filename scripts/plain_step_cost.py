"""Cost of a step that does not resample (cfg-3 model, 2^24 particles): gsmc_run_steps with threshold 0 (propagate + the
three early-exit launches of the resampling path) against a loop of plain step() calls (propagate only).
  python scripts/plain_step_cost.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gen_b200 as g  # noqa: E402

N, T = 1 << 24, 100
model = g.LinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0)
rng = np.random.default_rng(0)
x, ys = rng.standard_normal(), []
for t in range(T):
    if t:
        x = 0.9 * x + rng.standard_normal()
    ys.append(x + rng.standard_normal())
ys = np.array(ys)
for hist in (False, True):
    st = g.ParticleFilterState(model, N, seed=0, keep_history=hist, history_capacity=T, device=0)
    res = {}
    for mode in ("run_steps(threshold 0)", "step() loop"):
        best = 1e9
        for rep in range(5):
            st.reset()
            st.init([ys[0]])
            st.synchronize()
            st.timer_start()
            if mode.startswith("run"):
                st.run_steps(ys[1:], 0.0)
            else:
                for t in range(1, T):
                    st.step([ys[t]])
            ms = st.timer_stop()
            if rep >= 2:
                best = min(best, ms)
        res[mode] = best / (T - 1) * 1e3
    print("keep_history=%s: " % hist + ", ".join("%s %.1f us/step" % kv for kv in res.items()), flush=True)
    st.close()
