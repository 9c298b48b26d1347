import sys, time
sys.path.insert(0, '.')
import numpy as np
import gen_b200 as g
from oracle import closed_forms as cf
LG=[0.0,1.0,0.9,0.0,1.0,1.0,1.0]; T=100; N=1<<24
ys=cf.simulate_lgssm(T, LG, 0)
model=g.LinearGaussianSSM(*LG)
def tm(): return time.perf_counter()
for rep in range(4):
    t0=tm(); st=g.ParticleFilterState(model, N, seed=0, keep_history=True, history_capacity=T); t1=tm()
    st.init([ys[0]]); st.synchronize(); t2=tm()
    tr=ts=0.0
    for t in range(1,T):
        a=tm(); st.maybe_resample(N/2); b=tm(); st.step([ys[t]]); c=tm(); tr+=b-a; ts+=c-b
    lml=st.log_ml_estimate(); t3=tm()
    st.close(); t4=tm()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} init {1e3*(t2-t1):.1f} loop {1e3*(t3-t2):.1f} (resample calls {1e3*tr:.1f}, step calls {1e3*ts:.1f}) close {1e3*(t4-t3):.1f} total {1e3*(t4-t0):.1f}")
# API-level
for rep in range(3):
    t0=tm()
    state = g.initialize_particle_filter(model, (1,), g.choicemap(("y_init", float(ys[0]))), N, seed=0, keep_history=True, history_capacity=T)
    for Tn in range(2, T + 1):
        g.maybe_resample_b(state)
        g.particle_filter_step_b(state, (Tn,), (g.UnknownChange(),), g.choicemap((("chain", Tn - 1, "y"), float(ys[Tn - 1]))))
    out = g.log_ml_estimate(state); state.close()
    print(f"api rep {rep}: {1e3*(tm()-t0):.1f} ms")
import cProfile, pstats
def api_run():
    state = g.initialize_particle_filter(model, (1,), g.choicemap(("y_init", float(ys[0]))), N, seed=0, keep_history=True, history_capacity=T)
    for Tn in range(2, T + 1):
        g.maybe_resample_b(state)
        g.particle_filter_step_b(state, (Tn,), (g.UnknownChange(),), g.choicemap((("chain", Tn - 1, "y"), float(ys[Tn - 1]))))
    out = g.log_ml_estimate(state); state.close()
pr=cProfile.Profile(); pr.enable(); api_run(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
