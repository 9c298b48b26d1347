import os, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch, torch.distributed as dist
import gen_b200 as g
from gen_b200.distributed import Communicator
from oracle import closed_forms as cf
rank, world, local = int(os.environ.get("RANK",0)), int(os.environ.get("WORLD_SIZE",1)), int(os.environ.get("LOCAL_RANK",0))
torch.cuda.set_device(local)
comm=None
if world>1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm=Communicator(dist, rank, world, device=local)
LG=[0.0,1.0,0.9,0.0,1.0,1.0,1.0]; T=100; N=(1<<24)*world
ys=cf.simulate_lgssm(T, LG, 0)
st=g.ParticleFilterState(g.LinearGaussianSSM(*LG), N, seed=0, keep_history=True, history_capacity=T, device=local, comm=comm)
def run():
    st.reset(); st.init([ys[0]]); st.run_steps(ys[1:], N/2); return st.log_ml_estimate()
for _ in range(3): run()
for prof in (False, True):
    st.set_profiling(prof)
    if world>1: dist.barrier()
    st.synchronize(); t0=time.perf_counter(); st.timer_start()
    for _ in range(3): run()
    ms=st.timer_stop(); wall=(time.perf_counter()-t0)*1e3
    s=st.stats()
    print(f"rank {rank} prof={prof} dev_ms/run {ms/3:.2f} wall/run {wall/3:.2f}", {k[3:]:round(v/3,2) for k,v in s.items() if k.startswith('ms_')} if prof else '', flush=True)
st.close()
if world>1: dist.destroy_process_group()
