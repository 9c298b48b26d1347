"""Where a sharded run loses time against one GPU: the cfg-3 filter at 2^24 particles per GPU with ess_threshold = 0
(never resamples: propagate + the triple exchange of every step), N (always) and N/2, through gsmc_run_steps.
  torchrun --nproc-per-node R scripts/mgpu_breakdown.py      (or plain python for one GPU)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gen_b200 as g  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
comm = dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    from gen_b200.distributed import Communicator
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Communicator(dist, rank, world, device=local)
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
N, T = (1 << log2n) * world, 100
model = g.LinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0)
rng = np.random.default_rng(0)
x, ys = rng.standard_normal(), []
for t in range(T):
    if t:
        x = 0.9 * x + rng.standard_normal()
    ys.append(x + rng.standard_normal())
ys = np.array(ys)
out = {}
for name, thr in (("never", 0.0), ("always", float(N)), ("half", N / 2)):
    st = g.ParticleFilterState(model, N, seed=0, keep_history=False, device=local, comm=comm)
    best = 1e9
    for rep in range(6):
        st.reset()
        st.init([ys[0]])
        if dist is not None:
            dist.barrier()
        st.synchronize()
        st.timer_start()
        st.run_steps(ys[1:], thr)
        ms = st.timer_stop()
        if dist is not None:
            t_ = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms = float(t_.item())
        if rep >= 2:
            best = min(best, ms)
    out[name] = (best / (T - 1) * 1e3, st.stats()["num_resamples"])
    st.close()
if rank == 0:
    print("R=%d N=2^%d per GPU: never %.1f us/step, always %.1f us/step, ESS<N/2 %.1f us/step (%d resamples); one event = %.1f us" % (
        world, log2n, out["never"][0], out["always"][0], out["half"][0], out["half"][1], out["always"][0] - out["never"][0]), flush=True)
if dist is not None:
    comm.close()
    dist.destroy_process_group()
