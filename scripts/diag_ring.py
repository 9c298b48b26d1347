import sys, numpy as np
sys.path.insert(0, '.')
import gen_b200 as g
from oracle import oracle as O, closed_forms as cf
LG=[0.0,1.0,0.9,0.0,1.0,1.0,1.0]
orc=O.Oracle()
ys=cf.simulate_lgssm(40, LG, 3)
for N in (6000, 50000):
  for mode in ('sorted','iid'):
    b=g.ParticleFilterState(g.LinearGaussianSSM(*LG), N, seed=2, keep_history=False)
    pf=orc.particle_filter(O.LGSSM, LG, N, seed=2)
    b.init([ys[0]]); pf.init([ys[0]])
    rng=np.random.default_rng(1)
    for t in range(1,7):
        u=rng.random(N) if mode=='iid' else None
        if u is not None: b.set_replay(None,u)
        db=b.maybe_resample(N*0.9); do=pf.maybe_resample(N*0.9, u_replay=u)
        msg=''
        if db:
            ag, ao = b.ancestors(), pf.parents()
            bad=np.nonzero(ag!=ao)[0]
            msg='anc bad %d first %s gpu %s orc %s'%(bad.size, bad[:4], ag[bad[:4]], ao[bad[:4]])
        b.step([ys[t]]); pf.step([ys[t]])
        lw_ok=np.array_equal(b.log_weights().view(np.uint64), pf.log_weights().view(np.uint64))
        print(N, mode,'t',t,'res',db,do,'ess %.6f %.6f'%(b.last_ess, pf.last_ess), msg,'lw',lw_ok)
