"""One process per GPU: torch.distributed is only the plumbing that carries the NCCL unique id to
every rank; the collectives of the filter (allgather of logsumexp partials, of per-rank integer
weight totals and of spacing totals) and the peer-memory ancestor gather run inside libgensmc.so."""
import ctypes as C

import numpy as np

from . import _lib
from . import shard


class Communicator:
    """Process-wide NCCL communicator of libgensmc.so (one per rank)."""

    def __init__(self, dist, rank, world_size, device=-1):
        self.dist, self.rank, self.world_size = dist, int(rank), int(world_size)
        self.lib = _lib.load()
        self.handle = None
        raw = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(self.lib.gsmc_comm_unique_id(raw, 128))
        uid = self.broadcast_bytes(raw.raw, 128)
        buf = C.create_string_buffer(uid, 128)
        h = C.c_void_p()
        _lib.check(self.lib.gsmc_comm_create(buf, 128, self.rank, self.world_size, int(device), C.byref(h)))
        self.handle = h

    def broadcast_bytes(self, payload, nbytes):
        """Rank 0's `payload` (bytes) to every rank through torch.distributed (gloo or nccl)."""
        import torch
        backend = self.dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        if self.rank == 0:
            buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        self.dist.broadcast(buf, src=0)
        return bytes(buf.cpu().numpy().tobytes())

    def attach(self, state):
        shard.check_partition(state.num_particles, self.world_size)
        _lib.check(self.lib.gsmc_comm_attach(state.handle, self.handle), state.handle)

    def close(self):
        if self.handle:
            self.lib.gsmc_comm_destroy(self.handle)
            self.handle = None


class _GroupRank:
    """What ParticleFilterState(comm=...) needs from a communicator, for one logical rank of a LocalShardGroup."""

    def __init__(self, group, rank):
        self.group, self.rank = group, rank

    def attach(self, state):
        self.group._attach(self.rank, state)


class LocalShardGroup:
    """Shard emulation (gsmc_group_* in include/gen_b200.h): the R ranks of ONE sharded filter as R handles on one
    device and one stream. The data path is the multi-GPU one; the scalar exchanges are direct reads, ordered by the
    group calls (every rank's producers before any rank's consumers). For the 1-GPU parity tests of the sharded path.

        grp = LocalShardGroup(4)
        shards = [ParticleFilterState(model, N, comm=grp.rank(r), ...) for r in range(4)]
        grp.init([y0]); grp.maybe_resample(N / 2); grp.step([y1]); shards[2].log_weights() ..."""

    def __init__(self, world_size, device=-1):
        self.lib = _lib.load()
        self.world_size = int(world_size)
        h = C.c_void_p()
        _lib.check(self.lib.gsmc_group_create(self.world_size, int(device), C.byref(h)))
        self.handle = h
        self.states = [None] * self.world_size

    def rank(self, r):
        return _GroupRank(self, int(r))

    def _attach(self, r, state):
        shard.check_partition(state.num_particles, self.world_size)
        _lib.check(self.lib.gsmc_group_attach(self.handle, r, state.handle), state.handle)
        self.states[r] = state
        state.close = lambda: None          # the group owns its members (gsmc_group_destroy destroys them)

    def _propagate(self, fn, obs, proposal):
        o = np.ascontiguousarray(obs, dtype=np.float64)
        if proposal is None:
            rc = fn(self.handle, _lib.dptr(o), o.size, _lib.PROPOSAL_DEFAULT, None, 0)
        else:
            pp = np.ascontiguousarray(proposal.params, dtype=np.float64)
            rc = fn(self.handle, _lib.dptr(o), o.size, proposal.proposal_id, _lib.dptr(pp), pp.size)
        _lib.check(rc)
        for st in self.states:
            st.T += 1
            st.observations.append(o.copy())

    def init(self, obs, proposal=None):
        self._propagate(self.lib.gsmc_group_init, obs, proposal)

    def step(self, obs, proposal=None):
        self._propagate(self.lib.gsmc_group_step, obs, proposal)

    def maybe_resample(self, ess_threshold):
        did, ess = C.c_int(), C.c_double()
        _lib.check(self.lib.gsmc_group_maybe_resample(self.handle, float(ess_threshold), C.byref(did), C.byref(ess)))
        for st in self.states:
            st.last_ess = ess.value
        return bool(did.value)

    def sample_unweighted(self, num_samples):
        out = np.empty(int(num_samples), dtype=np.int64)
        _lib.check(self.lib.gsmc_group_sample_unweighted(self.handle, int(num_samples), _lib.iptr(out)))
        return out

    def close(self):
        if self.handle:
            self.lib.gsmc_group_destroy(self.handle)
            self.handle = None
            for st in self.states:
                if st is not None:
                    st.handle = None

    __del__ = close
