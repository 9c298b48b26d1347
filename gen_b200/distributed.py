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
