"""One process per GPU. Clients are sharded so that an edge server and all of its clients live on the
same GPU: the inner loop needs no cross-GPU traffic; the only exchange is the Cloud / FL
aggregation, a pre-weighted partial sum per rank followed by one all-reduce of the packed vector.

torch.distributed is the rendezvous plumbing (it carries the 128-byte NCCL unique id); the data-path
collective itself is issued by the engine on the compute stream (cgl_mix_allreduce)."""
import ctypes as C
import os

import torch

from . import abi


def shard_range(n_items, world, rank):
    """Contiguous block of `n_items` owned by `rank` (servers are dealt in contiguous blocks, like the
    reference deals workers to servers, CGLGAN/2DMG/main.py:468-474)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_plan(num_workers, num_servers, world):
    """[(server_lo, server_hi, client_lo, client_hi)] per rank; servers never straddle ranks."""
    per = num_workers // num_servers
    plan = []
    for r in range(world):
        lo, hi = shard_range(num_servers, world, r)
        plan.append((lo, hi, lo * per, hi * per))
    return plan


def global_weights(local_sizes, group=None):
    """Each rank's slice of the normalised aggregation weights A = size / sum over ALL ranks
    (Cloud.run: A /= A.sum(), CGLGAN/2DMG/main.py:117-122). Works on any backend (gloo on CPU)."""
    import torch.distributed as dist
    t = torch.as_tensor(local_sizes, dtype=torch.float32)
    tot = t.sum().reshape(1).clone()
    if dist.is_available() and dist.is_initialized():
        dev = t.device
        if dist.get_backend(group) == "nccl":
            tot = tot.cuda()
        dist.all_reduce(tot, group=group)
        tot = tot.to(dev)
    return t / tot


class ShardComm:
    """The engine-side communicator (wraps an ncclComm_t created by the C ABI)."""

    def __init__(self, world=None, rank=None):
        import torch.distributed as dist
        assert dist.is_initialized(), "init torch.distributed first (it carries the NCCL unique id)"
        self.world = dist.get_world_size() if world is None else world
        self.rank = dist.get_rank() if rank is None else rank
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            abi.check(abi.lib.cgl_comm_unique_id(uid))
        t = torch.tensor(list(uid), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        buf = (C.c_uint8 * 128)(*t.cpu().tolist())
        self.handle = C.c_void_p()
        abi.check(abi.lib.cgl_comm_init(self.world, self.rank, buf, C.byref(self.handle)))

    def allreduce_(self, t):
        from .engine import _stream
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        abi.check(abi.lib.cgl_allreduce_sum(self.handle, abi.ptr(t), t.numel(), _stream()))
        return t

    def close(self):
        if self.handle:
            abi.lib.cgl_comm_destroy(self.handle)
            self.handle = None


def env_rank():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))
