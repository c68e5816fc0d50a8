"""The data side path of a round, kept on the GPU.

ResidentPartitions: the whole dataset lives in HBM once; every simulated client owns a list of row ids (its
partition from partition.allocate_dataset). A round's real minibatches are produced exactly as the reference
produces them -- `DataLoader(dataset, batch_size, shuffle=True)` per Worker, `next(self.data)`, a NEW DataLoader
when the epoch is exhausted (CGLGAN/2DMG/main.py:299-301,350-355) -- except that the DataLoaders iterate over
ROW IDS, not samples: same classes, same consumption of the torch global RNG, hence bit-identical batches, and
only 8 bytes per sample travel host -> device; `cgl_gather_rows` assembles [C, B, d] on the GPU.
shuffle=False gives the sequential full pass of FL-GAN (FLGAN/MNIST/flgan.py:250).

kl_score_2d: plot_2d's KL score (CGLGAN/2DMG/main.py:68-94) on the GPU.
"""
import ctypes as C

import torch

from . import abi
from .engine import _stream


STAGING_SLOTS = 4


class ResidentPartitions:
    def __init__(self, data, partitions, batch_size, device="cuda", shuffle=True):
        """data: [n, ...] float tensor (moved to the device once); partitions: one sequence of row ids per client."""
        abi.require_device()
        self.device = torch.device(device)
        self.n = data.shape[0]
        self.data = data.reshape(self.n, -1).to(self.device, torch.float32).contiguous()
        self.d = self.data.shape[1]
        self.parts = [torch.as_tensor(p, dtype=torch.int64) for p in partitions]
        self.C, self.B, self.shuffle = len(self.parts), int(batch_size), shuffle
        # Worker.__init__: one DataLoader per client and its iterator, created in client order. The host side keeps no
        # DataLoader objects (1024 Python iterators cost ~80 ms per round, three times the round itself): it makes the
        # same draws from torch's global RNG, in the same order, that DataLoader(shuffle=...) makes --
        #   iter(DataLoader)         : _BaseDataLoaderIter draws its base seed        torch.empty((), int64).random_()
        #   first next() of an epoch : RandomSampler draws its seed the same way and takes torch.randperm(n, generator)
        # -- so the row ids are bit-identical to the reference's (tests/test_gpu_data.py checks them against real DataLoaders).
        self._perm = [None] * self.C          # row ids of the running epoch in visiting order (None: not drawn yet)
        self._cur = [0] * self.C
        for _ in range(self.C):
            self._draw_seed()                 # iter(self.dataloader) in Worker.__init__
        # A ring of pinned staging slots: the host runs ahead of the stream (a round enqueues dozens of kernels), so a
        # slot is rewritten only after the event recorded behind its previous host -> device copy has completed.
        self._slots = [dict(idx=torch.empty(self.C, self.B, dtype=torch.int64).pin_memory(),
                            n=torch.empty(self.C, dtype=torch.int32).pin_memory(), ev=None) for _ in range(STAGING_SLOTS)]
        self._slot = 0
        self._ready = None

    def prefetch(self):
        """Draws the NEXT round's row ids now (host work only, the same draws in the same order): call it right after a
        round was enqueued so that it runs while the GPU works; next_batches() then only ships them."""
        if self._ready is None:
            slot = self._slots[self._slot]
            self._slot = (self._slot + 1) % len(self._slots)
            self.next_indices(slot)
            self._ready = slot

    @staticmethod
    def _draw_seed():
        return int(torch.empty((), dtype=torch.int64).random_().item())

    def _start_epoch(self, c):
        part = self.parts[c]
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self._draw_seed())                       # RandomSampler.__iter__
            self._perm[c] = part[torch.randperm(part.numel(), generator=g)]
        else:
            self._perm[c] = part                                   # SequentialSampler
        self._cur[c] = 0

    def next_indices(self, slot=None):
        """One `next(self.data)` per client, in client order (Worker.train, main.py:350-355). Returns the pinned
        [C, B] row ids (-1 = padding of a short last batch) and the valid counts [C] of a staging slot that no
        copy in flight still reads."""
        if slot is None:
            slot = self._slots[self._slot]
            self._slot = (self._slot + 1) % len(self._slots)
        if slot["ev"] is not None:
            slot["ev"].synchronize()
            slot["ev"] = None
        idx, n = slot["idx"], slot["n"]
        idx.fill_(-1)
        B = self.B
        for c in range(self.C):
            perm = self._perm[c]
            if perm is None:
                self._start_epoch(c)                               # the first next() of the iterator made in __init__
            elif self._cur[c] >= perm.numel():                     # StopIteration: a NEW DataLoader and iterator, then next()
                self._draw_seed()
                self._start_epoch(c)
            perm, cur = self._perm[c], self._cur[c]
            rows = perm[cur:cur + B]
            k = rows.numel()
            idx[c, :k] = rows
            n[c] = k
            self._cur[c] = cur + k
        return idx, n

    def next_batches(self, out=None):
        """-> (real [C, B, d] on the device, n_real [C] int32 on the device) for MDStyleSim.round / FLStyleSim."""
        if self._ready is not None:
            slot, self._ready = self._ready, None
            idx, n = slot["idx"], slot["n"]
        else:
            slot = self._slots[self._slot]
            self._slot = (self._slot + 1) % len(self._slots)
            idx, n = self.next_indices(slot)
        idx_dev = idx.to(self.device, non_blocking=True)      # per-round device tensors: nothing aliases across rounds
        n_dev = n.to(self.device, non_blocking=True)
        slot["ev"] = torch.cuda.Event()
        slot["ev"].record()
        if out is None:
            out = torch.empty(self.C, self.B, self.d, device=self.device)
        abi.check(abi.lib.cgl_gather_rows(self.C * self.B, self.d, abi.ptr(self.data), self.n, abi.ptr(idx_dev),
                                          abi.ptr(out), _stream()))
        return out, n_dev

    def gather(self, idx):
        """[n, B] row ids on the host (-1 = padding) -> [n, B, d] on the device (cgl_gather_rows)."""
        idx_dev = idx.contiguous().to(self.device, non_blocking=False)
        out = torch.empty(idx.shape[0], idx.shape[1], self.d, device=self.device)
        abi.check(abi.lib.cgl_gather_rows(idx.numel(), self.d, abi.ptr(self.data), self.n, abi.ptr(idx_dev), abi.ptr(out),
                                          _stream()))
        return out

    @property
    def h2d_bytes_per_round(self):
        return self.C * self.B * 8 + self.C * 4


class KLScore2D:
    """KL score of generated 2-D points against a fixed real sample (plot_2d, CGLGAN/2DMG/main.py:68-94)."""

    def __init__(self, real_points, device="cuda"):
        abi.require_device()
        self.device = torch.device(device)
        real = real_points.to(self.device, torch.float32).contiguous()
        self.real_hist = torch.zeros(256, dtype=torch.int32, device=self.device)
        self._scratch = torch.zeros(256, dtype=torch.int32, device=self.device)
        self._out = torch.zeros(1, dtype=torch.float64, device=self.device)
        abi.check(abi.lib.cgl_hist2d(real.shape[0], abi.ptr(real), real.shape[1], abi.ptr(self.real_hist), _stream()))

    def __call__(self, points):
        """points: [n, 2] device tensor -> 0-dim float64 device tensor."""
        pts = points.to(self.device, torch.float32).contiguous()
        abi.check(abi.lib.cgl_kl_score_2d(pts.shape[0], abi.ptr(pts), pts.shape[1], abi.ptr(self.real_hist),
                                          abi.ptr(self._scratch), abi.ptr(self._out), _stream()))
        return self._out[0].clone()
