"""Server-side generators, stacked over edge servers and run by the CUDA engine.

The reference builds one `Generator` nn.Module and one optim.Adam per `Server` thread
(CGLGAN/2DMG/main.py:191-192, mixed-gan.py:180-182) and lets autograd walk every client's D back into it.
Here the S servers' generators live in packed buffers (same packed-row convention as the clients'
discriminators) and one call advances all of them:

  forward        cgl_mlp_forward   Linear [+ BatchNorm1d(eps 0.8, batch stats)] + LeakyReLU / Tanh, grouped
  backward_step  cgl_mlp_backward  data / weight gradients with the Adam step fused into the epilogues

Packed rows follow parameters() order of the reference module, split at the trunk/head boundary:
  trunk bank [S, .]   : model.*                    (model/mnist_model.py:17-24 / 45-49)
  head bank  [S*N, .] : paths.i.* of server s in row s*N+i (model/mnist_model.py:52-57)
BatchNorm running_mean/var are kept in a parallel stats row (train-mode batch statistics with
eps=0.8, momentum 0.1; both the no_grad Xd pass and the Xg pass update them, CGLGAN/2DMG/main.py:229-234).
There is no PyTorch or CPU implementation of the math here: without the CUDA library this module fails.
"""
import ctypes as C

import torch

from . import abi
from . import models
from .layout import RowLayout, flatten_bn_stats, flatten_params, load_bn_stats, load_flat_params


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Bank:
    """[rows, ld] packed parameters + Adam moments + step + BatchNorm running stats of one MLP stack."""

    def __init__(self, desc, rows, device):
        self.desc = desc
        self.lay = RowLayout(desc)
        self.rows = rows
        lay = self.lay
        self.params = torch.zeros(rows, lay.ld, device=device)
        self.adam_m = torch.zeros(rows, lay.ld, device=device)
        self.adam_v = torch.zeros(rows, lay.ld, device=device)
        self.step = torch.zeros(rows, dtype=torch.int32, device=device)
        self.stats = torch.zeros(rows, lay.ld_stats, device=device)
        self.ws = None

    def workspace(self, B):
        n = abi.lib.cgl_mlp_workspace_bytes(C.byref(self.desc), self.rows, B)
        if self.ws is None or self.ws.numel() < n:
            self.ws = torch.empty(n, dtype=torch.uint8, device=self.params.device)
        return self.ws

    def n_kernels(self, backward):
        """Upper bound of kernels per call (bench bookkeeping; the exact count is cgl_launch_count)."""
        L, nbn = self.lay.n_layers, sum(self.lay.bn)
        return (L + nbn) if not backward else (2 + nbn + 3 * L)


class StackedGenerator:
    """S generators with `n_heads` heads each (n_heads == 0: plain single-path generator).
    Trunks live in one bank [S, ld_trunk], heads in another [S*N, ld_head] (head i of server s is row
    s*N+i), so both run as grouped products without gathering."""

    def __init__(self, img_shape, n_servers, n_heads, device="cuda", lr=0.0002, b1=0.5, b2=0.999, eps=1e-8):
        abi.require_device()
        self.img_shape = tuple(img_shape)
        d = 1
        for s in self.img_shape:
            d *= s
        self.d = d
        self.S, self.N = int(n_servers), int(n_heads)
        self.device = torch.device(device)
        two_d = d == 2
        if self.N == 0:
            self.trunk = _Bank(abi.arch_describe(abi.ARCH_G_2D_MD if two_d else abi.ARCH_G_MNIST), self.S, self.device)
            self.heads = None
        else:
            self.trunk = _Bank(abi.arch_describe(abi.ARCH_G_2D_TRUNK if two_d else abi.ARCH_G_MNIST_TRUNK),
                               self.S, self.device)
            self.heads = _Bank(abi.arch_describe(abi.ARCH_G_2D_HEAD if two_d else abi.ARCH_G_MNIST_HEAD),
                               self.S * self.N, self.device)
            self.head_src = (torch.arange(self.S * self.N, device=self.device, dtype=torch.int32) // self.N).contiguous()
            self.head_ptr = torch.arange(0, self.S * self.N + 1, self.N, device=self.device, dtype=torch.int32)
        self.P_trunk = self.trunk.lay.n_params
        self.P_head = self.heads.lay.n_params if self.heads else 0
        self.P = self.P_trunk + self.N * self.P_head
        self.cfg = abi.TrainCfg(0, 1.0, lr, b1, b2, eps)
        self.training = True
        self._last = None     # (z, trunk_out, head_out) of the latest forward: what backward_step differentiates
        self.launches = 0

    def banks(self):
        return [self.trunk] + ([self.heads] if self.heads else [])

    # ---- reference-module I/O ---------------------------------------------------------------
    def make_module(self):
        if self.N == 0:
            return models.Generator(self.img_shape)
        return models.MixGenerator(self.img_shape, self.N)

    def load_modules(self, mods):
        """Server s <- reference-style module mods[s] (Generator / MixGenerator)."""
        assert len(mods) == self.S
        cache = {}

        def rows_of(m):
            if id(m) not in cache:
                t = flatten_params(m.model).float()
                ts = flatten_bn_stats(m.model).float()
                hs = [(flatten_params(p).float(), flatten_bn_stats(p).float()) for p in m.paths] if self.N else []
                cache[id(m)] = (t, ts, hs)
            return cache[id(m)]

        tp = torch.zeros(self.S, self.trunk.lay.ld)
        tst = torch.zeros(self.S, self.trunk.lay.ld_stats)
        if self.N:
            hp = torch.zeros(self.S * self.N, self.heads.lay.ld)
            hst = torch.zeros(self.S * self.N, self.heads.lay.ld_stats)
        for s, m in enumerate(mods):
            t, ts, hs = rows_of(m)
            tp[s, :self.P_trunk] = t
            if ts.numel():
                tst[s, :ts.numel()] = ts
            if self.N:
                assert len(m.paths) == self.N
                for i, (h, hst_i) in enumerate(hs):
                    hp[s * self.N + i, :self.P_head] = h
                    if hst_i.numel():
                        hst[s * self.N + i, :hst_i.numel()] = hst_i
        self.trunk.params.copy_(tp)
        self.trunk.stats.copy_(tst)
        if self.N:
            self.heads.params.copy_(hp)
            self.heads.stats.copy_(hst)

    def store_module(self, s, m):
        load_flat_params(m.model, self.trunk.params[s, :self.P_trunk].cpu())
        if self.trunk.lay.n_stats:
            load_bn_stats(m.model, self.trunk.stats[s, :self.trunk.lay.n_stats].cpu())
        if self.N:
            for i, path in enumerate(m.paths):
                r = s * self.N + i
                load_flat_params(path, self.heads.params[r, :self.P_head].cpu())
                if self.heads.lay.n_stats:
                    load_bn_stats(path, self.heads.stats[r, :self.heads.lay.n_stats].cpu())

    def flat_rows(self):
        """[S, P] serialize_model view of every server's generator (parameters() order)."""
        t = self.trunk.params[:, :self.P_trunk]
        if not self.N:
            return t.clone()
        h = self.heads.params[:, :self.P_head].reshape(self.S, self.N * self.P_head)
        return torch.cat([t, h], dim=1)

    # ---- forward ---------------------------------------------------------------------------------
    def _fwd(self, bank, x, x_idx, B, out_dim, ids=None):
        G = bank.rows if ids is None else ids.numel()      # ids: the bank rows that take part (FeGAN's group of the round)
        y = torch.empty(G, B, out_dim, device=self.device)
        ws = bank.workspace(B)
        lay = bank.lay
        abi.check(abi.lib.cgl_mlp_forward(C.byref(bank.desc), G, abi.ptr(bank.params), lay.ld, abi.ptr(ids),
                                          abi.ptr(bank.stats), lay.ld_stats, 1 if self.training else 0, abi.ptr(x),
                                          B * lay.dims[0], abi.ptr(x_idx), B, abi.ptr(y), abi.ptr(ws), ws.numel(),
                                          _stream()))
        self.launches += bank.n_kernels(False)
        return y

    def forward(self, z, ids=None):
        """z [S, B, 100] -> plain: [S, B, d]; multi-head: [S, N, B, d] (head i of server s feeds its
        i-th client, torch.chunk(net_g(z), N) in CGLGAN/2DMG/main.py:231). The activations of the LATEST
        forward are what backward_step differentiates (the reference's Xg pass comes after its no_grad
        Xd pass, CGLGAN/2DMG/main.py:229-234)."""
        S, N = self.S, self.N
        if ids is not None:
            assert N == 0, "row subsets are for single-path generators (FL-style clients)"
            ids = ids.to(device=self.device, dtype=torch.int32).contiguous()
            S = ids.numel()
        z = z.reshape(S, -1, self.trunk.lay.dims[0]).contiguous().float()
        B = z.shape[1]
        t_out = self._fwd(self.trunk, z, None, B, self.trunk.lay.dims[-1], ids)
        self._last_ids = ids
        if N == 0:
            self._last = (z, t_out, None)
            return t_out
        h_out = self._fwd(self.heads, t_out, self.head_src, B, self.d)
        self._last = (z, t_out, h_out)
        return h_out.view(S, N, B, self.d)

    __call__ = forward

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    # ---- backward + optimiser --------------------------------------------------------------------
    def _bwd(self, bank, x, x_idx, y, dy, B, want_dx, ids=None):
        lay = bank.lay
        G = bank.rows if ids is None else ids.numel()
        dx = torch.empty(G, B, lay.dims[0], device=self.device) if want_dx else None
        ws = bank.workspace(B)
        abi.check(abi.lib.cgl_mlp_backward(C.byref(bank.desc), G, abi.ptr(bank.params), abi.ptr(bank.adam_m),
                                           abi.ptr(bank.adam_v), lay.ld, abi.ptr(bank.step), abi.ptr(ids), C.byref(self.cfg),
                                           abi.ptr(x), B * lay.dims[0], abi.ptr(x_idx), B, abi.ptr(y), abi.ptr(dy),
                                           abi.ptr(dx), abi.ptr(ws), ws.numel(), _stream()))
        self.launches += bank.n_kernels(True)
        return dx

    def backward_step(self, dy, trunk_w=None):
        """Back-propagate dLoss/d(output of the latest forward) and take the generators' Adam step
        (F_max.backward(); opti.step() -- CGLGAN/2DMG/main.py:254-276, mixed-gan.py:263-288).
        dy: [S, B, d] (plain) or [S, N, B, d]. trunk_w [S, N] (multi-head only): the heads are trained on
        d(sum_i loss_i) while the trunk receives sum_i w_i * d loss_i (heads-only backward of `losses`, then
        the trunk-only backward of F_max with the paths frozen)."""
        assert self._last is not None, "forward() has not run"
        z, t_out, h_out = self._last
        S, N = self.S, self.N
        B = z.shape[1]
        dy = dy.contiguous().float()
        if N == 0:
            ids = getattr(self, "_last_ids", None)
            self._bwd(self.trunk, z, None, t_out, dy.reshape(z.shape[0], B, -1), B, False, ids)
        else:
            dh = self._bwd(self.heads, t_out, self.head_src, h_out, dy.reshape(S * N, B, self.d), B, True)
            hid = t_out.shape[2]
            dt = torch.empty(S, B, hid, device=self.device)
            w = None if trunk_w is None else trunk_w.reshape(S * N).contiguous().float()
            abi.check(abi.lib.cgl_dxg_reduce(S, abi.ptr(self.head_ptr), None, abi.ptr(w), abi.ptr(dh), B * hid,
                                             abi.ptr(dt), _stream()))
            self.launches += 1
            self._bwd(self.trunk, z, None, t_out, dt, B, False)
        self._last = None
