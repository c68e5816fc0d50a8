"""Server-side generators, stacked over edge servers.

The reference builds one `Generator` nn.Module and one optim.Adam per `Server` thread
(CGLGAN/2DMG/main.py:191-192, mixed-gan.py:180-182). Here the S servers' generators live in ONE packed
buffer [S, ld] (same packed-row convention as the clients' discriminators) and run as batched
matrix products over the server axis, so a round costs O(1) launches instead of O(S) modules.
The batched contractions use torch.bmm (library plumbing, row a4 of SURVEY.md section 8 "may stay in PyTorch
initially"); the optimizer step is the engine's fused Adam (cgl_adam_rows).

Packed rows follow parameters() order of the reference module, split at the trunk/head boundary:
  trunk bank [S, .]   : model.*                    (model/mnist_model.py:17-24 / 45-49)
  head bank  [S*N, .] : paths.i.* of server s in row s*N+i (model/mnist_model.py:52-57)
BatchNorm running_mean/var are kept in a parallel stats row (train-mode batch statistics with
eps=0.8, momentum 0.1; both the no_grad Xd pass and the Xg pass update them, CGLGAN/2DMG/main.py:229-234).
"""
import torch

from . import abi
from .engine import adam_rows
from .layout import RowLayout, flatten_bn_stats, flatten_params, load_bn_stats, load_flat_params, padded
from . import models


class _ScaleGrad(torch.autograd.Function):
    """Identity forward; backward multiplies the incoming gradient by a weight tensor that is filled
    in AFTER the forward (the server only knows the loss weights once the clients have answered)."""

    @staticmethod
    def forward(ctx, x, holder):
        ctx.holder = holder
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        w = ctx.holder["w"]
        return (g if w is None else g * w), None


class _Bank:
    """[rows, ld] packed parameters + Adam moments + step + BatchNorm running stats of one MLP stack."""

    def __init__(self, lay, rows, device):
        self.lay, self.rows = lay, rows
        self.params = torch.zeros(rows, lay.ld, device=device, requires_grad=True)
        self.adam_m = torch.zeros(rows, lay.ld, device=device)
        self.adam_v = torch.zeros(rows, lay.ld, device=device)
        self.step = torch.zeros(rows, dtype=torch.int32, device=device)
        self.stats = torch.zeros(rows, lay.ld_stats, device=device)


class StackedGenerator:
    """S generators with `n_heads` heads each (n_heads == 0: plain single-path generator).
    Trunks live in one bank [S, ld_trunk], heads in another [S*N, ld_head] (head i of server s is row
    s*N+i), so both run as strided-batched products without gathering."""

    def __init__(self, img_shape, n_servers, n_heads, device="cuda", lr=0.0002, b1=0.5, b2=0.999):
        self.img_shape = tuple(img_shape)
        d = 1
        for s in self.img_shape:
            d *= s
        self.d = d
        self.S, self.N = int(n_servers), int(n_heads)
        self.device = torch.device(device)
        two_d = d == 2
        if self.N == 0:
            self.trunk = _Bank(RowLayout(abi.arch_describe(abi.ARCH_G_2D_MD if two_d else abi.ARCH_G_MNIST)),
                               self.S, self.device)
            self.heads = None
        else:
            self.trunk = _Bank(RowLayout(abi.arch_describe(
                abi.ARCH_G_2D_TRUNK if two_d else abi.ARCH_G_MNIST_TRUNK)), self.S, self.device)
            self.heads = _Bank(RowLayout(abi.arch_describe(
                abi.ARCH_G_2D_HEAD if two_d else abi.ARCH_G_MNIST_HEAD)), self.S * self.N, self.device)
        self.P_trunk = self.trunk.lay.n_params
        self.P_head = self.heads.lay.n_params if self.heads else 0
        self.P = self.P_trunk + self.N * self.P_head
        self.lr, self.b1, self.b2 = lr, b1, b2
        self.training = True
        self.trunk_scale = {"w": None}  # per-(server, head) weight applied to the trunk's gradient
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
        with torch.no_grad():
            for s, m in enumerate(mods):
                self.trunk.params[s, :self.P_trunk].copy_(flatten_params(m.model).float())
                st = flatten_bn_stats(m.model).float()
                if st.numel():
                    self.trunk.stats[s, :st.numel()].copy_(st)
                if self.N:
                    assert len(m.paths) == self.N
                    for i, path in enumerate(m.paths):
                        self.heads.params[s * self.N + i, :self.P_head].copy_(flatten_params(path).float())
                        st = flatten_bn_stats(path).float()
                        if st.numel():
                            self.heads.stats[s * self.N + i, :st.numel()].copy_(st)

    def store_module(self, s, m):
        load_flat_params(m.model, self.trunk.params[s, :self.P_trunk].detach().cpu())
        if self.trunk.lay.n_stats:
            load_bn_stats(m.model, self.trunk.stats[s, :self.trunk.lay.n_stats].cpu())
        if self.N:
            for i, path in enumerate(m.paths):
                r = s * self.N + i
                load_flat_params(path, self.heads.params[r, :self.P_head].detach().cpu())
                if self.heads.lay.n_stats:
                    load_bn_stats(path, self.heads.stats[r, :self.heads.lay.n_stats].cpu())

    def flat_rows(self):
        """[S, P] serialize_model view of every server's generator (parameters() order)."""
        t = self.trunk.params.detach()[:, :self.P_trunk]
        if not self.N:
            return t.clone()
        h = self.heads.params.detach()[:, :self.P_head].reshape(self.S, self.N * self.P_head)
        return torch.cat([t, h], dim=1)

    # ---- forward ---------------------------------------------------------------------------------
    def _run_stack(self, bank, x):
        """x [rows, B, in] -> [rows, B, out] through the bank's Linear[+BN]+act stack."""
        lay, prm, stats, rows = bank.lay, bank.params, bank.stats, bank.rows
        for l in range(lay.n_layers):
            din, dout = lay.dims[l], lay.dims[l + 1]
            W = prm[:, lay.w_off[l]: lay.w_off[l] + din * dout].view(rows, dout, din)
            b = prm[:, lay.b_off[l]: lay.b_off[l] + dout].view(rows, 1, dout)
            x = torch.baddbmm(b, x, W.transpose(1, 2))
            self.launches += 1
            if lay.bn[l]:
                gamma = prm[:, lay.bn_w_off[l]: lay.bn_w_off[l] + dout].view(rows, 1, dout)
                beta = prm[:, lay.bn_b_off[l]: lay.bn_b_off[l] + dout].view(rows, 1, dout)
                rm = stats[:, lay.bn_mean_off[l]: lay.bn_mean_off[l] + dout]
                rv = stats[:, lay.bn_var_off[l]: lay.bn_var_off[l] + dout]
                eps, mom = lay.desc.bn_eps, lay.desc.bn_momentum
                if self.training:
                    n = x.shape[1]
                    mean = x.mean(dim=1, keepdim=True)
                    var = x.var(dim=1, unbiased=False, keepdim=True)
                    with torch.no_grad():
                        rm.mul_(1 - mom).add_(mean.squeeze(1), alpha=mom)
                        rv.mul_(1 - mom).add_(var.squeeze(1) * (n / (n - 1)), alpha=mom)
                else:
                    mean, var = rm.unsqueeze(1), rv.unsqueeze(1)
                x = (x - mean) / torch.sqrt(var + eps) * gamma + beta
            a = lay.act[l]
            if a == abi.ACT_LRELU:
                x = torch.nn.functional.leaky_relu(x, lay.desc.lrelu_slope)
            elif a == abi.ACT_TANH:
                x = torch.tanh(x)
            elif a == abi.ACT_SIGMOID:
                x = torch.sigmoid(x)
        return x

    def forward(self, z):
        """z [S, B, 100] -> plain: [S, B, d]; multi-head: [S, N, B, d] (head i of server s feeds its
        i-th client, torch.chunk(net_g(z), N) in CGLGAN/2DMG/main.py:231)."""
        S, N = self.S, self.N
        hidden = self._run_stack(self.trunk, z)
        if N == 0:
            return hidden
        B = hidden.shape[1]
        hidden = _ScaleGrad.apply(hidden.unsqueeze(1).expand(S, N, B, hidden.shape[2]), self.trunk_scale)
        out = self._run_stack(self.heads, hidden.reshape(S * N, B, -1))
        return out.view(S, N, B, self.d)

    __call__ = forward

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    # ---- optimiser -------------------------------------------------------------------------------
    def zero_grad(self):
        for b in self.banks():
            b.params.grad = None

    def adam_step(self):
        """opti_g.step(): torch.optim.Adam semantics on every row (CGLGAN/2DMG/main.py:276)."""
        with torch.no_grad():
            for b in self.banks():
                g = b.params.grad
                assert g is not None, "backward() has not run"
                adam_rows(b.params.detach(), g.contiguous(), b.adam_m, b.adam_v, b.step, self.lr, self.b1, self.b2)
                self.launches += 2
