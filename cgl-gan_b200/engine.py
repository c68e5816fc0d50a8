"""ClientBank: all simulated clients' discriminators as packed rows in HBM, stepped through the C ABI.

Replaces the per-thread `Worker` objects of the reference (CGLGAN/2DMG/main.py:282-375): instead of one
Python thread + one nn.Module + one optim.Adam per client, one bank holds [C, ld] fp32 rows for the
parameters and both Adam moments plus a per-client step counter, and one call advances every client.
"""
import ctypes as C

import torch

from . import abi
from .layout import RowLayout, flatten_params, load_flat_params


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _i32(t, device):
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.int32).contiguous()


class ClientBank:
    def __init__(self, arch, n_clients, batch_size, device="cuda", loss_kind=abi.LOSS_BCE, d_loss_scale=1.0,
                 lr=0.0002, b1=0.5, b2=0.999, eps=1e-8):
        abi.require_device()
        self.desc = abi.arch_describe(arch) if isinstance(arch, int) else arch
        self.lay = RowLayout(self.desc)
        self.C = int(n_clients)
        self.B = int(batch_size)
        self.d = self.lay.dims[0]
        self.device = torch.device(device)
        self.ld = self.lay.ld
        self.P = self.lay.n_params
        self.params = torch.zeros(self.C, self.ld, device=self.device)
        self.adam_m = torch.zeros(self.C, self.ld, device=self.device)
        self.adam_v = torch.zeros(self.C, self.ld, device=self.device)
        self.step = torch.zeros(self.C, dtype=torch.int32, device=self.device)
        self.cfg = abi.TrainCfg(loss_kind, d_loss_scale, lr, b1, b2, eps)
        self._ws = None
        self._scratch = None  # second parameter buffer for out-of-place mixing
        self.launches = 0     # kernels launched through this bank (bench bookkeeping)

    # ---- parameter I/O -------------------------------------------------------------------------
    def load_modules(self, modules):
        """Row c <- parameters of modules[c] (reference Discriminator instances or ours)."""
        assert len(modules) == self.C
        rows = torch.stack([flatten_params(m).float().cpu() for m in modules])
        self.load_rows(rows)

    def load_rows(self, rows):
        assert rows.shape == (self.C, self.P), (rows.shape, (self.C, self.P))
        self.params.zero_()
        self.params[:, :self.P].copy_(rows)

    def rows(self):
        return self.params[:, :self.P]

    def store_module(self, c, module):
        load_flat_params(module, self.params[c, :self.P].cpu())

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- K1/K2 -----------------------------------------------------------------------------------
    def d_step(self, real, fake, n_real=None, fake_idx=None, client_ids=None):
        """One Adam step of every listed client's D on (real[g], fake[fake_idx[g]]).
        Reference: Worker.train D loop, CGLGAN/2DMG/main.py:357-366; capgan.py:329-341.
        real [G,B,d] fp32 (rows >= n_real[g] must be finite, they carry no gradient), fake [F,B,d]."""
        G = real.shape[0]
        B = self.B
        real = real.reshape(G, B, self.d).contiguous()
        fake = fake.reshape(-1, B, self.d).contiguous()
        assert real.dtype == torch.float32 and fake.dtype == torch.float32
        n_real = _i32(n_real, self.device)
        fake_idx = _i32(fake_idx, self.device)
        client_ids = _i32(client_ids, self.device)
        if fake_idx is None:
            assert fake.shape[0] >= G
        out = torch.empty(G, device=self.device)
        nbytes = abi.lib.cgl_d_step_workspace_bytes(C.byref(self.desc), G, B)
        ws = self._workspace(nbytes)
        abi.check(abi.lib.cgl_d_step(C.byref(self.desc), G, abi.ptr(self.params), abi.ptr(self.adam_m),
                                     abi.ptr(self.adam_v), self.ld, abi.ptr(self.step), abi.ptr(client_ids),
                                     abi.ptr(real), abi.ptr(n_real), abi.ptr(fake), abi.ptr(fake_idx), B,
                                     C.byref(self.cfg), abi.ptr(out), abi.ptr(ws), ws.numel(), _stream()))
        self.launches += 3 * self.desc.n_layers - 2   # bump, L-1 fwd, head, L-2 bwd-data, L-1 wgrad+Adam
        return out

    def g_loss_raw(self, xg, xg_idx=None, client_ids=None, need_grad=True):
        """loss[g] = loss(D_g(xg[xg_idx[g]]), valid) and dloss/dxg [G,B,d].
        Reference: Worker.train tail, CGLGAN/2DMG/main.py:368-373."""
        B = self.B
        xg = xg.reshape(-1, B, self.d).contiguous()
        xg_idx = _i32(xg_idx, self.device)
        client_ids = _i32(client_ids, self.device)
        G = xg_idx.numel() if xg_idx is not None else (client_ids.numel() if client_ids is not None else xg.shape[0])
        loss = torch.empty(G, device=self.device)
        dxg = torch.empty(G, B, self.d, device=self.device) if need_grad else None
        nbytes = abi.lib.cgl_g_loss_workspace_bytes(C.byref(self.desc), G, B)
        ws = self._workspace(nbytes)
        abi.check(abi.lib.cgl_g_loss(C.byref(self.desc), G, abi.ptr(self.params), self.ld, abi.ptr(client_ids),
                                     abi.ptr(xg), abi.ptr(xg_idx), B, self.cfg.loss_kind, abi.ptr(loss),
                                     abi.ptr(dxg), abi.ptr(ws), ws.numel(), _stream()))
        self.launches += self.desc.n_layers + (self.desc.n_layers - 1 if need_grad else 0)
        return loss, dxg

    def client_step(self, real, fake, xg, n_real=None, idx=None, client_ids=None, need_grad=True):
        """d_step(real, fake[idx]) followed by g_loss_raw(xg[idx]) in ONE ABI call (cgl_client_step): one Worker.train
        call with epoch == 1, CGLGAN/2DMG/main.py:344-375. idx None: client g owns fake[g] / xg[g].
        For the 2DMG discriminator this is one kernel launch with the client's weights resident in shared memory."""
        G, B = real.shape[0], self.B
        real = real.reshape(G, B, self.d).contiguous()
        fake = fake.reshape(-1, B, self.d).contiguous()
        xg = xg.reshape(-1, B, self.d).contiguous()
        assert real.dtype == torch.float32 and fake.dtype == torch.float32 and xg.dtype == torch.float32
        n_real, idx, client_ids = _i32(n_real, self.device), _i32(idx, self.device), _i32(client_ids, self.device)
        if idx is None:
            assert fake.shape[0] >= G and xg.shape[0] >= G
        d_loss = torch.empty(G, device=self.device)
        g_loss = torch.empty(G, device=self.device)
        dxg = torch.empty(G, B, self.d, device=self.device) if need_grad else None
        nbytes = abi.lib.cgl_d_step_workspace_bytes(C.byref(self.desc), G, B)
        ws = self._workspace(nbytes)
        abi.check(abi.lib.cgl_client_step(C.byref(self.desc), G, abi.ptr(self.params), abi.ptr(self.adam_m),
                                          abi.ptr(self.adam_v), self.ld, abi.ptr(self.step), abi.ptr(client_ids),
                                          abi.ptr(real), abi.ptr(n_real), abi.ptr(fake), abi.ptr(idx), abi.ptr(xg),
                                          abi.ptr(idx), B, C.byref(self.cfg), abi.ptr(d_loss), abi.ptr(g_loss),
                                          abi.ptr(dxg), abi.ptr(ws), ws.numel(), _stream()))
        return d_loss, g_loss, dxg

    def g_loss(self, xg, xg_idx=None, client_ids=None):
        """Graph-attached client losses: the tensor the reference's workers put on servers[id].queen_g
        (CGLGAN/2DMG/main.py:373). Backward delivers sum_g grad[g] * dloss_g/dxg to xg."""
        return _GLossFn.apply(xg, self, xg_idx, client_ids)

    # ---- K3 ---------------------------------------------------------------------------------------
    def _other(self):
        if self._scratch is None:
            self._scratch = torch.zeros_like(self.params)
        return self._scratch

    def mix(self, matrix):
        """params <- M @ params for a dense [C,C] (or CSR triple (row_ptr, col, vals)) mixing matrix: FedAvg rows, swap
        permutations, neighbour / group means (a8, a10); vals None = the mean of the listed rows, summed in column order
        and divided once (receive_parameter, CGLGAN/2DMG/main.py:171-179). Out of place, then the buffers swap roles."""
        row_ptr, col, vals = dense_to_csr(matrix) if torch.is_tensor(matrix) else matrix
        row_ptr, col = _i32(row_ptr, self.device), _i32(col, self.device)
        vals = None if vals is None else vals.to(self.device, torch.float32).contiguous()   # None: row means (cgl_mix_csr)
        R = row_ptr.numel() - 1
        assert R == self.C
        dst = self._other()
        abi.check(abi.lib.cgl_mix_csr(R, self.ld, abi.ptr(row_ptr), abi.ptr(col), abi.ptr(vals), abi.ptr(self.params),
                                      self.ld, abi.ptr(dst), self.ld, _stream()))
        self.params, self._scratch = dst, self.params
        self.launches += 1

    def weighted_sum(self, weights, rows=None, out=None):
        """out[:] = sum_c weights[c] * params[rows[c]] in ascending c (Cloud.run / fedavg_aggregate)."""
        w = weights.to(self.device, torch.float32).contiguous()
        rows = _i32(rows, self.device)
        if out is None:
            out = torch.empty(self.ld, device=self.device)
        abi.check(abi.lib.cgl_wsum(w.numel(), self.ld, abi.ptr(w), abi.ptr(rows), abi.ptr(self.params), self.ld,
                                   abi.ptr(out), _stream()))
        self.launches += 1
        return out

    def broadcast(self, g, sigma=0.0, rows=None):
        """params[rows] <- sigma*params[rows] + (1-sigma)*g (segema mix + load_state_dict)."""
        rows = _i32(rows, self.device)
        R = rows.numel() if rows is not None else self.C
        abi.check(abi.lib.cgl_bcast_mix(R, self.ld, abi.ptr(rows), float(sigma), abi.ptr(g), abi.ptr(self.params),
                                        self.ld, _stream()))
        self.launches += 1


class _GLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xg, bank, xg_idx, client_ids):
        loss, dxg = bank.g_loss_raw(xg.detach(), xg_idx, client_ids, need_grad=True)
        ctx.bank = bank
        ctx.xg_shape = xg.shape
        ctx.xg_idx = xg_idx
        ctx.save_for_backward(dxg)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        (dxg,) = ctx.saved_tensors
        bank = ctx.bank
        G = dxg.shape[0]
        n = dxg.shape[1] * dxg.shape[2]
        F = 1
        for s in ctx.xg_shape:
            F *= s
        F //= n
        grad_loss = grad_loss.contiguous().float()
        if ctx.xg_idx is None:
            # every client owns its slice of xg (multi-head generators): scale in place of a reduce
            gx = dxg * grad_loss.view(G, 1, 1)
            return gx.view(ctx.xg_shape), None, None, None
        # shared xg: grad[f] = sum over clients with xg_idx == f, in client order (deterministic)
        idx = torch.as_tensor(ctx.xg_idx).to("cpu", torch.int64)
        order = torch.argsort(idx, stable=True)
        counts = torch.bincount(idx, minlength=F)
        srv_ptr = torch.zeros(F + 1, dtype=torch.int32)
        srv_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        clients = order.to(torch.int32)
        out = torch.empty(F, n, device=dxg.device)
        srv_ptr_d, clients_d = srv_ptr.to(dxg.device), clients.to(dxg.device)
        abi.check(abi.lib.cgl_dxg_reduce(F, abi.ptr(srv_ptr_d), abi.ptr(clients_d), abi.ptr(grad_loss), abi.ptr(dxg),
                                         n, abi.ptr(out), _stream()))
        bank.launches += 1
        return out.view(ctx.xg_shape), None, None, None


def dense_to_csr(M):
    """Dense mixing matrix -> (row_ptr, col, vals) keeping column order (the accumulation order)."""
    M = M.detach().cpu().float()
    nz = M != 0
    counts = nz.sum(1)
    row_ptr = torch.zeros(M.shape[0] + 1, dtype=torch.int32)
    row_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    rows, cols = torch.nonzero(nz, as_tuple=True)
    return row_ptr, cols.to(torch.int32), M[rows, cols].contiguous()


def adam_rows(p, g, m, v, step, lr, b1, b2, eps=1e-8):
    """Fused torch.optim.Adam step over packed rows [R, ld] (server-side generators, a7)."""
    R, ld = p.shape
    assert g.shape == p.shape and m.shape == p.shape and v.shape == p.shape
    assert p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous()
    abi.check(abi.lib.cgl_adam_rows(R, ld, ld, abi.ptr(p), abi.ptr(g), abi.ptr(m), abi.ptr(v), abi.ptr(step),
                                    lr, b1, b2, eps, _stream()))
