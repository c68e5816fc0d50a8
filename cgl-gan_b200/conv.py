"""The convolutional LSGAN networks of model/lsgan.py on the engine.

The reference ships `model/lsgan.py` (Generator :3-27, Discriminator :73-99) without a call site; a user who swaps them
into one of the scripts gets, per simulated client, a conv discriminator trained with the least-squares loss and, per
server, a conv generator. Here both are packed banks like the MLP ones (engine.ClientBank / generators.StackedGenerator):
rows of parameters() order, Adam moments and step counters, BatchNorm running statistics in a parallel row. Every 3x3
convolution runs as an implicit GEMM on the grouped Linear kernels (csrc/conv.cu explains the layout: activations are
[group][image*pixel][channel] rows; Conv2d.weight [Cout][Cin][3][3] is the Linear weight [out][in = Cin*9]); this module is
the host-side composition of those C-ABI calls -- there is no PyTorch / CPU implementation of the math here.

  ConvDiscriminatorBank.d_step   D_loss = MSE(D(real), 1) + MSE(D(fake), 0); backward; Adam   (the BCE body of Worker.train,
                                 CGLGAN/2DMG/main.py:357-366, with nn.MSELoss on the raw adv_layer output: LSGAN)
  ConvDiscriminatorBank.g_loss   G_loss = MSE(D(Xg), 1) and dG_loss/dXg                        (:368-373)
  ConvGeneratorStack.forward / backward_step                                                   (Server.train's generator part)
net_d(real) and net_d(fake) are two forward calls in the reference, so BatchNorm2d normalises each with its own batch
statistics (and updates the running statistics twice); the gradients of both meet in one optimizer step. Dropout2d(0.25)
masks are INJECTED ([G, B, C] per block, already scaled by 1 / 0.75; `sample_masks` draws them the way F.dropout2d does).
Full real batches only (n_real == batch size)."""
import ctypes as C

import torch

from . import abi

LRELU, TANH, NONE = abi.ACT_LRELU, abi.ACT_TANH, abi.ACT_NONE
BN_EPS, BN_MOM, SLOPE = 0.8, 0.1, 0.2


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pad32(n):
    return (n + 31) // 32 * 32


class _ConvBank:
    """Packed rows of a conv net: `entries` lists (name, numel) in parameters() order; every entry starts on a multiple of
    4 floats (the reference sizes do), `stats` lists (name, numel) of the BatchNorm running statistics."""

    def __init__(self, rows, entries, stats, device, lr, b1, b2, eps):
        self.rows, self.device = rows, torch.device(device)
        self.off, o = {}, 0
        for name, n in entries:
            self.off[name] = o
            o += n
        self.P = o
        self.ld = _pad32(o)
        self.soff, o = {}, 0
        for name, n in stats:
            self.soff[name] = o
            o += n
        self.n_stats = o
        self.ld_stats = _pad32(max(o, 1))
        self.params = torch.zeros(rows, self.ld, device=self.device)
        self.adam_m = torch.zeros_like(self.params)
        self.adam_v = torch.zeros_like(self.params)
        self.step = torch.zeros(rows, dtype=torch.int32, device=self.device)
        self.stats = torch.zeros(rows, self.ld_stats, device=self.device)
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps

    def load_modules(self, mods):
        assert len(mods) == self.rows
        p = torch.zeros(self.rows, self.ld)
        s = torch.zeros(self.rows, self.ld_stats)
        for r, m in enumerate(mods):
            flat = torch.cat([q.detach().reshape(-1) for q in m.parameters()]).float()
            assert flat.numel() == self.P, (flat.numel(), self.P)
            p[r, :self.P] = flat
            st = [b.reshape(-1) for n, b in m.named_buffers() if "running" in n]
            if st:
                st = torch.cat(st).float()
                s[r, :st.numel()] = st
        self.params.copy_(p)
        self.stats.copy_(s)

    def flat_rows(self):
        return self.params[:, :self.P]

    # ---- grouped Linear products on packed rows -------------------------------------------------------------------
    def lin_fwd(self, x, rows, cin, cout, w, b, act):
        G = self.rows
        y = torch.empty(G, rows, cout, device=self.device)
        abi.check(abi.lib.cgl_linear_fwd(G, rows, cin, cout, abi.ptr(x), rows * cin, abi.ptr(self.params), self.ld, None,
                                         self.off[w], self.off[b], act, SLOPE, abi.ptr(y), rows * cout, _st()))
        return y

    def lin_bwd_data(self, dy, rows, cin, cout, w):
        G = self.rows
        dx = torch.empty(G, rows, cin, device=self.device)
        abi.check(abi.lib.cgl_linear_bwd_data(G, rows, cin, cout, abi.ptr(dy), rows * cout, abi.ptr(self.params), self.ld, None,
                                              self.off[w], None, 0, NONE, 0.0, abi.ptr(dx), rows * cin, _st()))
        return dx

    def lin_wgrad_adam(self, dy, x, rows, cin, cout, w, b):
        abi.check(abi.lib.cgl_linear_wgrad_adam(self.rows, rows, cin, cout, abi.ptr(dy), rows * cout, abi.ptr(x), rows * cin,
                                                abi.ptr(self.params), abi.ptr(self.adam_m), abi.ptr(self.adam_v), self.ld,
                                                abi.ptr(self.step), None, self.off[w], self.off[b], self.lr, self.b1, self.b2,
                                                self.eps, None, _st()))

    def bn_fwd(self, u, rows, F, gname, bname, mname, vname, act, train=True):
        G = self.rows
        h = torch.empty_like(u)
        mean = torch.empty(G, F, device=self.device)
        invstd = torch.empty(G, F, device=self.device)
        abi.check(abi.lib.cgl_bn_forward(G, rows, F, abi.ptr(u), abi.ptr(h), abi.ptr(self.params), self.ld, None,
                                         self.off[gname], self.off[bname], abi.ptr(self.stats), self.ld_stats, self.soff[mname],
                                         self.soff[vname], abi.ptr(mean), abi.ptr(invstd), BN_EPS, BN_MOM, 1 if train else 0,
                                         act, SLOPE, _st()))
        return h, mean, invstd


def im2col(x, N, H, W, Cc, stride):
    OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
    col = torch.empty(N * OH * OW, Cc * 9, device=x.device)
    abi.check(abi.lib.cgl_im2col3x3(N, H, W, Cc, stride, abi.ptr(x), abi.ptr(col), _st()))
    return col, OH, OW


def col2im(dcol, N, H, W, Cc, stride):
    dx = torch.empty(N * H * W, Cc, device=dcol.device)
    abi.check(abi.lib.cgl_col2im3x3(N, H, W, Cc, stride, abi.ptr(dcol), abi.ptr(dx), _st()))
    return dx


def act_backward(dy, y, act):
    dz = torch.empty_like(dy)
    abi.check(abi.lib.cgl_act_backward(dy.numel(), abi.ptr(dy), abi.ptr(y), abi.ptr(dz), act, SLOPE, _st()))
    return dz


def sample_masks(G, B, train=True, generator=None, device="cuda", p=0.25):
    """Dropout2d masks of the four discriminator blocks, drawn like F.dropout2d does: noise [B, C, 1, 1] = bernoulli(1-p) / (1-p)."""
    out = []
    for Cc in (16, 32, 64, 128):
        if train:
            m = torch.bernoulli(torch.full((G, B, Cc), 1 - p), generator=generator) / (1 - p)
        else:
            m = torch.ones(G, B, Cc)
        out.append(m.to(device))
    return out


class ConvDiscriminatorBank(_ConvBank):
    """model/lsgan.py:73-99 for G clients: 4 x [Conv2d(3, 2, 1), LeakyReLU(0.2), Dropout2d(0.25), BatchNorm2d(C, 0.8) (not in
    block 1)], Linear(128 * 2 * 2, 1) on 1 x 32 x 32 images."""
    CH = (1, 16, 32, 64, 128)
    CONV = ("model.0", "model.3", "model.7", "model.11")
    BN = (None, "model.6", "model.10", "model.14")

    def __init__(self, n_clients, batch_size, device="cuda", lr=0.0002, b1=0.5, b2=0.999, eps=1e-8):
        abi.require_device()
        ent, st = [], []
        for k in range(4):
            ci, co = self.CH[k], self.CH[k + 1]
            ent += [(self.CONV[k] + ".weight", co * ci * 9), (self.CONV[k] + ".bias", co)]
            if self.BN[k]:
                ent += [(self.BN[k] + ".weight", co), (self.BN[k] + ".bias", co)]
                st += [(self.BN[k] + ".running_mean", co), (self.BN[k] + ".running_var", co)]
        ent += [("adv_layer.weight", 512), ("adv_layer.bias", 1)]
        super().__init__(n_clients, ent, st, device, lr, b1, b2, eps)
        self.B = batch_size

    def _forward(self, img, masks, train=True):
        """img [G, B, 1024] -> logits [G, B]; returns the tensors the backward needs."""
        G, B = self.rows, img.shape[1]
        x = img.reshape(G * B * 1024, 1).contiguous()
        H = 32
        saved = []
        for k in range(4):
            ci, co = self.CH[k], self.CH[k + 1]
            col, OH, OW = im2col(x, G * B, H, H, ci, 2)
            rows = B * OH * OW
            a = self.lin_fwd(col, rows, ci * 9, co, self.CONV[k] + ".weight", self.CONV[k] + ".bias", LRELU)
            abi.check(abi.lib.cgl_channel_scale(G * B, OH * OW, co, abi.ptr(masks[k]), abi.ptr(a), _st()))   # Dropout2d, in place
            if self.BN[k]:
                h, mean, invstd = self.bn_fwd(a, rows, co, self.BN[k] + ".weight", self.BN[k] + ".bias",
                                              self.BN[k] + ".running_mean", self.BN[k] + ".running_var", NONE, train)
            else:
                h, mean, invstd = a, None, None
            saved.append(dict(x=x, H=H, a=a, mean=mean, invstd=invstd, rows=rows))
            x, H = h.reshape(G * rows, co), OH
        flat = torch.empty(G * B, 512, device=self.device)                # out.view(B, -1) of [B, 128, 2, 2]
        abi.check(abi.lib.cgl_nhwc_to_nchw(G * B, 128, 4, abi.ptr(x), abi.ptr(flat), _st()))
        logits = self.lin_fwd(flat, B, 512, 1, "adv_layer.weight", "adv_layer.bias", NONE)
        return logits.reshape(G, B), flat, saved

    def _backward(self, passes, train_step):
        """passes: [(dlogits [G, B], flat, saved, masks)] of one or two forward calls. train_step: Adam on every parameter
        (the gradients of the passes are summed); returns dLoss/dimg [G, B, 1024] of the LAST pass."""
        G = self.rows
        Bs = [p[0].shape[1] for p in passes]
        cat = (lambda ts: ts[0] if len(ts) == 1 else torch.cat(ts, dim=1).contiguous())
        dl = cat([p[0].reshape(G, b, 1) for p, b in zip(passes, Bs)])
        flat = cat([p[1].reshape(G, b, 512) for p, b in zip(passes, Bs)])
        rows = sum(Bs)
        dflat = self.lin_bwd_data(dl, rows, 512, 1, "adv_layer.weight")
        if train_step:
            self.lin_wgrad_adam(dl, flat, rows, 512, 1, "adv_layer.weight", "adv_layer.bias")
        # back to NHWC per pass
        g_list, o = [], 0
        for b in Bs:
            d = torch.empty(G * b * 4, 128, device=self.device)
            seg = dflat[:, o:o + b].contiguous()
            abi.check(abi.lib.cgl_nchw_to_nhwc(G * b, 128, 4, abi.ptr(seg), abi.ptr(d), _st()))
            g_list.append(d)
            o += b
        for k in (3, 2, 1, 0):
            ci, co = self.CH[k], self.CH[k + 1]
            sv = [p[2][k] for p in passes]
            rws = [s["rows"] for s in sv]
            g_cat = cat([g.reshape(G, r, co) for g, r in zip(g_list, rws)])
            a_cat = cat([s["a"].reshape(G, r, co) for s, r in zip(sv, rws)])
            if self.BN[k]:
                if train_step:
                    m1 = sv[1]["mean"] if len(sv) > 1 else None
                    i1 = sv[1]["invstd"] if len(sv) > 1 else None
                    abi.check(abi.lib.cgl_bn_backward_seg(G, rws[0], rws[1] if len(rws) > 1 else 0, co, abi.ptr(g_cat), abi.ptr(a_cat),
                                                          abi.ptr(sv[0]["mean"]), abi.ptr(sv[0]["invstd"]), abi.ptr(m1), abi.ptr(i1),
                                                          abi.ptr(self.params), abi.ptr(self.adam_m), abi.ptr(self.adam_v), self.ld,
                                                          None, self.off[self.BN[k] + ".weight"], self.off[self.BN[k] + ".bias"],
                                                          abi.ptr(self.step), self.lr, self.b1, self.b2, self.eps, _st()))
                else:
                    g_cat = self._bn_bwd_data_only(g_cat, a_cat, sv[0], rws[0], co, self.BN[k] + ".weight")
            # Dropout2d backward (same mask), then LeakyReLU'(a): sign(a * mask) == sign(a) wherever the mask is not zero
            o = 0
            for p, s, r, b in zip(passes, sv, rws, Bs):
                seg = g_cat[:, o:o + r]
                if len(passes) > 1:
                    seg = seg.contiguous()
                abi.check(abi.lib.cgl_channel_scale(G * b, r // b, co, abi.ptr(p[3][k]), abi.ptr(seg), _st()))
                if len(passes) > 1:
                    g_cat[:, o:o + r] = seg
                o += r
            dconv = act_backward(g_cat, a_cat, LRELU)
            # x_col of every pass again (recomputed from the saved block input: an im2col buffer is 9 x its input)
            cols = []
            for s, b in zip(sv, Bs):
                col, _, _ = im2col(s["x"], G * b, s["H"], s["H"], ci, 2)
                cols.append(col.reshape(G, s["rows"], ci * 9))
            col_cat = cat(cols)
            tot = sum(rws)
            dcol = self.lin_bwd_data(dconv, tot, ci * 9, co, self.CONV[k] + ".weight") if k > 0 or not train_step else None
            if train_step:
                self.lin_wgrad_adam(dconv, col_cat, tot, ci * 9, co, self.CONV[k] + ".weight", self.CONV[k] + ".bias")
            if dcol is None:
                return None
            g_list, o = [], 0
            for s, b, r in zip(sv, Bs, rws):
                seg = dcol[:, o:o + r].contiguous()
                g_list.append(col2im(seg.reshape(G * r, ci * 9), G * b, s["H"], s["H"], ci, 2))
                o += r
        return g_list[-1].reshape(G, Bs[-1], 1024)

    def _bn_bwd_data_only(self, g, a, sv, rows, F, gname):
        """BatchNorm backward without a parameter update (the generator-loss pass): through scratch copies of the affine row."""
        scratch_p = self.params.clone()
        scratch_m, scratch_v = torch.zeros_like(self.params), torch.zeros_like(self.params)
        one = torch.ones(self.rows, dtype=torch.int32, device=self.device)
        bname = gname.replace(".weight", ".bias")
        g = g.contiguous()
        abi.check(abi.lib.cgl_bn_backward(self.rows, rows, F, abi.ptr(g), abi.ptr(a), abi.ptr(sv["mean"]), abi.ptr(sv["invstd"]),
                                          abi.ptr(scratch_p), abi.ptr(scratch_m), abi.ptr(scratch_v), self.ld, None, self.off[gname],
                                          self.off[bname], abi.ptr(one), 0.0, self.b1, self.b2, self.eps, _st()))
        return g

    def d_step(self, real, fake, masks_real, masks_fake):
        """real, fake [G, B, 1024]; masks_*: sample_masks(G, B). One Adam step; returns D_loss [G]."""
        G, B = self.rows, self.B
        lr_, fr, sr = self._forward(real.reshape(G, B, 1024), masks_real)
        lf, ff, sf = self._forward(fake.reshape(G, B, 1024), masks_fake)
        loss = ((lr_ - 1.0) ** 2).mean(1) + (lf ** 2).mean(1)          # nn.MSELoss()(D(real), 1) + nn.MSELoss()(D(fake), 0)
        d_real = 2.0 * (lr_ - 1.0) / B
        d_fake = 2.0 * lf / B
        self.step += 1                                                 # the fused Adam epilogues read the incremented counter
        self._backward([(d_real.contiguous(), fr, sr, masks_real), (d_fake.contiguous(), ff, sf, masks_fake)], train_step=True)
        return loss

    def g_loss(self, xg, masks):
        """G_loss = MSE(D(Xg), 1) through the (updated) discriminators, train mode like the reference's net_d(Xg); returns
        (loss [G], dG_loss/dXg [G, B, 1024])."""
        G, B = self.rows, xg.shape[1]
        lg, fg, sg = self._forward(xg.reshape(G, B, 1024), masks)
        loss = ((lg - 1.0) ** 2).mean(1)
        dl = (2.0 * (lg - 1.0) / B).contiguous()
        dx = self._backward([(dl, fg, sg, masks)], train_step=False)
        return loss, dx


class ConvGeneratorStack(_ConvBank):
    """model/lsgan.py:3-27 for S servers: Linear(100, 128 * 8 * 8) -> [128, 8, 8] -> Upsample, Conv(128, 128), BN, LeakyReLU,
    Upsample, Conv(128, 64), BN, LeakyReLU, Conv(64, 1), Tanh -> 1 x 32 x 32."""

    def __init__(self, n_servers, device="cuda", lr=0.0002, b1=0.5, b2=0.999, eps=1e-8):
        abi.require_device()
        ent = [("l1.0.weight", 8192 * 100), ("l1.0.bias", 8192),
               ("conv_blocks.1.weight", 128 * 128 * 9), ("conv_blocks.1.bias", 128),
               ("conv_blocks.2.weight", 128), ("conv_blocks.2.bias", 128),
               ("conv_blocks.5.weight", 64 * 128 * 9), ("conv_blocks.5.bias", 64),
               ("conv_blocks.6.weight", 64), ("conv_blocks.6.bias", 64),
               ("conv_blocks.8.weight", 64 * 9), ("conv_blocks.8.bias", 1)]
        st = [("conv_blocks.2.running_mean", 128), ("conv_blocks.2.running_var", 128),
              ("conv_blocks.6.running_mean", 64), ("conv_blocks.6.running_var", 64)]
        super().__init__(n_servers, ent, st, device, lr, b1, b2, eps)
        self._last = None

    def forward(self, z, train=True):
        """z [S, B, 100] -> images [S, B, 1024] (1 x 32 x 32). Train mode: batch statistics, running statistics updated."""
        S, B = self.rows, z.shape[1]
        z = z.reshape(S, B, 100).contiguous().float()
        l1 = self.lin_fwd(z, B, 100, 8192, "l1.0.weight", "l1.0.bias", NONE)           # [S, B, 128*64] NCHW
        x0 = torch.empty(S * B * 64, 128, device=self.device)
        abi.check(abi.lib.cgl_nchw_to_nhwc(S * B, 128, 64, abi.ptr(l1), abi.ptr(x0), _st()))
        up1 = torch.empty(S * B * 256, 128, device=self.device)
        abi.check(abi.lib.cgl_upsample2x(S * B, 8, 8, 128, abi.ptr(x0), abi.ptr(up1), _st()))
        col1, _, _ = im2col(up1, S * B, 16, 16, 128, 1)
        u1 = self.lin_fwd(col1, B * 256, 1152, 128, "conv_blocks.1.weight", "conv_blocks.1.bias", NONE)
        del col1
        h1, m1, i1 = self.bn_fwd(u1, B * 256, 128, "conv_blocks.2.weight", "conv_blocks.2.bias", "conv_blocks.2.running_mean",
                                 "conv_blocks.2.running_var", LRELU, train)
        up2 = torch.empty(S * B * 1024, 128, device=self.device)
        abi.check(abi.lib.cgl_upsample2x(S * B, 16, 16, 128, abi.ptr(h1), abi.ptr(up2), _st()))
        col2, _, _ = im2col(up2, S * B, 32, 32, 128, 1)
        u2 = self.lin_fwd(col2, B * 1024, 1152, 64, "conv_blocks.5.weight", "conv_blocks.5.bias", NONE)
        del col2
        h2, m2, i2 = self.bn_fwd(u2, B * 1024, 64, "conv_blocks.6.weight", "conv_blocks.6.bias", "conv_blocks.6.running_mean",
                                 "conv_blocks.6.running_var", LRELU, train)
        col3, _, _ = im2col(h2, S * B, 32, 32, 64, 1)
        img = self.lin_fwd(col3, B * 1024, 576, 1, "conv_blocks.8.weight", "conv_blocks.8.bias", TANH)
        del col3
        self._last = dict(z=z, B=B, up1=up1, u1=u1, h1=h1, m1=m1, i1=i1, up2=up2, u2=u2, h2=h2, m2=m2, i2=i2, img=img)
        return img.reshape(S, B, 1024)

    __call__ = forward

    def backward_step(self, dy):
        """dy = dLoss/d(images of the latest forward) [S, B, 1024]; one Adam step on every generator parameter."""
        L = self._last
        assert L is not None, "forward() has not run"
        S, B = self.rows, L["B"]
        self.step += 1
        bn_args = (self.lr, self.b1, self.b2, self.eps)
        d3 = act_backward(dy.reshape(S, B * 1024, 1).contiguous().float(), L["img"], TANH)
        col3, _, _ = im2col(L["h2"], S * B, 32, 32, 64, 1)
        dcol3 = self.lin_bwd_data(d3, B * 1024, 576, 1, "conv_blocks.8.weight")
        self.lin_wgrad_adam(d3, col3, B * 1024, 576, 1, "conv_blocks.8.weight", "conv_blocks.8.bias")
        del col3
        dh2 = col2im(dcol3.reshape(S * B * 1024, 576), S * B, 32, 32, 64, 1)
        del dcol3
        dz2 = act_backward(dh2.reshape(S, B * 1024, 64), L["h2"], LRELU)
        abi.check(abi.lib.cgl_bn_backward(S, B * 1024, 64, abi.ptr(dz2), abi.ptr(L["u2"]), abi.ptr(L["m2"]), abi.ptr(L["i2"]),
                                          abi.ptr(self.params), abi.ptr(self.adam_m), abi.ptr(self.adam_v), self.ld, None,
                                          self.off["conv_blocks.6.weight"], self.off["conv_blocks.6.bias"], abi.ptr(self.step),
                                          *bn_args, _st()))
        col2, _, _ = im2col(L["up2"], S * B, 32, 32, 128, 1)
        dcol2 = self.lin_bwd_data(dz2, B * 1024, 1152, 64, "conv_blocks.5.weight")
        self.lin_wgrad_adam(dz2, col2, B * 1024, 1152, 64, "conv_blocks.5.weight", "conv_blocks.5.bias")
        del col2
        dup2 = col2im(dcol2.reshape(S * B * 1024, 1152), S * B, 32, 32, 128, 1)
        del dcol2
        dh1 = torch.empty(S * B * 256, 128, device=self.device)
        abi.check(abi.lib.cgl_upsample2x_bwd(S * B, 16, 16, 128, abi.ptr(dup2), abi.ptr(dh1), _st()))
        dz1 = act_backward(dh1.reshape(S, B * 256, 128), L["h1"], LRELU)
        abi.check(abi.lib.cgl_bn_backward(S, B * 256, 128, abi.ptr(dz1), abi.ptr(L["u1"]), abi.ptr(L["m1"]), abi.ptr(L["i1"]),
                                          abi.ptr(self.params), abi.ptr(self.adam_m), abi.ptr(self.adam_v), self.ld, None,
                                          self.off["conv_blocks.2.weight"], self.off["conv_blocks.2.bias"], abi.ptr(self.step),
                                          *bn_args, _st()))
        col1, _, _ = im2col(L["up1"], S * B, 16, 16, 128, 1)
        dcol1 = self.lin_bwd_data(dz1, B * 256, 1152, 128, "conv_blocks.1.weight")
        self.lin_wgrad_adam(dz1, col1, B * 256, 1152, 128, "conv_blocks.1.weight", "conv_blocks.1.bias")
        del col1
        dup1 = col2im(dcol1.reshape(S * B * 256, 1152), S * B, 16, 16, 128, 1)
        dx0 = torch.empty(S * B * 64, 128, device=self.device)
        abi.check(abi.lib.cgl_upsample2x_bwd(S * B, 8, 8, 128, abi.ptr(dup1), abi.ptr(dx0), _st()))
        dl1 = torch.empty(S, B, 8192, device=self.device)
        abi.check(abi.lib.cgl_nhwc_to_nchw(S * B, 128, 64, abi.ptr(dx0), abi.ptr(dl1), _st()))
        self.lin_wgrad_adam(dl1, L["z"], B, 100, 8192, "l1.0.weight", "l1.0.bias")
        self._last = None
