"""Whole communication rounds of the reference, restated serially on the CPU (servers 0..S-1, then their
clients in index order). One object per simulation; every random input is injected.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

MD-style round = one iteration of Server.run's while-loop (CGLGAN/2DMG/main.py:196-212) for every server,
with the Worker.train calls it triggers (:344-375) and the Cloud.run iteration it feeds (:124-136).
FL-style round = one iteration of FL Server.run (FLGAN/MNIST/flgan.py:141-163) with Worker.run/train (:220-270).
"""
import copy

import torch
from torch import optim

from . import models as om
from . import steps as st


class OracleMD:
    def __init__(self, algo, num_workers, num_servers, batch_size, img_shape, iid=1, part_sizes=None,
                 segema=0.0, cloud_epoch=1, cloud_mode="intended", lr=0.0002, b1=0.5, b2=0.999, weights_init=False,
                 num_communication=20000, E=0, d_share="group_mean"):
        self.algo = algo
        self.S, self.N = num_servers, num_workers // num_servers
        self.C = self.S * self.N
        self.B = batch_size
        self.segema, self.cloud_epoch, self.cloud_mode = segema, cloud_epoch, cloud_mode
        self.num_communication, self.E, self.d_share = num_communication, E, d_share
        d = 1
        for s in img_shape:
            d *= s
        two_d = d == 2
        N = self.N
        if algo == "cglgan":
            heads = N if iid != 0 else 1
            mk_g = (lambda: om.Generator2DCGL(img_shape, heads)) if two_d else (lambda: om.MixGeneratorMNIST(img_shape, heads))
            self.multi_head = iid != 0
        elif algo == "mixed":
            mk_g = lambda: om.MixGeneratorMNIST(img_shape, N)
            self.multi_head = True
        else:
            mk_g = (lambda: om.Generator2DMD(img_shape)) if two_d else (lambda: om.GeneratorMNIST(img_shape))
            self.multi_head = False
        if two_d:
            self.kind, mk_d = st.LOSS_BCE, lambda: om.Discriminator2D()
        elif algo in ("capgan", "capgan_copy", "mixed", "acgan"):
            self.kind, mk_d = st.LOSS_CE, lambda: om.DiscriminatorMNIST2(img_shape)
        else:
            self.kind, mk_d = st.LOSS_BCE, lambda: om.DiscriminatorMNIST1(img_shape)
        self.d_scale = 0.5 if (algo in ("capgan", "capgan_copy", "mixed") and not two_d) else 1.0
        self.net_g = [mk_g() for _ in range(self.S)]
        self.net_d = [mk_d() for _ in range(self.C)]
        if weights_init:
            for m in self.net_g + self.net_d:
                m.apply(om.weights_init)
        self.opti_g = [st.make_adam(g.parameters(), lr, b1, b2) for g in self.net_g]
        self.opti_d = [st.make_adam(d_.parameters(), lr, b1, b2) for d_ in self.net_d]
        self.loss = st.make_loss(self.kind)
        if part_sizes is None:
            part_sizes = [1] * self.C
        self.beta, self.data_len = [], torch.zeros(self.S)
        for s in range(self.S):
            b = torch.zeros(N)
            for c in range(N):
                b[c] = part_sizes[s * N + c]          # CGLGAN/2DMG/main.py:184-188
            self.data_len[s] = b.sum()
            self.beta.append(b / b.sum())
        self.A = self.data_len / self.data_len.sum()   # Cloud.run, main.py:117-122
        self.Lambda = [torch.tensor(0., requires_grad=algo in ("capgan", "capgan_copy", "mixed")) for _ in range(self.S)]
        self.opti_L = [optim.SGD([L], lr=0.1) for L in self.Lambda]   # main.py:159-160
        self.t = 0
        self.F_max = [None] * self.S

    def cloud(self):
        """Server.run's cloud block + Cloud.run (CGLGAN/2DMG/main.py:201-208,124-136); capgan.py:170-175,110-117."""
        if self.algo in ("capgan", "capgan_copy"):
            flats = [st.serialize_model(g) for g in self.net_g]
            para_sum = st.fedavg_aggregate(flats, self.data_len.clone())
            for s, g in enumerate(self.net_g):
                st.deserialize_model(g, self.segema * flats[s] + (1 - self.segema) * para_sum)
            return
        if self.algo in ("mdgan", "acgan"):
            return   # no cloud in these scripts
        if self.cloud_mode == "as_written":
            return   # net_g.load_state_dict(recv_p, strict=False) with trunk-relative keys loads nothing
        self_ps = [st.copy_parameters(g.model) for g in self.net_g]
        p = st.cloud_aggregate([copy.deepcopy(x) for x in self_ps], self.A)
        for s, g in enumerate(self.net_g):
            recv = st.segema_mix(self_ps[s], {k: v.clone() for k, v in p.items()}, self.segema)
            g.model.load_state_dict(recv, strict=False)   # the intended target of main.py:208

    def share(self):
        """Neighbour-D share (commented out as shipped, README.md:26 "cancel the note to test").
        swap      : MDGAN/MNIST/mdgan.py:158-164,258-262 -- the server collects its clients' D dicts in client order,
                    self.rd.shuffle(p_ds) (Random(rank + 100), :122-123), client idx loads p_ds[idx].
        group_mean: ACGAN/MNIST/acgan.py:240-263 -- w <- p; s <- mean of the group's w; p += s - w, i.e. every client of the
                    server's group ends with the uniform mean (== the dead receive_parameter of CGLGAN/2DMG/main.py:171-179)."""
        import random
        if not hasattr(self, "rd"):
            self.rd = [random.Random(s + 100) for s in range(self.S)]
        N = self.N
        for s in range(self.S):
            cl = list(range(s * N, (s + 1) * N))
            dicts = [st.copy_parameters(self.net_d[c]) for c in cl]
            if self.d_share == "swap":
                p_ds = st.mdgan_swap(dicts, self.rd[s])
                for j, c in enumerate(cl):
                    self.net_d[c].load_state_dict(p_ds[j], strict=False)
            else:
                mean = st.group_mean(dicts)
                for c in cl:
                    self.net_d[c].load_state_dict({k: v.clone() for k, v in mean.items()}, strict=False)

    def round(self, real, n_real, z_d, z_g):
        """real [epoch, C, B, d], n_real [epoch, C], z_d / z_g [S, B, 100]. Returns client G losses [S, N]."""
        S, N, B = self.S, self.N, self.B
        t = self.num_communication - self.t          # the reference's counter counts down (while t > 0: ...; t -= 1)
        if self.cloud_epoch:
            if self.algo == "capgan":                # capgan.py:169: t % (self.data_len * cloud_epoch / batch_size) == 0
                due = [bool(t % (self.data_len[s] * self.cloud_epoch / self.B) == 0) for s in range(S)]
                assert all(due) or not any(due), "the reference's Cloud rendezvous would deadlock"
                if all(due):
                    self.cloud()
            elif t % self.cloud_epoch == 0:          # CGLGAN/2DMG/main.py:201, mixed-gan.py:193, CAPGAN/MNIST/capgan.py:169
                self.cloud()
        # MDGAN/MNIST/mdgan.py:158,258: `if t % E == 0`; ACGAN/MNIST/acgan.py:240: `if (num_communication - t) % E == 0`
        if self.E and ((self.t % self.E == 0) if self.algo == "acgan" else (t % self.E == 0)):
            self.share()
        out = torch.zeros(S, N)
        for s in range(S):
            g = self.net_g[s]
            Xd, Xg = st.server_generate(g, z_d[s], z_g[s], N, self.multi_head)
            loss = torch.zeros(N)
            for i in range(N):
                c = s * N + i
                batches = [real[e, c, :int(n_real[e, c])] for e in range(real.shape[0])]
                _, gl = st.worker_train(self.net_d[c], self.opti_d[c], self.loss, self.kind, batches,
                                        [Xd[i].reshape(B, -1)], [Xg[i].reshape(B, -1)], B, self.d_scale)
                loss[i] = gl[0].clone()
            if self.algo == "cglgan":
                self.Lambda[s], self.F_max[s] = st.server_update_cglgan(g, self.opti_g[s], loss, self.beta[s],
                                                                        self.Lambda[s], self.multi_head)
            elif self.algo in ("mdgan", "acgan"):
                self.F_max[s] = st.server_update_mean(g, self.opti_g[s], loss)
            elif self.algo == "capgan":
                self.F_max[s] = st.server_update_capgan(g, self.opti_g[s], loss, self.beta[s], self.Lambda[s], self.opti_L[s])
            elif self.algo == "capgan_copy":
                self.F_max[s] = st.server_update_capgan_copy(g, self.opti_g[s], loss, self.beta[s], self.Lambda[s], self.opti_L[s])
            elif self.algo == "mixed":
                self.F_max[s] = st.server_update_mixed(g, self.opti_g[s], loss, self.beta[s], self.Lambda[s], self.opti_L[s])
            out[s] = loss.detach()
        self.t += 1
        return out


class OracleFL:
    """FLGAN: FLGAN/MNIST/flgan.py:134-163 (Server.run), :211-270 (Worker.run/train)."""

    def __init__(self, num_workers, batch_size, img_shape, lr=0.0002, b1=0.5, b2=0.999):
        d = 1
        for s in img_shape:
            d *= s
        two_d = d == 2
        self.C, self.B = num_workers, batch_size
        mk_g = (lambda: om.Generator2DMD(img_shape)) if two_d else (lambda: om.GeneratorMNIST(img_shape))
        mk_d = (lambda: om.Discriminator2D()) if two_d else (lambda: om.DiscriminatorMNIST1(img_shape))
        self.srv_g, self.srv_d = mk_g(), mk_d()
        self.net_g = [mk_g() for _ in range(self.C)]
        self.net_d = [mk_d() for _ in range(self.C)]
        self.opti_g = [st.make_adam(g.parameters(), lr, b1, b2) for g in self.net_g]
        self.opti_d = [st.make_adam(d_.parameters(), lr, b1, b2) for d_ in self.net_d]
        self.loss = st.make_loss(st.LOSS_BCE)
        self.p_g = st.copy_parameters(self.srv_g)
        self.p_d = st.copy_parameters(self.srv_d)

    def load_global(self):
        for c in range(self.C):
            self.net_d[c].load_state_dict(self.p_d, strict=False)   # flgan.py:222-223
            self.net_g[c].load_state_dict(self.p_g, strict=False)

    def local_minibatch(self, real, n_real, z_d, z_g):
        dl, gl = torch.zeros(self.C), torch.zeros(self.C)
        for c in range(self.C):
            dl[c], gl[c] = st.fl_local_minibatch(self.net_d[c], self.net_g[c], self.loss, self.opti_g[c], self.opti_d[c],
                                                 real[c, :int(n_real[c])], z_d[c], z_g[c], self.B)
        return dl, gl

    def aggregate(self):
        self.p_d = st.fl_aggregate([st.copy_parameters(n) for n in self.net_d], self.C)
        self.p_g = st.fl_aggregate([st.copy_parameters(n) for n in self.net_g], self.C)
        self.load_global()


class OracleFeGAN:
    """FeGAN: Server.run's group loop (fegan.py:125-165) and Worker.run / train (:220-303), clients served serially."""

    def __init__(self, num_workers, batch_size, img_shape, sk, groups, lr=0.0002, b1=0.5, b2=0.999):
        d = 1
        for s in img_shape:
            d *= s
        two_d = d == 2
        self.C, self.B = num_workers, batch_size
        mk_g = (lambda: om.Generator2DMD(img_shape)) if two_d else (lambda: om.GeneratorMNIST(img_shape))
        mk_d = (lambda: om.Discriminator2D()) if two_d else (lambda: om.DiscriminatorMNIST1(img_shape))
        self.srv_g, self.srv_d = mk_g(), mk_d()                       # Server.run: fresh net_g / net_d, :127-131
        self.net_g = [mk_g() for _ in range(self.C)]                  # Worker.run: its own networks, :222-223
        self.net_d = [mk_d() for _ in range(self.C)]
        self.opti_g = [st.make_adam(g.parameters(), lr, b1, b2) for g in self.net_g]
        self.opti_d = [st.make_adam(d_.parameters(), lr, b1, b2) for d_ in self.net_d]
        self.loss = st.make_loss(st.LOSS_BCE)
        self.p_g = st.serialize_model(self.srv_g)                     # :133-134
        self.p_d = st.serialize_model(self.srv_d)
        self.sk = torch.as_tensor(sk, dtype=torch.float32)
        self.groups = [list(g) for g in groups]
        self.t = 0

    def round(self, minibatches):
        """minibatches: list of (real [N, B, d], n_real [N], z_d [N, B, 100], z_g [N, B, 100]) for the round's group."""
        group = self.groups[self.t % len(self.groups)]
        weight = torch.exp(self.sk[group])                            # :142-146
        weight /= weight.sum()
        for c in group:                                               # Worker.run, :228-233
            st.deserialize_model(self.net_g[c], self.p_g)
            st.deserialize_model(self.net_d[c], self.p_d)
        out = []
        for real, n_real, z_d, z_g in minibatches:                    # Worker.train, :282-303
            dl, gl = torch.zeros(len(group)), torch.zeros(len(group))
            for j, c in enumerate(group):
                dl[j], gl[j] = st.fl_local_minibatch(self.net_d[c], self.net_g[c], self.loss, self.opti_g[c], self.opti_d[c],
                                                     real[j, :int(n_real[j])], z_d[j], z_g[j], self.B)
            out.append((dl, gl))
        self.p_g = st.fedavg_aggregate([st.serialize_model(self.net_g[c]) for c in group], weights=weight)   # :163-164
        self.p_d = st.fedavg_aggregate([st.serialize_model(self.net_d[c]) for c in group], weights=weight)
        self.t += 1
        return out
