"""CPU restatement of the reference's step bodies (torch autograd + torch.optim), fp32.

The reference runs every actor as a thread and exchanges tensors through queues; the RNG order is a
thread race. The oracle therefore takes every random input explicitly (initial
parameters, z, real batches) and visits servers 0..S-1 and clients 0..W-1 in index order.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import copy

import torch
import torch.nn.functional as F
from torch import nn, optim

LOSS_BCE, LOSS_CE, LOSS_MSE = 0, 1, 2


def make_loss(kind):
    """nn.BCELoss (CGLGAN/2DMG/main.py:336), nn.CrossEntropyLoss (capgan.py:312), nn.MSELoss (LSGAN)."""
    return {LOSS_BCE: nn.BCELoss, LOSS_CE: nn.CrossEntropyLoss, LOSS_MSE: nn.MSELoss}[kind]()


def _targets(kind, n, value):
    # BCE/MSE: Tensor(n,1).fill_(v) (CGLGAN/2DMG/main.py:357,362); CE: LongTensor(n).fill_(v) (capgan.py:329,334)
    if kind == LOSS_CE:
        return torch.full((n,), int(value), dtype=torch.long)
    return torch.full((n, 1), float(value), dtype=torch.float32)


def make_adam(params, lr=0.0002, b1=0.5, b2=0.999):
    """optim.Adam(net.parameters(), lr=lr, betas=(b1, b2)): CGLGAN/2DMG/main.py:192,337"""
    return optim.Adam(params, lr=lr, betas=(b1, b2))


def worker_d_step(net_d, opti_d, loss, kind, real_imgs, X, batch_size, d_loss_scale=1.0):
    """One discriminator update.
    BCE: CGLGAN/2DMG/main.py:357-366, MDGAN/MNIST/mdgan.py:278-288.
    CE : capgan.py:329-341 (scale 0.5), ACGAN/MNIST/acgan.py:281-292 (scale 1)."""
    valid = _targets(kind, real_imgs.shape[0], 1)
    opti_d.zero_grad()
    real_loss = loss(net_d(real_imgs), valid)
    fake = _targets(kind, batch_size, 0)
    fake_loss = loss(net_d(X), fake)
    D_loss = (real_loss + fake_loss) if d_loss_scale == 1.0 else (real_loss + fake_loss) * d_loss_scale
    D_loss.backward()
    opti_d.step()
    return D_loss.detach()


def worker_g_loss(net_d, loss, kind, Xg, batch_size):
    """G_loss = loss(net_d(Xg), valid), graph attached: CGLGAN/2DMG/main.py:368-372, capgan.py:343-346."""
    valid = _targets(kind, batch_size, 1)
    return loss(net_d(Xg), valid)


def worker_train(net_d, opti_d, loss, kind, real_batches, Xs, Xgs, batch_size, d_loss_scale=1.0):
    """Worker.train with `epoch` = len(real_batches) and one entry of Xs / Xgs per serving edge server.
    CGLGAN/2DMG/main.py:344-375: for each epoch, for each (server) X: a D step; then one G loss per server."""
    d_losses = []
    for real in real_batches:
        for X in Xs:
            d_losses.append(worker_d_step(net_d, opti_d, loss, kind, real, X, batch_size, d_loss_scale))
    g_losses = [worker_g_loss(net_d, loss, kind, Xg, batch_size) for Xg in Xgs]
    return d_losses, g_losses


# ------------------------------------------------------------------------------------------------
# Server.train variants. `client_losses(Xd_list, Xg_list) -> [N] graph-attached losses`
# stands for the queue round trip to the N clients of this server.
# ------------------------------------------------------------------------------------------------

def server_generate(net_g, z_d, z_g, n_clients, multi_head):
    """CGLGAN/2DMG/main.py:229-234: Xd under no_grad, Xg with graph; both in train mode (BN batch stats,
    running stats updated on both passes)."""
    with torch.no_grad():
        out = net_g(z_d)
        Xd = list(torch.chunk(out, n_clients, dim=0)) if multi_head else [out] * n_clients
    out = net_g(z_g)
    Xg = list(torch.chunk(out, n_clients, dim=0)) if multi_head else [out] * n_clients
    return Xd, Xg


def server_update_cglgan(net_g, opti_g, loss, beta, Lambda, multi_head):
    """CGLGAN/2DMG/main.py:245-276 == CGLGAN/MNIST/main.py:263-293. `loss` is the [N] tensor of client
    G losses (graph attached). Returns (new Lambda, F_max)."""
    opti_g.zero_grad()
    if multi_head:
        losses = loss.sum()
        net_g.model.requires_grad_(False)
        losses.backward(retain_graph=True)
        net_g.model.requires_grad_(True)
    gamma = F.softmax(Lambda * loss, dim=0).detach()
    F_beta = (beta * loss).sum()
    F_gamma = (gamma * loss).sum()
    F_max = (F_beta + F_gamma) / 2
    if multi_head:
        net_g.paths.requires_grad_(False)
        F_max.backward()
        net_g.paths.requires_grad_(True)
    else:
        F_max.backward()
    grad = (loss * loss * gamma).sum() - (loss * gamma * F_gamma).sum()
    new_lambda = (Lambda + 10 * grad).detach()
    opti_g.step()
    return new_lambda, F_max.detach()


def server_update_mean(net_g, opti_g, loss):
    """MDGAN/MNIST/mdgan.py:196-205, ACGAN/MNIST/acgan.py:166-175: losses = loss.mean(); backward; step."""
    opti_g.zero_grad()
    losses = loss.mean()
    losses.backward()
    opti_g.step()
    return losses.detach()


def server_update_capgan(net_g, opti_g, loss, beta, Lambda, opti_L):
    """capgan.py:229-260: alpha = softmax(softmax(Lambda*loss) * beta); F_max = sum(alpha*loss) - 0.001*Lambda;
    opti_L (SGD lr 0.1 on Lambda) steps too."""
    opti_g.zero_grad()
    opti_L.zero_grad()
    alpha = F.softmax(Lambda.detach() * loss.detach(), dim=0)
    alpha = F.softmax(alpha * beta, dim=0)
    F_max = (alpha * loss).sum() - 0.001 * Lambda
    F_max.backward()
    opti_L.step()
    opti_g.step()
    return F_max.detach()


def server_update_capgan_copy(net_g, opti_g, loss, beta, Lambda, opti_L):
    """CAPGAN/MNIST/capgan.py:241-243 (the packaged copy): gamma = softmax(Lambda*loss);
    s = softmax(beta*gamma); F_max = sum(s*loss) - 0.001*Lambda."""
    opti_g.zero_grad()
    opti_L.zero_grad()
    gamma = F.softmax(Lambda.detach() * loss.detach(), dim=0)
    s = F.softmax(beta * gamma, dim=0)
    F_max = (s * loss).sum() - 0.001 * Lambda
    F_max.backward()
    opti_L.step()
    opti_g.step()
    return F_max.detach()


def server_update_mixed(net_g, opti_g, loss, beta, Lambda, opti_L):
    """mixed-gan.py:254-288: heads get d(sum loss), trunk gets d(F_max) with
    alpha = softmax(beta * Lambda * loss); opti_L then opti_g step."""
    opti_g.zero_grad()
    losses = loss.sum()
    net_g.model.requires_grad_(False)
    losses.backward(retain_graph=True)
    net_g.model.requires_grad_(True)
    opti_L.zero_grad()
    alpha = F.softmax(beta * Lambda.detach() * loss.detach(), dim=0)
    F_max = (alpha * loss).sum() - 0.001 * Lambda
    net_g.paths.requires_grad_(False)
    F_max.backward()
    net_g.paths.requires_grad_(True)
    opti_L.step()
    opti_g.step()
    return F_max.detach()


# ------------------------------------------------------------------------------------------------
# FL-style local step (a5)
# ------------------------------------------------------------------------------------------------

def fl_local_minibatch(net_d, net_g, loss, opti_g, opti_d, real_imgs, z_d, z_g, batch_size, kind=LOSS_BCE):
    """One minibatch of FLGAN Worker.train: FLGAN/MNIST/flgan.py:251-269 == FLGAN/2DMG/flgan.py:239-256.
    Xd is NOT detached in the reference; the G grads it leaves behind are zeroed at opti_g.zero_grad()."""
    fake = _targets(kind, batch_size, 0)
    Xd = net_g(z_d)
    valid = _targets(kind, real_imgs.shape[0], 1)
    opti_d.zero_grad()
    real_loss = loss(net_d(real_imgs), valid)
    fake_loss = loss(net_d(Xd), fake)
    D_loss = real_loss + fake_loss
    D_loss.backward()
    opti_d.step()

    valid = _targets(kind, batch_size, 1)
    opti_g.zero_grad()
    Xg = net_g(z_g)
    g_loss = loss(net_d(Xg), valid)
    g_loss.backward()
    opti_g.step()
    return D_loss.detach(), g_loss.detach()


# ------------------------------------------------------------------------------------------------
# Aggregation (a8, a9, a10)
# ------------------------------------------------------------------------------------------------

def copy_parameters(net):
    """CGLGAN/2DMG/main.py:164-169: every state_dict entry with at least one dimension
    (drops BatchNorm's 0-dim num_batches_tracked, keeps running_mean / running_var)."""
    return {k: v.clone() for k, v in net.state_dict().items() if len(v.size()) != 0}


def cloud_aggregate(dicts, A):
    """Cloud.run, CGLGAN/2DMG/main.py:126-133: p[key] (+)= paras[key] * A[idx], servers in index order."""
    p = {}
    for idx, paras in enumerate(dicts):
        for key in paras:
            if key in p:
                p[key] += paras[key] * A[idx]
            else:
                p[key] = paras[key] * A[idx]
    return p


def segema_mix(self_p, recv_p, segema):
    """CGLGAN/2DMG/main.py:206-207: recv_p[key] = segema * self_p[key] + (1 - segema) * recv_p[key]"""
    return {k: segema * self_p[k] + (1 - segema) * recv_p[k] for k in recv_p}


def fl_aggregate(dicts, n_clients):
    """FL Server.run, FLGAN/MNIST/flgan.py:148-162: p[key] (+)= paras[key] / len(client_list)."""
    p = {}
    for paras in dicts:
        for key in paras:
            if key in p:
                p[key] += paras[key] / n_clients
            else:
                p[key] = paras[key] / n_clients
    return p


def serialize_model(net):
    """fedlab SerializationTool.serialize_model (fedlab <= 1.2, un-vendored):
    cat([p.data.view(-1) for p in model.parameters()]). Call sites capgan.py:170, fegan.py:133-134."""
    return torch.cat([p.data.view(-1) for p in net.parameters()]).clone()


def deserialize_model(net, flat):
    """fedlab SerializationTool.deserialize_model(mode='copy'). Call sites capgan.py:175, fegan.py:232-233."""
    i = 0
    for p in net.parameters():
        n = p.numel()
        p.data.copy_(flat[i:i + n].view_as(p))
        i += n


def fedavg_aggregate(flats, weights=None):
    """fedlab Aggregators.fedavg_aggregate: weights /= sum(weights); sum(stack(list, -1) * weights, -1).
    Call sites capgan.py:114, fegan.py:163-164."""
    if weights is None:
        weights = torch.ones(len(flats))
    weights = torch.as_tensor(weights, dtype=torch.float32)
    weights = weights / torch.sum(weights)
    return torch.sum(torch.stack(flats, dim=-1) * weights, dim=-1)


def mdgan_swap(p_ds, rd):
    """The commented-out discriminator swap, MDGAN/MNIST/mdgan.py:158-164,258-262: the server collects
    all D dicts in client order, self.rd.shuffle(p_ds) (Random(rank+100)), worker idx loads p_ds[idx]."""
    p_ds = list(p_ds)
    rd.shuffle(p_ds)
    return p_ds


def group_mean(dicts):
    """ACGAN delta-gossip fixed point (ACGAN/MNIST/acgan.py:240-263) == dead receive_parameter mean
    (CGLGAN/2DMG/main.py:171-179): every member of the group ends with the uniform mean of the group."""
    p = copy.deepcopy(dicts[0])
    for d in dicts[1:]:
        for key in p:
            p[key] += d[key]
    for key in p:
        p[key] /= len(dicts)
    return p
