"""Parity at BASELINE.json's full size (CGLGAN MNIST, 1024 clients / 256 edge servers on one GPU: the bench
workload), through size-independent properties -- the oracle cannot run 1024 clients in seconds:

  * replication: the big simulation is 128 copies of an 8-client / 2-server simulation (same initial modules,
    batches, latents, data shares). Every copy must end BIT-IDENTICAL to copy 0 (a group's result may not depend
    on its index, its CTA, its SM or its neighbours), and copy 0 must match the oracle's 8-client round within
    the bars of test_gpu_rounds (losses 1e-4, parameters 1e-5 bulk).
  * aggregation checksum: the cloud FedAvg over all 256 trunks equals the float64 weighted sum of the rows, and a
    second aggregation with segema = 0 is idempotent.
"""
import copy

import pytest
import torch

from helpers import assert_params_close, quantile_err
from oracle.rounds import OracleMD
from test_gpu_rounds import _compare_generators, _inputs, _self_noise

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("segema", [0.3])
def test_cglgan_mnist_1024_clients_replication_and_oracle(lib, segema):
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    torch.manual_seed(20211212)
    shape, B, d = (1, 28, 28), 100, 784
    W0, S0, R = 8, 2, 128                      # small simulation, replicated R times -> 1024 clients / 256 servers
    N = W0 // S0
    sizes0 = [1000 + 137 * i for i in range(W0)]
    orc = OracleMD("cglgan", W0, S0, B, shape, iid=1, part_sizes=sizes0, segema=segema)
    orc1 = copy.deepcopy(orc)
    k = Knobs(num_workers=W0 * R, num_servers=S0 * R, batch_size=B, epoch=1, segema=segema, iid=1, img_shape=shape)
    sim = MDStyleSim("cglgan", k, part_sizes=sizes0 * R, total_data_len=sum(sizes0) * R)
    sim.load([orc.net_g[s % S0] for s in range(S0 * R)], [orc.net_d[c % W0] for c in range(W0 * R)])

    real, n_real, z_d, z_g = _inputs(W0, S0, B, d, 1, seed=77)
    l_ref = orc.round(real, n_real, z_d, z_g)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        orc1.round(real, n_real, z_d, z_g)
    finally:
        torch.set_num_threads(threads)
    l_gpu = sim.round(real.repeat(1, R, 1, 1).cuda(), n_real.repeat(1, R).cuda(), z_d.repeat(R, 1, 1).cuda(),
                      z_g.repeat(R, 1, 1).cuda())
    torch.cuda.synchronize()

    # ---- replication: bit-identical copies ----
    l = l_gpu.view(R, S0, N)
    assert torch.equal(l, l[0:1].expand_as(l)), "a client's G loss depends on where its group runs"
    rows = sim.bank.rows()                      # [1024, P] discriminator parameters
    P = rows.shape[1]
    rv = rows.view(R, W0, P)
    assert torch.equal(rv, rv[0:1].expand_as(rv)), "discriminator rows of replicated clients differ"
    for bank in (sim.G.trunk, sim.G.heads):
        if bank is None:
            continue
        per = bank.rows // R
        pv = bank.params.detach().view(R, per, -1)
        assert torch.equal(pv, pv[0:1].expand_as(pv)), "generator rows of replicated servers differ"

    # ---- copy 0 against the oracle's 8-client round ----
    assert (l[0].cpu() - l_ref).abs().max() < 1e-4
    for c in range(W0):
        ref = torch.cat([p.detach().reshape(-1) for p in orc.net_d[c].parameters()])
        ref1 = torch.cat([p.detach().reshape(-1) for p in orc1.net_d[c].parameters()])
        assert_params_close(rows[c], ref, steps=1, tag=("D", c), strict=False, bulk=max(1e-5, 3 * _self_noise(ref, ref1)))
    for s in range(S0):
        m = sim.G.make_module()
        sim.G.store_module(s, m)
        _compare_generators(m, orc.net_g[s], 1, ("G", s), bulk=1e-5, ref1=orc1.net_g[s])

    # ---- aggregation checksum over all 256 trunks ----
    trunk = sim.G.trunk
    before = trunk.params.detach().clone()
    want = (sim.A.double().view(-1, 1) * before.double()).sum(0)             # float64 weighted sum
    sim.k.segema = 0.0
    sim.cloud_aggregate()
    torch.cuda.synchronize()
    after = trunk.params.detach()
    assert torch.equal(after, after[0:1].expand_as(after))
    # the engine accumulates in row order in fp32, as Cloud.run's `p += paras * A` loop does: 256 roundings, <= 256 * 2^-24
    assert quantile_err(after[0], want.float(), 1.0) < 256 * 2.0 ** -24
    again = after.clone()
    sim.cloud_aggregate()                       # all rows equal and the weights sum to 1: a fixed point
    torch.cuda.synchronize()
    assert quantile_err(trunk.params.detach()[0], again[0], 1.0) < 256 * 2.0 ** -24
