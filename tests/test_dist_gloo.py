"""N > 1 host logic on CPU: two gloo ranks shard the servers, pre-weight their partial sums and all-reduce;
the result must equal the single-process Cloud aggregation of the oracle. (The GPU path runs the same plan
with cgl_mix_allreduce over NCCL.)"""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cgl_gan_b200.dist import global_weights, shard_plan
    from oracle import models as om
    from oracle import steps as st
    S, W = 5, 10
    plan = shard_plan(W, S, world)
    lo, hi, clo, chi = plan[rank]
    torch.manual_seed(0)                      # every rank builds the same S generators, keeps its shard
    nets = [om.Generator2DCGL((2,), 2) for _ in range(S)]
    data_len = torch.tensor([1000., 3000., 500., 700., 1200.])
    A_local = global_weights(data_len[lo:hi])
    # local pre-weighted partial sum (what cgl_wsum computes on each GPU), then the all-reduce
    part = None
    for j, s in enumerate(range(lo, hi)):
        v = st.serialize_model(nets[s].model) * A_local[j]
        part = v if part is None else part + v
    dist.all_reduce(part)
    A = data_len / data_len.sum()
    p = st.cloud_aggregate([st.copy_parameters(n.model) for n in nets], A)
    ref = torch.cat([p[k].reshape(-1) for k, _ in nets[0].model.named_parameters()])
    q.put((rank, plan, float((part - ref).abs().max() / ref.abs().max()), A_local.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_cloud_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plan = res[0][1]
    assert plan == [(0, 3, 0, 6), (3, 5, 6, 10)]          # servers never straddle ranks; clients follow
    for rank, _, err, A_local in res:
        assert err < 1e-6, (rank, err)
    assert abs(sum(res[0][3]) + sum(res[1][3]) - 1.0) < 1e-6


def test_shard_plan_covers_everything():
    from cgl_gan_b200.dist import shard_plan
    for world in (1, 2, 4, 8):
        plan = shard_plan(1024, 256, world)
        assert plan[0][0] == 0 and plan[-1][1] == 256 and plan[-1][3] == 1024
        for a, b in zip(plan, plan[1:]):
            assert a[1] == b[0] and a[3] == b[2]
