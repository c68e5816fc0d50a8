"""Eager round() against round_graph() (CUDA-graph replay) of the CGLGAN simulation, for launch-bound topologies.
    python profiles/graph_bench.py [dataset: 2dmg|mnist] [workers] [servers] [rounds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi, models  # noqa: E402
from cgl_gan_b200.sim import Knobs, MDStyleSim  # noqa: E402

ds = sys.argv[1] if len(sys.argv) > 1 else "2dmg"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 10
S = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 50
abi.require_device()
shape = (2,) if ds == "2dmg" else (1, 28, 28)
d = 2 if ds == "2dmg" else 784
B = 100
torch.manual_seed(0)
k = Knobs(num_workers=W, num_servers=S, batch_size=B, epoch=1, cloud_epoch=1, iid=1, img_shape=shape)


def make():
    sim = MDStyleSim("cglgan", k, part_sizes=[1000] * W)
    sim.load([sim.G.make_module() for _ in range(S)], [models.Discriminator(shape) for _ in range(W)])
    return sim


real = torch.tanh(torch.randn(W, B, d, device="cuda"))
for name in ("eager", "graph"):
    sim = make()
    step = sim.round if name == "eager" else sim.round_graph
    for _ in range(5):
        step(real)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rounds):
        step(real)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / rounds
    print(f"{ds} {W} workers / {S} servers, {name}: {ms:.3f} ms per round, {W / ms * 1e3:.0f} client-steps/s")
