"""The on-GPU data path and the KL score (8f.4) against the reference's own mechanisms:
torch's DataLoader over the samples, numpy histogram2d + scipy entropy."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from oracle.metrics import kl_score_2d

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,shuffle", [(2, True), (784, True), (784, False), (3, True)])
def test_resident_batches_equal_dataloader_batches(lib, d, shuffle):
    """Same seed -> the gathered [C, B, d] batches are bit-identical to what the reference's per-Worker
    DataLoader(dataset, batch_size=100, shuffle=True) yields, over more than one epoch (ragged last batch, restart)."""
    from cgl_gan_b200.data import ResidentPartitions
    g = torch.Generator().manual_seed(5)
    n, B = 1000, 100
    data = torch.randn(n, d, generator=g)
    perm = torch.randperm(n, generator=g)
    parts = [perm[:241], perm[241:530], perm[530:570], perm[570:]]      # 241, 289, 40 (< B), 430 samples
    torch.manual_seed(123)
    rp = ResidentPartitions(data, parts, B, shuffle=shuffle)
    got = []
    for _ in range(6):
        real, n_real = rp.next_batches()
        got.append((real.cpu().clone(), n_real.cpu().clone()))
    torch.manual_seed(123)
    loaders = [DataLoader(dataset=data[p], batch_size=B, shuffle=shuffle) for p in parts]   # Worker.__init__
    iters = [iter(dl) for dl in loaders]
    for r in range(6):
        for c in range(len(parts)):
            try:
                want = next(iters[c])
            except StopIteration:                                                            # Worker.train
                loaders[c] = DataLoader(dataset=data[parts[c]], batch_size=B, shuffle=shuffle)
                iters[c] = iter(loaders[c])
                want = next(iters[c])
            k = want.shape[0]
            assert int(got[r][1][c]) == k
            assert torch.equal(got[r][0][c, :k], want)
            assert not got[r][0][c, k:].any()                                                # zero padding
    assert rp.h2d_bytes_per_round == 4 * B * 8 + 16


def test_gather_rows_edge_cases(lib):
    import ctypes as C
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    data = torch.arange(50, dtype=torch.float32, device="cuda").view(10, 5)
    idx = torch.tensor([9, -1, 0, 10, 3], dtype=torch.int64, device="cuda")                   # -1 and out of range -> zeros
    out = torch.full((5, 5), 7.0, device="cuda")
    lib.check(lib.lib.cgl_gather_rows(5, 5, lib.ptr(data), 10, lib.ptr(idx), lib.ptr(out), st))
    want = torch.stack([data[9], torch.zeros(5, device="cuda"), data[0], torch.zeros(5, device="cuda"), data[3]])
    assert torch.equal(out, want)
    lib.check(lib.lib.cgl_gather_rows(0, 5, None, 10, None, None, st))                        # empty: no launch


def test_kl_score_matches_numpy_scipy(lib):
    from cgl_gan_b200.data import KLScore2D
    g = torch.Generator().manual_seed(3)
    ang = torch.rand(4000, generator=g) * 6.2831853
    real = torch.stack([ang.cos(), ang.sin()], 1) * 0.9 + 0.01 * torch.randn(4000, 2, generator=g)
    gen = torch.tanh(torch.randn(3000, 2, generator=g) * 0.8)
    gen[:7] = torch.tensor([[1.0, 1.0], [-1.0, -1.0], [1.0, -1.0], [0.125, 0.25], [1.0000001, 0.0], [-1.5, 0.2], [0.0, 0.0]])
    score = KLScore2D(real)
    got = score(gen.cuda()).item()
    want = kl_score_2d(real.numpy(), gen.numpy())
    assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (got, want)
    h = torch.zeros(256, dtype=torch.int32, device="cuda")
    import ctypes as C
    lib.check(lib.lib.cgl_hist2d(gen.shape[0], lib.ptr(gen.cuda().contiguous()), 2, lib.ptr(h),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    cnt, _, _ = np.histogram2d(gen[:, 0].numpy(), gen[:, 1].numpy(), bins=16, range=[[-1, 1], [-1, 1]])
    assert np.array_equal(h.cpu().numpy().reshape(16, 16), cnt.astype(np.int32))              # bit-exact bin counts
