"""The TMA-fed tcgen05 kernels (csrc/tc_tma.cuh) through the C ABI: every layer shape of the networks against float64 and
against the exact-fp32 FFMA kernel, the ragged tiles that use the tail tensor maps (M = 784, the data gradient's K = 784),
concatenated real|fake batches (one tensor map per source), index-selected groups, and -- the regression test of the
round-2 finding that TMA touches the addresses of out-of-bounds box rows -- operands placed at the very END of their own
cudaMalloc allocation (reference products: Linear forward / backward of model/mnist_model.py:5-29,71-88)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

ACT = 0.2


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _fwd(lib, G, rows, K, out, x_ptr, prm_ptr, ldp, ids=None):
    y = torch.empty(G, rows, out, device="cuda")
    lib.check(lib.lib.cgl_linear_fwd(G, rows, K, out, x_ptr, rows * K, prm_ptr, ldp, lib.ptr(ids), 0, K * out, lib.ACT_LRELU, ACT,
                                     lib.ptr(y), rows * out, _st()))
    return y


def _bwd(lib, G, rows, K, out, dy_ptr, prm_ptr, ldp, saved):
    dx = torch.empty(G, rows, K, device="cuda")
    lib.check(lib.lib.cgl_linear_bwd_data(G, rows, K, out, dy_ptr, rows * out, prm_ptr, ldp, None, 0, lib.ptr(saved), rows * K,
                                          lib.ACT_LRELU, ACT, lib.ptr(dx), rows * K, _st()))
    return dx


def _refs(prm, x, dy, K, out):
    G = prm.shape[0]
    W = prm[:, :K * out].view(G, out, K).double()
    b = prm[:, K * out:K * out + out].double()
    y = torch.nn.functional.leaky_relu(torch.baddbmm(b.unsqueeze(1), x.double(), W.transpose(1, 2)), ACT)
    dx = torch.bmm(dy.double(), W) * torch.where(x > 0, 1.0, ACT)
    return y, dx


SHAPES = [(1024, 784), (784, 512), (512, 1024), (512, 256), (528, 512), (800, 128), (256, 512)]


@pytest.mark.parametrize("K,out", SHAPES, ids=[f"{k}x{o}" for k, o in SHAPES])
@pytest.mark.parametrize("rows", [100, 200, 37])
def test_tma_products_match_float64_and_ffma(lib, K, out, rows):
    torch.manual_seed(K + out + rows)
    G = 3
    ldp = (K * out + out + 31) // 32 * 32
    prm = (torch.randn(G, ldp) * 0.1).cuda()
    x, dy = torch.randn(G, rows, K).cuda(), torch.randn(G, rows, out).cuda()
    y_ref, dx_ref = _refs(prm, x, dy, K, out)
    got = {}
    for mode in (0, 1):                                   # automatic (TMA-fed tcgen05 for these shapes) / FFMA only
        lib.check(lib.lib.cgl_set_gemm_mode(mode))
        got[mode] = (_fwd(lib, G, rows, K, out, lib.ptr(x), lib.ptr(prm), ldp), _bwd(lib, G, rows, K, out, lib.ptr(dy), lib.ptr(prm), ldp, x))
    lib.check(lib.lib.cgl_set_gemm_mode(0))
    torch.cuda.synchronize()
    for mode in (0, 1):
        e_f = ((got[mode][0].double() - y_ref).abs().max() / y_ref.abs().max()).item()
        e_b = ((got[mode][1].double() - dx_ref).abs().max() / dx_ref.abs().max()).item()
        assert e_f < 2e-6 and e_b < 2e-6, (mode, e_f, e_b)          # fp32-grade: the 3xTF32 split drops 2^-22 terms
    # the two kernels agree with each other within the sum of their distances from float64
    assert ((got[0][0] - got[1][0]).abs().max() / y_ref.abs().max()).item() < 4e-6


def test_tma_index_selected_groups(lib):
    """ids: the groups of the call are rows of a larger bank (FeGAN's group of the round) -- the TMA group coordinate."""
    torch.manual_seed(5)
    K, out, rows, bank = 512, 256, 100, 7
    ldp = (K * out + out + 31) // 32 * 32
    prm = (torch.randn(bank, ldp) * 0.1).cuda()
    ids = torch.tensor([5, 0, 3], dtype=torch.int32, device="cuda")
    x = torch.randn(3, rows, K).cuda()
    y = _fwd(lib, 3, rows, K, out, lib.ptr(x), lib.ptr(prm), ldp, ids)
    y_ref, _ = _refs(prm[ids.long()], x, torch.zeros(3, rows, out, device="cuda"), K, out)
    assert ((y.double() - y_ref).abs().max() / y_ref.abs().max()).item() < 2e-6


@pytest.mark.parametrize("kind,mode,K,out,rows", [
    ("fwd", "x", 1024, 784, 100), ("fwd", "p", 1024, 784, 100), ("fwd", "x", 784, 512, 200), ("fwd", "p", 784, 512, 100),
    ("fwd", "x", 512, 256, 37), ("bwd", "x", 512, 1024, 100), ("bwd", "p", 784, 512, 100), ("bwd", "p", 1024, 784, 100),
    ("bwd", "x", 1024, 784, 100), ("fwd", "x", 100, 128, 100), ("fwd", "p", 100, 128, 100)])
def test_no_access_behind_an_operand_at_the_end_of_its_allocation(lib, kind, mode, K, out, rows):
    """The batch operand (x) or the parameter bank (p) ends exactly where its own cudaMalloc allocation ends. A TMA box
    whose out-of-bounds rows lie behind the allocation faults (measured), so the kernels use tail maps / fall back."""
    rt = C.CDLL("libcudart.so.12")
    torch.manual_seed(11)
    G = 2
    ldp = (K * out + out + 31) // 32 * 32
    prm = (torch.randn(G, ldp) * 0.1).cuda()
    x, dy = torch.randn(G, rows, K).cuda(), torch.randn(G, rows, out).cuda()
    src = {"x": x if kind == "fwd" else dy, "p": prm}[mode]
    nbytes = src.numel() * 4
    size = (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20) + (2 << 20)
    p = C.c_void_p()
    assert rt.cudaMalloc(C.byref(p), C.c_size_t(size)) == 0
    try:
        at = p.value + size - nbytes
        assert rt.cudaMemcpy(C.c_void_p(at), C.c_void_p(src.data_ptr()), C.c_size_t(nbytes), 3) == 0
        xp = C.c_void_p(at) if (mode == "x" and kind == "fwd") else lib.ptr(x)
        dyp = C.c_void_p(at) if (mode == "x" and kind == "bwd") else lib.ptr(dy)
        pp = C.c_void_p(at) if mode == "p" else lib.ptr(prm)
        y_ref, dx_ref = _refs(prm, x, dy, K, out)
        if kind == "fwd":
            got, ref = _fwd(lib, G, rows, K, out, xp, pp, ldp), y_ref
        else:
            got, ref = _bwd(lib, G, rows, K, out, dyp, pp, ldp, x), dx_ref
        assert rt.cudaDeviceSynchronize() == 0, "a kernel touched memory behind the operand"
        assert ((got.double() - ref).abs().max() / ref.abs().max()).item() < 2e-6
    finally:
        rt.cudaFree(p)


# The opt-in variants of the TMA-fed kernel (CGL_TUNE is read once per process, so each runs in its own interpreter through
# profiles/pair_check.py: every layer shape of a round against float64). 2: look-ahead barrier tests; 4: two issuing warps with a
# token barrier (with 2 and, | 2097152, with 4 A stages); 262144: pipelined A converters; 33554432: truncated-hi batch operand.
@pytest.mark.parametrize("extra", [2, 4, 4 | 2097152, 262144, 33554432])
def test_opt_in_variants_match_float64(extra):
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = 1 | 8 | 32 | 64 | 256 | 512 | 1024 | 131072           # TC_TUNE_DEFAULT (csrc/tc_gemm.cuh)
    env = dict(os.environ, CGL_TUNE=str(base | extra))
    res = subprocess.run([sys.executable, os.path.join(root, "profiles", "pair_check.py")], capture_output=True, text=True,
                         timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    worst = float(re.search(r"worst error ([0-9.e+-]+)", res.stdout).group(1))
    assert worst < 2e-6, res.stdout        # the default kernel: 1.0e-6 at K = 1024 (profiles/tma_shapes.py)
