"""The C-ABI library loads on a CPU-only host and exports every symbol include/cgl_b200.h declares;
argument validation works without a GPU; compute calls fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cgl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgl_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib.lib, n), f"{n} declared in include/cgl_b200.h but not exported"


def test_arch_tables_match_reference_parameter_counts(lib):
    """parameter counts measured on the reference's classes."""
    expect = {lib.ARCH_D_2D: 33665, lib.ARCH_D_MNIST1: 533505, lib.ARCH_D_MNIST2: 533762,
              lib.ARCH_G_2D_MD: 59010, lib.ARCH_G_MNIST: 1510032, lib.ARCH_G_2D_TRUNK: 3232,
              lib.ARCH_G_2D_HEAD: 66, lib.ARCH_G_MNIST_TRUNK: 179072, lib.ARCH_G_MNIST_HEAD: 1330960}
    for arch, n in expect.items():
        assert lib.layout_of(lib.arch_describe(arch)).n_params == n
    assert lib.layout_of(lib.arch_describe(lib.ARCH_G_MNIST)).n_bn_stats == 3584
    with pytest.raises(lib.CglError):
        lib.arch_describe(99)


def test_argument_validation_without_gpu(lib):
    d = lib.arch_describe(lib.ARCH_D_MNIST1)
    cfg = lib.TrainCfg(lib.LOSS_CE, 1.0, 2e-4, 0.5, 0.999, 1e-8)     # CE on a 1-logit sigmoid D
    rc = lib.lib.cgl_d_step(C.byref(d), 1, None, None, None, 0, None, None, None, None, None, None, 100,
                            C.byref(cfg), None, None, 0, None)
    assert rc == -1 and b"CrossEntropy" in lib.lib.cgl_last_error()
    assert lib.lib.cgl_d_step_workspace_bytes(C.byref(d), 1024, 100) > 1024 * 200 * (512 + 256) * 4
    rc = lib.lib.cgl_mix_csr(2, 8, None, None, None, None, 8, None, 8, None)
    assert rc == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert not lib.device_ok()
    from cgl_gan_b200.engine import ClientBank
    with pytest.raises(lib.CglError):
        ClientBank(lib.ARCH_D_2D, 2, 100)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cgl-gan_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
