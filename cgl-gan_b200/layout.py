"""Packed parameter rows <-> reference modules / state_dicts.

A packed row is the flat fp32 vector of a module's parameters in `parameters()` order -- the order of
fedlab's SerializationTool.serialize_model (reference capgan.py:170, fegan.py:133-134) -- padded to a
multiple of 32 floats so that every client's row starts 128-byte aligned in HBM. BatchNorm running
statistics (buffers, not parameters) live in a second, parallel row.
"""
import torch

from . import abi

ROW_ALIGN = 32  # floats


def padded(n, align=ROW_ALIGN):
    return (n + align - 1) // align * align


class RowLayout:
    """Offsets of one MLP stack (an abi.MlpDesc) inside a packed row."""

    def __init__(self, desc):
        self.desc = desc
        lay = abi.layout_of(desc)
        self.n_layers = desc.n_layers
        self.dims = [desc.dims[i] for i in range(desc.n_layers + 1)]
        self.act = [desc.act[i] for i in range(desc.n_layers)]
        self.bn = [bool(desc.bn[i]) for i in range(desc.n_layers)]
        self.n_params = int(lay.n_params)
        self.n_stats = int(lay.n_bn_stats)
        self.w_off = [int(lay.w_off[i]) for i in range(self.n_layers)]
        self.b_off = [int(lay.b_off[i]) for i in range(self.n_layers)]
        self.bn_w_off = [int(lay.bn_w_off[i]) for i in range(self.n_layers)]
        self.bn_b_off = [int(lay.bn_b_off[i]) for i in range(self.n_layers)]
        self.bn_mean_off = [int(lay.bn_mean_off[i]) for i in range(self.n_layers)]
        self.bn_var_off = [int(lay.bn_var_off[i]) for i in range(self.n_layers)]
        self.ld = padded(self.n_params)
        self.ld_stats = padded(max(self.n_stats, 1))


def flatten_params(module):
    """serialize_model: cat of parameters() (reference call site capgan.py:170)."""
    return torch.cat([p.detach().reshape(-1) for p in module.parameters()])


def load_flat_params(module, flat):
    """deserialize_model (reference call site capgan.py:175)."""
    i = 0
    with torch.no_grad():
        for p in module.parameters():
            n = p.numel()
            p.copy_(flat[i:i + n].view_as(p))
            i += n
    assert i == flat.numel(), (i, flat.numel())


def flatten_bn_stats(module):
    """running_mean, running_var of every BatchNorm in module order (num_batches_tracked is dropped,
    like copy_parameters' len(var.size()) != 0 test, CGLGAN/2DMG/main.py:167)."""
    out = []
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            out += [m.running_mean.reshape(-1), m.running_var.reshape(-1)]
    return torch.cat(out) if out else torch.zeros(0)


def load_bn_stats(module, flat):
    i = 0
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                n = m.running_mean.numel()
                m.running_mean.copy_(flat[i:i + n]); i += n
                m.running_var.copy_(flat[i:i + n]); i += n
