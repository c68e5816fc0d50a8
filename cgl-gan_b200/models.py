"""Generator / Discriminator classes with the reference's interfaces (constructor arguments,
forward(z) / forward(img), the `.model` trunk and `.paths` head list, identical state_dict keys), built
from the engine's architecture tables so that shapes have one source of truth (csrc/arch.cu).

Reference: model/mnist_model.py:5-88, CGLGAN/2DMG/model.py:26-71, CGLGAN/MNIST/mnist_model.py:30-86,
MDGAN/2DMG/model.py:4-41. They are host-side containers: initialisation (torch default / weights_init),
loading a user's checkpoint, and handing parameters to / from the packed rows. The compute runs in the
CUDA engine (engine.ClientBank, generators.StackedGenerator).
"""
import numpy as np
import torch
import torch.nn as nn

from . import abi

_ACT = {abi.ACT_LRELU: lambda inplace: nn.LeakyReLU(0.2, inplace=inplace), abi.ACT_TANH: lambda _: nn.Tanh(),
        abi.ACT_SIGMOID: lambda _: nn.Sigmoid()}


def sequential_from_desc(desc, inplace_act=False):
    layers = []
    for i in range(desc.n_layers):
        layers.append(nn.Linear(desc.dims[i], desc.dims[i + 1]))
        if desc.bn[i]:
            layers.append(nn.BatchNorm1d(desc.dims[i + 1], desc.bn_eps))
        if desc.act[i] != abi.ACT_NONE:
            layers.append(_ACT[desc.act[i]](inplace_act))
    return nn.Sequential(*layers)


class Discriminator(nn.Module):
    """Discriminator(img_shape[, ns]) -> forward(img) = validity [B,1] (sigmoid) or [B,2] logits.
    arch selects the reference variant; default picks by img_shape like the reference files do."""

    def __init__(self, img_shape=(2,), ns=1, arch=None):
        super().__init__()
        self.img_shape = img_shape
        if arch is None:
            arch = abi.ARCH_D_2D if int(np.prod(img_shape)) == 2 else abi.ARCH_D_MNIST1
        self.arch = arch
        self.desc = abi.arch_describe(arch)
        assert self.desc.dims[0] == int(np.prod(img_shape)), "img_shape does not match the architecture"
        self.model = sequential_from_desc(self.desc)

    def forward(self, img):
        return self.model(img.view(img.shape[0], -1))


class Generator(nn.Module):
    """Single-path generator: Generator(img_shape) (model/mnist_model.py:5-29, MDGAN/2DMG/model.py:4-21)."""

    def __init__(self, img_shape=(2,), arch=None):
        super().__init__()
        self.img_shape = img_shape
        if arch is None:
            arch = abi.ARCH_G_2D_MD if int(np.prod(img_shape)) == 2 else abi.ARCH_G_MNIST
        self.arch = arch
        self.desc = abi.arch_describe(arch)
        self.model = sequential_from_desc(self.desc, inplace_act=True)

    def forward(self, z):
        img = self.model(z)
        return img.view((img.shape[0], *self.img_shape))


class MixGenerator(nn.Module):
    """Shared trunk + one head per client: MixGenerator(img_shape, num_client)
    (model/mnist_model.py:32-66; CGLGAN's `Generator(img_shape, num_client)` is the same class,
    CGLGAN/MNIST/mnist_model.py:30-64, CGLGAN/2DMG/model.py:26-50). forward returns the heads' outputs
    concatenated along dim 0 ([num_client*B, ...])."""

    def __init__(self, img_shape=(2,), num_client=1):
        super().__init__()
        self.img_shape = img_shape
        two_d = int(np.prod(img_shape)) == 2
        self.trunk_desc = abi.arch_describe(abi.ARCH_G_2D_TRUNK if two_d else abi.ARCH_G_MNIST_TRUNK)
        self.head_desc = abi.arch_describe(abi.ARCH_G_2D_HEAD if two_d else abi.ARCH_G_MNIST_HEAD)
        self.model = sequential_from_desc(self.trunk_desc, inplace_act=True)
        self.paths = nn.ModuleList([sequential_from_desc(self.head_desc, inplace_act=True)
                                    for _ in range(num_client)])

    def forward(self, z):
        hidden = self.model(z)
        img = []
        for path in self.paths:
            out = path(hidden)
            img.append(out.view((out.shape[0], *self.img_shape)))
        return torch.cat(img, dim=0)


def weights_init(m):
    """N(0,0.02) Linear/Conv weights, zero bias, BN gamma N(1,0.02): reference mixed-gan.py:68-77."""
    name = m.__class__.__name__
    if name.find('Conv') != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif name.find('BatchNorm') != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)
    elif name.find('Linear') != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
        nn.init.constant_(m.bias.data, 0)
