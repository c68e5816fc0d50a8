"""Restated model classes of the reference (same constructor signatures, attribute names and
state_dict keys), torch CPU.  a6 in SURVEY.md section 8.

  Discriminator2D      CGLGAN/2DMG/model.py:54-71, MDGAN/2DMG/model.py:26-41
  DiscriminatorMNIST1  CGLGAN/MNIST/mnist_model.py:69-86, MDGAN/MNIST/mnist_model.py:34-50
  DiscriminatorMNIST2  model/mnist_model.py:71-88
  Generator2DMD        MDGAN/2DMG/model.py:4-21
  Generator2DCGL       CGLGAN/2DMG/model.py:26-50
  GeneratorMNIST       model/mnist_model.py:5-29
  MixGeneratorMNIST    model/mnist_model.py:32-66 == CGLGAN/MNIST/mnist_model.py:30-64
"""
import numpy as np
import torch
import torch.nn as nn


def _block(in_feat, out_feat, normalize=True):
    # model/mnist_model.py:10-15 -- note BatchNorm1d(out_feat, 0.8): the 0.8 is eps
    layers = [nn.Linear(in_feat, out_feat)]
    if normalize:
        layers.append(nn.BatchNorm1d(out_feat, 0.8))
    layers.append(nn.LeakyReLU(0.2, inplace=True))
    return layers


class Discriminator2D(nn.Module):
    def __init__(self, ns=1):
        super().__init__()
        self.model = nn.Sequential(
            nn.Linear(2, 128), nn.LeakyReLU(0.2),
            nn.Linear(128, 256), nn.LeakyReLU(0.2),
            nn.Linear(256, 1), nn.Sigmoid())

    def forward(self, img):
        return self.model(img.view(img.shape[0], -1))


class DiscriminatorMNIST1(nn.Module):
    def __init__(self, img_shape, ns=1):
        super().__init__()
        self.img_shape = img_shape
        self.model = nn.Sequential(
            nn.Linear(int(np.prod(img_shape)), 512), nn.LeakyReLU(0.2),
            nn.Linear(512, 256), nn.LeakyReLU(0.2),
            nn.Linear(256, 1), nn.Sigmoid())

    def forward(self, img):
        return self.model(img.view(img.shape[0], -1))


class DiscriminatorMNIST2(nn.Module):
    def __init__(self, img_shape):
        super().__init__()
        self.img_shape = img_shape
        self.model = nn.Sequential(
            nn.Linear(int(np.prod(img_shape)), 512), nn.LeakyReLU(0.2),
            nn.Linear(512, 256), nn.LeakyReLU(0.2),
            nn.Linear(256, 2))

    def forward(self, img):
        return self.model(img.view(img.shape[0], -1))


class DiscriminatorMNISTLS(nn.Module):
    """LSGAN variant: linear validity output (model/lsgan.py:96 `adv_layer` is a bare Linear)."""

    def __init__(self, img_shape):
        super().__init__()
        self.img_shape = img_shape
        self.model = nn.Sequential(
            nn.Linear(int(np.prod(img_shape)), 512), nn.LeakyReLU(0.2),
            nn.Linear(512, 256), nn.LeakyReLU(0.2),
            nn.Linear(256, 1))

    def forward(self, img):
        return self.model(img.view(img.shape[0], -1))


class Generator2DMD(nn.Module):
    def __init__(self, img_shape):
        super().__init__()
        self.img_shape = img_shape
        self.model = nn.Sequential(
            nn.Linear(100, 256), nn.LeakyReLU(0.2),
            nn.Linear(256, 128), nn.LeakyReLU(0.2),
            nn.Linear(128, 2), nn.Tanh())

    def forward(self, z):
        img = self.model(z)
        return img.view((img.shape[0], *self.img_shape))


class Generator2DCGL(nn.Module):
    def __init__(self, img_shape, num_client):
        super().__init__()
        self.img_shape = img_shape
        self.model = nn.Sequential(nn.Linear(100, 32), nn.LeakyReLU(0.2))
        self.paths = nn.ModuleList(
            [nn.Sequential(nn.Linear(32, 2), nn.Tanh()) for _ in range(num_client)])

    def forward(self, z):
        hidden = self.model(z)
        return torch.cat([path(hidden) for path in self.paths], dim=0)


class GeneratorMNIST(nn.Module):
    def __init__(self, img_shape):
        super().__init__()
        self.img_shape = img_shape
        self.model = nn.Sequential(
            *_block(100, 128, normalize=False), *_block(128, 256), *_block(256, 512),
            *_block(512, 1024), nn.Linear(1024, int(np.prod(img_shape))), nn.Tanh())

    def forward(self, z):
        img = self.model(z)
        return img.view((img.shape[0], *self.img_shape))


class MixGeneratorMNIST(nn.Module):
    def __init__(self, img_shape, num_client):
        super().__init__()
        self.img_shape = img_shape
        self.model = nn.Sequential(
            *_block(100, 128, normalize=False), *_block(128, 256), *_block(256, 512))
        self.paths = nn.ModuleList([
            nn.Sequential(*_block(512, 1024), nn.Linear(1024, int(np.prod(img_shape))), nn.Tanh())
            for _ in range(num_client)])

    def forward(self, z):
        hidden = self.model(z)
        img = []
        for path in self.paths:
            out = path(hidden)
            img.append(out.view((out.shape[0], *self.img_shape)))
        return torch.cat(img, dim=0)


def weights_init(m):
    """mixed-gan.py:68-77"""
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find('BatchNorm') != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)
    elif classname.find('Linear') != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class ConvGenerator(nn.Module):
    """model/lsgan.py:3-27 (the ims argument is unused there too)."""

    def __init__(self, ims=None):
        super().__init__()
        self.init_size = 32 // 4
        self.l1 = nn.Sequential(nn.Linear(100, 128 * self.init_size ** 2))
        self.conv_blocks = nn.Sequential(
            nn.Upsample(scale_factor=2),
            nn.Conv2d(128, 128, 3, stride=1, padding=1),
            nn.BatchNorm2d(128, 0.8),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Upsample(scale_factor=2),
            nn.Conv2d(128, 64, 3, stride=1, padding=1),
            nn.BatchNorm2d(64, 0.8),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(64, 1, 3, stride=1, padding=1),
            nn.Tanh(),
        )

    def forward(self, z):
        out = self.l1(z)
        out = out.view(out.shape[0], 128, self.init_size, self.init_size)
        return self.conv_blocks(out)


class ConvDiscriminator(nn.Module):
    """model/lsgan.py:73-99. forward(img, masks): masks = None runs the modules as the reference does (nn.Dropout2d draws
    its own noise); a list of four [B, C] tensors replaces the noise of the four Dropout2d layers (F.dropout2d multiplies by
    a [B, C, 1, 1] tensor of bernoulli(1 - p) / (1 - p)), so that the engine and the oracle drop the same channels."""

    def __init__(self, ims=None):
        super().__init__()

        def discriminator_block(in_filters, out_filters, bn=True):
            block = [nn.Conv2d(in_filters, out_filters, 3, 2, 1), nn.LeakyReLU(0.2, inplace=True), nn.Dropout2d(0.25)]
            if bn:
                block.append(nn.BatchNorm2d(out_filters, 0.8))
            return block

        self.model = nn.Sequential(
            *discriminator_block(1, 16, bn=False),
            *discriminator_block(16, 32),
            *discriminator_block(32, 64),
            *discriminator_block(64, 128),
        )
        ds_size = 32 // 2 ** 4
        self.adv_layer = nn.Linear(128 * ds_size ** 2, 1)

    def forward(self, img, masks=None):
        if masks is None:
            out = self.model(img)
        else:
            out, k = img, 0
            for m in self.model:
                if isinstance(m, nn.Dropout2d):
                    out = out * masks[k][:, :, None, None]
                    k += 1
                else:
                    out = m(out)
        out = out.view(out.shape[0], -1)
        return self.adv_layer(out)


def draw_dropout2d_masks(B, p=0.25):
    """The noise of the discriminator's four Dropout2d layers, drawn from torch's global RNG exactly as one training-mode
    forward draws it (feature dropout: bernoulli(1 - p) of shape [B, C, 1, 1], divided by 1 - p)."""
    import torch.nn.functional as F
    return [F.dropout2d(torch.ones(B, C, 1, 1), p, training=True).reshape(B, C) for C in (16, 32, 64, 128)]
