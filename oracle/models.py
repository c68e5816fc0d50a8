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
