"""Drop-in replacement for the reference's ``src/AE_model_unet.py`` running on hand-written sm_100a kernels.

Same public surface as the reference module (class names, constructor arguments, ``forward(x, istrain)``
signatures, returned tensors, ``state_dict`` keys -- SURVEY.md Appendix B), so ``GDN_main.py`` / ``trainer.py`` /
``eval.py`` and existing ``.pkl`` checkpoints keep working:

    from AE_model_unet import *      ->      from gdn_pytorch_b200.AE_model_unet import *

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ... objects below are PARAMETER CONTAINERS only (they give the reference's
key names, init RNG order and ``load_state_dict`` versioning for free); the arithmetic is done by ``engine.Engine``
through the C ABI of ``libgdn_b200.so``.  There is no PyTorch fallback: CPU tensors or a missing library raise.

Reference: /root/reference/src/AE_model_unet.py (blocks :45-94, AutoEncoder :96-261, AutoEncoder_2 :263-382,
AutoEncoder_Unet :385-483, AutoEncoder_DtoD :485-590, AutoEncoder_Resnet :592-697).
"""
# the reference does `from AE_model_unet import *` and relies on these names being re-exported (GDN_main.py:20)
import torch
from torch.autograd import Variable
import torch.nn as nn
import torch.nn.functional as F
import os
import math
import itertools
import numpy as np

from . import graph as _graph


def _norm(norm, c):
    if norm == 'Batch':
        return nn.BatchNorm2d(c, affine=True, track_running_stats=True)
    return nn.InstanceNorm2d(c, affine=True, track_running_stats=True)


class _Picklable(object):
    """torch.save(model, ...) pickles whole modules (trainer.py:542,868): drop the engine caches"""

    def __getstate__(self):
        d = dict(self.__dict__)
        for k in ("_gdn_engines", "_gdn_graph", "_gdn_epoch"):
            d.pop(k, None)
        return d


class _Block(_Picklable, nn.Module):
    """common forward for the three building blocks: run as a one-block graph on the engine"""
    _kind = None

    def forward(self, x):
        from .module_runtime import run_block
        return run_block(self, x)


class ResidualBlock(_Block):
    """x + BN(conv(ReLU(BN(conv(x))))) with zero padding (reference :45-57)"""
    _kind = "res"

    def __init__(self, dim_in, dim_out, kernel_size, padding):
        super(ResidualBlock, self).__init__()
        self._cfg = (dim_in, dim_out, kernel_size, padding, 1)
        layers = [nn.Conv2d(dim_in, dim_out, kernel_size, 1, padding, bias=False), _norm('Batch', dim_out),
                  nn.ReLU(inplace=True),
                  nn.Conv2d(dim_out, dim_out, kernel_size, 1, padding, bias=False), _norm('Batch', dim_out)]
        self.main = nn.Sequential(*layers)


class ConvBlock(_Block):
    """ReLU(BN(conv(reflect_pad(x)))) (reference :60-77)"""
    _kind = "conv"

    def __init__(self, dim_in, dim_out, kernel_size, padding, stride=1, norm='Batch'):
        super(ConvBlock, self).__init__()
        self._cfg = (dim_in, dim_out, kernel_size, padding, stride)
        self._norm = norm
        self.main = nn.Sequential(nn.ReflectionPad2d(padding),
                                  nn.Conv2d(dim_in, dim_out, kernel_size, stride, padding=0, bias=False),
                                  _norm(norm, dim_out), nn.ReLU(inplace=True))


class ConvTBlock(_Block):
    """ReLU(BN(conv_transpose(x))) (reference :79-94)"""
    _kind = "convT"

    def __init__(self, dim_in, dim_out, kernel_size, padding, stride=1, norm='Batch'):
        super(ConvTBlock, self).__init__()
        self._cfg = (dim_in, dim_out, kernel_size, padding, stride)
        self._norm = norm
        self.main = nn.Sequential(nn.ConvTranspose2d(dim_in, dim_out, kernel_size, stride, padding, bias=False),
                                  _norm(norm, dim_out), nn.ReLU(inplace=True))


def _reference_init(model):
    """the reference's _initialize_weights (:249-261): only nn.Conv2d (NOT ConvTranspose2d) is re-drawn,
    U(+-1/sqrt(cin*kh*kw)); module traversal order = registration order, so the RNG stream matches."""
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            n = m.in_channels
            for k in m.kernel_size:
                n *= k
            stdv = 1. / math.sqrt(n)
            m.weight.data.uniform_(-stdv, stdv)
            if m.bias is not None:
                m.bias.data.uniform_(-stdv, stdv)
        elif isinstance(m, nn.Linear):
            m.weight.data.normal_(0, 0.01)
            m.bias.data.zero_()


class _Net(_Picklable, nn.Module):
    """shared forward of the five autoencoders"""
    _graph_name = None
    _default_istrain = False

    def _announce(self, norm):
        print("- norm : Batch" if norm == 'Batch' else "- norm : Instance")
        self._norm = norm

    def _initialize_weights(self):
        _reference_init(self)

    def forward(self, x, istrain=None):
        from .module_runtime import run_network
        if istrain is None:
            istrain = self._default_istrain
        return run_network(self, x, istrain)

    def gdn_graph(self):
        g = _graph.GRAPHS[self._graph_name]
        return g(self._input_dim) if self._graph_name != "AutoEncoder" else g()


def _res_levels(net, names):
    for name, c, k in names:
        setattr(net, name, ResidualBlock(c, c, k, k // 2))


class AutoEncoder(_Net):
    """RtoD network of --mode RtoD_test / eval.py / depth_extract.py (reference :96-261).  forward default
    istrain=True; the reference's ``x.cuda()`` in forward (:161) is kept: CPU inputs are moved to the GPU."""
    _graph_name = "AutoEncoder"
    _default_istrain = True

    def __init__(self, init_weights=True, norm='Batch', height=128, width=416):
        super(AutoEncoder, self).__init__()
        self.height, self.width, self._input_dim = height, width, 3
        self.downconv0 = nn.Conv2d(3, 64, kernel_size=9, stride=1, padding=4, bias=False)
        self.downconv1 = nn.Conv2d(64, 128, kernel_size=7, stride=2, padding=3, bias=False)
        self.downconv2 = nn.Conv2d(128, 256, kernel_size=5, stride=2, padding=2, bias=False)
        self.downconv3 = nn.Conv2d(256, 512, kernel_size=3, stride=2, padding=1, bias=False)
        _res_levels(self, [("res64_down1", 64, 9), ("res64_down2", 64, 9), ("res64_up1", 64, 9), ("res64_up2", 64, 9),
                           ("res128_down1", 128, 7), ("res128_down2", 128, 7), ("res128_up1", 128, 7),
                           ("res128_up2", 128, 7), ("res256_down1", 256, 5), ("res256_down2", 256, 5),
                           ("res256_up1", 256, 5), ("res256_up2", 256, 5)] +
                    [("res512_%d" % i, 512, 3) for i in range(1, 7)])
        self.upconv0 = nn.ConvTranspose2d(512, 256, kernel_size=3, stride=1, padding=1, bias=False)
        self.upconv1 = nn.ConvTranspose2d(256, 128, kernel_size=5, stride=1, padding=2, bias=False)
        self.upconv2 = nn.ConvTranspose2d(128, 64, kernel_size=7, stride=1, padding=3, bias=False)
        self.upconv3 = nn.Conv2d(64, 1, kernel_size=9, stride=1, padding=4, bias=False)
        self.conv1x1_64 = nn.Conv2d(128, 64, kernel_size=1, stride=1, padding=0, bias=False)
        self.conv1x1_128 = nn.Conv2d(256, 128, kernel_size=1, stride=1, padding=0, bias=False)
        self.conv1x1_256 = nn.Conv2d(512, 256, kernel_size=1, stride=1, padding=0, bias=False)
        self.upsampling = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        self._announce(norm)
        for name, c in (("N64_down", 64), ("N128_down", 128), ("N256_down", 256), ("N512_down", 512), ("N64_up", 64),
                        ("N128_up", 128), ("N256_up", 256)):
            setattr(self, name, _norm(norm, c))
        self.ReLU = nn.ReLU(inplace=True)
        if init_weights:
            self._initialize_weights()

    def forward(self, x, istrain=True):
        return super(AutoEncoder, self).forward(x.cuda(), istrain)


def _encoder(net, input_dim, down):
    """downconv0..4 of the ConvBlock-based networks; ``down`` = [(cin, cout, k, pad)] for the stride-2 levels"""
    net.downconv0 = ConvBlock(input_dim, 64, kernel_size=9, stride=1, padding=4)
    for i, (ci, co, k, p) in enumerate(down):
        setattr(net, "downconv%d" % (i + 1), ConvBlock(ci, co, kernel_size=k, stride=2, padding=p))


_RES_2 = [("res64_down1", 64, 9), ("res64_up1", 64, 9), ("res128_down1", 128, 7), ("res128_up1", 128, 7),
          ("res256_down1", 256, 5), ("res256_up1", 256, 5), ("res512_down1", 512, 3), ("res512_up1", 512, 3),
          ("res512_down2", 512, 3), ("res512_up2", 512, 3)] + [("res512_%d" % i, 512, 3) for i in range(1, 7)]
_DOWN_K7 = [(64, 128, 7, 3), (128, 256, 5, 2), (256, 512, 3, 1), (512, 512, 3, 1)]
_DOWN_K4 = [(64, 128, 4, 1), (128, 256, 4, 1), (256, 512, 4, 1), (512, 512, 4, 1)]


def _skip_1x1(net):
    net.conv1x1_64 = ConvBlock(128, 64, kernel_size=1, stride=1, padding=0)
    net.conv1x1_128 = ConvBlock(256, 128, kernel_size=1, stride=1, padding=0)
    net.conv1x1_256 = ConvBlock(512, 256, kernel_size=1, stride=1, padding=0)
    net.conv1x1_512 = ConvBlock(1024, 512, kernel_size=1, stride=1, padding=0)


class AutoEncoder_2(_Net):
    """RtoD network used for TRAINING (--mode RtoD), reference :263-382"""
    _graph_name = "AutoEncoder_2"

    def __init__(self, init_weights=True, norm='Batch', input_dim=3, height=128, width=416):
        super(AutoEncoder_2, self).__init__()
        self._announce(norm)
        self.height, self.width, self._input_dim = height, width, input_dim
        _encoder(self, input_dim, _DOWN_K7)
        _res_levels(self, _RES_2)
        self.upconv0 = ConvBlock(512, 512, kernel_size=3, stride=1, padding=1)
        self.upconv1 = ConvBlock(512, 256, kernel_size=3, stride=1, padding=1)
        self.upconv2 = ConvBlock(256, 128, kernel_size=5, stride=1, padding=2)
        self.upconv3 = ConvBlock(128, 64, kernel_size=7, stride=1, padding=3)
        self.upconv4 = nn.Conv2d(64, 1, kernel_size=9, stride=1, padding=4, bias=False)
        _skip_1x1(self)
        self.upsampling = nn.functional.interpolate
        self.ReLU = nn.ReLU(inplace=True)
        if init_weights:
            self._initialize_weights()


class AutoEncoder_Unet(_Net):
    """ablation without residual blocks (reference :385-483; never instantiated by live code)"""
    _graph_name = "AutoEncoder_Unet"

    def __init__(self, init_weights=True, norm='Batch', input_dim=3, height=128, width=416):
        super(AutoEncoder_Unet, self).__init__()
        self._announce(norm)
        self.height, self.width, self._input_dim = height, width, input_dim
        _encoder(self, input_dim, _DOWN_K7)
        for i in range(1, 7):
            setattr(self, "conv512_%d" % i, ConvBlock(512, 512, kernel_size=3, stride=1, padding=1))
        self.upconv0 = ConvBlock(512, 512, kernel_size=3, stride=1, padding=1)
        self.upconv1 = ConvBlock(512, 256, kernel_size=3, stride=1, padding=1)
        self.upconv2 = ConvBlock(256, 128, kernel_size=5, stride=1, padding=2)
        self.upconv3 = ConvBlock(128, 64, kernel_size=7, stride=1, padding=3)
        self.upconv4 = nn.Conv2d(64, 1, kernel_size=9, stride=1, padding=4, bias=False)
        _skip_1x1(self)
        self.upsampling = nn.functional.interpolate
        self.ReLU = nn.ReLU(inplace=True)
        if init_weights:
            self._initialize_weights()


class AutoEncoder_DtoD(_Net):
    """depth-to-depth autoencoder (reference :485-590): k4/s2 down-convs, k4/s2 ConvTranspose up-convs, no skips"""
    _graph_name = "AutoEncoder_DtoD"

    def __init__(self, init_weights=True, norm='Batch', input_dim=1, height=128, width=416):
        super(AutoEncoder_DtoD, self).__init__()
        self._announce(norm)
        self.height, self.width, self._input_dim = height, width, input_dim
        _encoder(self, input_dim, _DOWN_K4)
        _res_levels(self, _RES_2)
        self.upconv0 = ConvTBlock(512, 512, kernel_size=4, stride=2, padding=1)
        self.upconv1 = ConvTBlock(512, 256, kernel_size=4, stride=2, padding=1)
        self.upconv2 = ConvTBlock(256, 128, kernel_size=4, stride=2, padding=1)
        self.upconv3 = ConvTBlock(128, 64, kernel_size=4, stride=2, padding=1)
        self.upconv4 = nn.ConvTranspose2d(64, 1, kernel_size=9, stride=1, padding=4, bias=False)
        self.upsampling = nn.functional.interpolate
        self.ReLU = nn.ReLU(inplace=True)
        if init_weights:
            self._initialize_weights()


class AutoEncoder_Resnet(_Net):
    """ablation without skip connections (reference :592-697; never instantiated by live code)"""
    _graph_name = "AutoEncoder_Resnet"

    def __init__(self, init_weights=True, norm='Batch', input_dim=3, height=128, width=416):
        super(AutoEncoder_Resnet, self).__init__()
        self._announce(norm)
        self.height, self.width, self._input_dim = height, width, input_dim
        _encoder(self, input_dim, _DOWN_K7)
        _res_levels(self, _RES_2)
        self.upconv0 = ConvTBlock(512, 512, kernel_size=3, stride=1, padding=1)
        self.upconv1 = ConvTBlock(512, 256, kernel_size=3, stride=1, padding=1)
        self.upconv2 = ConvTBlock(256, 128, kernel_size=5, stride=1, padding=2)
        self.upconv3 = ConvTBlock(128, 64, kernel_size=7, stride=1, padding=3)
        self.upconv4 = nn.Conv2d(64, 1, kernel_size=9, stride=1, padding=4, bias=False)
        self.upsampling = nn.functional.interpolate
        self.ReLU = nn.ReLU(inplace=True)
        if init_weights:
            self._initialize_weights()
