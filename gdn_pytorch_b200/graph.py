"""Layer graphs of the GDN autoencoders as lists of fused conv units.

A *unit* is one convolution together with everything the reference wraps around it
(input upsampling / padding, BatchNorm, ReLU, residual add, tanh).  The graphs restate the data flow of
/root/reference/src/AE_model_unet.py (forward methods at :160-246, :312-368, :425-470, :527-574, :633-682);
the engine (engine.py) turns them into kernel launches.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple


@dataclass
class Unit:
    conv: str                       # state_dict prefix of the conv weight, e.g. "res64_down1.main.0"
    srcs: Tuple[str, ...]           # input tensor name(s); two names = channel concat (torch.cat(..., 1))
    out: str                        # output tensor name
    cin: int
    cout: int
    k: int
    stride: int = 1
    pad: int = 0
    reflect: bool = False           # nn.ReflectionPad2d(pad) in front of a pad-0 conv
    transposed: bool = False        # nn.ConvTranspose2d (weight is (cin, cout, k, k))
    up: int = 0                     # input bilinear x2 first: 1 = align_corners False, 2 = True
    bn: Optional[str] = None        # state_dict prefix of the BatchNorm2d
    relu: bool = False
    resid: Optional[str] = None     # tensor added after BN (ResidualBlock identity)
    tanh: bool = False


@dataclass
class Graph:
    name: str
    cin: int
    units: List[Unit] = field(default_factory=list)
    outputs: Tuple[str, ...] = ()   # the 8 tensors returned when istrain is True; the last one is the depth map
    encoder_outputs: Tuple[str, ...] = ()  # (x1, x2, x4, x6): all the guidance loss needs (trainer.py:700,703)

    # ---- builders mirroring the reference blocks
    def convblock(self, src, out, name, cin, cout, k, pad, stride=1, up=0):
        """ConvBlock (AE_model_unet.py:60-77): ReflectionPad2d -> Conv2d -> BN -> ReLU; keys main.1 / main.2"""
        self.units.append(Unit(name + ".main.1", tuple(src) if isinstance(src, (tuple, list)) else (src,), out, cin, cout,
                               k, stride, pad, reflect=pad > 0, up=up, bn=name + ".main.2", relu=True))

    def resblock(self, src, out, name, c, k):
        """ResidualBlock (AE_model_unet.py:45-57): x + BN(conv(ReLU(BN(conv(x))))); keys main.0/1/3/4"""
        p = k // 2
        mid = name + ":h"
        self.units.append(Unit(name + ".main.0", (src,), mid, c, c, k, 1, p, bn=name + ".main.1", relu=True))
        self.units.append(Unit(name + ".main.3", (mid,), out, c, c, k, 1, p, bn=name + ".main.4", resid=src))

    def convtblock(self, src, out, name, cin, cout, k, pad, stride, up=0):
        """ConvTBlock (AE_model_unet.py:79-94): ConvTranspose2d -> BN -> ReLU; keys main.0 / main.1"""
        self.units.append(Unit(name + ".main.0", (src,), out, cin, cout, k, stride, pad, transposed=True, up=up,
                               bn=name + ".main.1", relu=True))


def graph_autoencoder_2(cin=3) -> Graph:
    """AutoEncoder_2 (AE_model_unet.py:263-368)"""
    g = Graph("AutoEncoder_2", cin)
    g.convblock("in", "x1c", "downconv0", cin, 64, 9, 4)
    g.resblock("x1c", "x1", "res64_down1", 64, 9)
    g.convblock("x1", "x2c", "downconv1", 64, 128, 7, 3, 2)
    g.resblock("x2c", "x2", "res128_down1", 128, 7)
    g.convblock("x2", "x3c", "downconv2", 128, 256, 5, 2, 2)
    g.resblock("x3c", "x3", "res256_down1", 256, 5)
    g.convblock("x3", "x4c", "downconv3", 256, 512, 3, 1, 2)
    g.resblock("x4c", "x4a", "res512_down1", 512, 3)
    g.resblock("x4a", "x4", "res512_down2", 512, 3)
    g.convblock("x4", "x5", "downconv4", 512, 512, 3, 1, 2)
    prev = "x5"
    for i in range(1, 7):
        nxt = "x6" if i == 6 else "x6_%d" % i
        g.resblock(prev, nxt, "res512_%d" % i, 512, 3)
        prev = nxt
    g.convblock("x6", "x7", "upconv0", 512, 512, 3, 1, up=1)
    g.convblock(("x7", "x4c"), "x8a", "conv1x1_512", 1024, 512, 1, 0)
    g.resblock("x8a", "x8b", "res512_up1", 512, 3)
    g.resblock("x8b", "x8", "res512_up2", 512, 3)
    g.convblock("x8", "x9", "upconv1", 512, 256, 3, 1, up=1)
    g.convblock(("x9", "x3c"), "x10a", "conv1x1_256", 512, 256, 1, 0)
    g.resblock("x10a", "x10", "res256_up1", 256, 5)
    g.convblock("x10", "x11", "upconv2", 256, 128, 5, 2, up=1)
    g.convblock(("x11", "x2c"), "x12a", "conv1x1_128", 256, 128, 1, 0)
    g.resblock("x12a", "x12", "res128_up1", 128, 7)
    g.convblock("x12", "x13", "upconv3", 128, 64, 7, 3, up=1)
    g.convblock(("x13", "x1c"), "x14a", "conv1x1_64", 128, 64, 1, 0)
    g.resblock("x14a", "x14", "res64_up1", 64, 9)
    g.units.append(Unit("upconv4", ("x14",), "x15", 64, 1, 9, 1, 4, tanh=True))
    g.outputs = ("x1", "x2", "x4", "x6", "x8", "x12", "x14", "x15")
    g.encoder_outputs = ("x1", "x2", "x4", "x6")
    return g


def graph_autoencoder_dtod(cin=1) -> Graph:
    """AutoEncoder_DtoD (AE_model_unet.py:485-574): k4/s2 reflect-pad down-convs, k4/s2 ConvTranspose up-convs"""
    g = Graph("AutoEncoder_DtoD", cin)
    g.convblock("in", "x1c", "downconv0", cin, 64, 9, 4)
    g.resblock("x1c", "x1", "res64_down1", 64, 9)
    g.convblock("x1", "x2c", "downconv1", 64, 128, 4, 1, 2)
    g.resblock("x2c", "x2", "res128_down1", 128, 7)
    g.convblock("x2", "x3c", "downconv2", 128, 256, 4, 1, 2)
    g.resblock("x3c", "x3", "res256_down1", 256, 5)
    g.convblock("x3", "x4c", "downconv3", 256, 512, 4, 1, 2)
    g.resblock("x4c", "x4a", "res512_down1", 512, 3)
    g.resblock("x4a", "x4", "res512_down2", 512, 3)
    g.convblock("x4", "x5", "downconv4", 512, 512, 4, 1, 2)
    prev = "x5"
    for i in range(1, 7):
        nxt = "x6" if i == 6 else "x6_%d" % i
        g.resblock(prev, nxt, "res512_%d" % i, 512, 3)
        prev = nxt
    g.convtblock("x6", "x7", "upconv0", 512, 512, 4, 1, 2)
    g.resblock("x7", "x8b", "res512_up1", 512, 3)
    g.resblock("x8b", "x8", "res512_up2", 512, 3)
    g.convtblock("x8", "x9", "upconv1", 512, 256, 4, 1, 2)
    g.resblock("x9", "x10", "res256_up1", 256, 5)
    g.convtblock("x10", "x11", "upconv2", 256, 128, 4, 1, 2)
    g.resblock("x11", "x12", "res128_up1", 128, 7)
    g.convtblock("x12", "x13", "upconv3", 128, 64, 4, 1, 2)
    g.resblock("x13", "x14", "res64_up1", 64, 9)
    g.units.append(Unit("upconv4", ("x14",), "x15", 64, 1, 9, 1, 4, transposed=True, tanh=True))
    g.outputs = ("x1", "x2", "x4", "x6", "x8", "x12", "x14", "x15")
    g.encoder_outputs = ("x1", "x2", "x4", "x6")
    return g


def graph_autoencoder() -> Graph:
    """AutoEncoder (AE_model_unet.py:96-246), the RtoD_test / eval / demo class.  Top-level convs + BNs,
    two ResidualBlocks per level, nn.Upsample(align_corners=True), stride-1 ConvTranspose2d up-convs, plain 1x1
    convs on the concats.  (:191-195: ReLU is in place, so res512_1 consumes the ReLU'd tensor.)"""
    g = Graph("AutoEncoder", 3)
    U = g.units
    U.append(Unit("downconv0", ("in",), "x3", 3, 64, 9, 1, 4, bn="N64_down", relu=True))
    g.resblock("x3", "x4", "res64_down1", 64, 9)
    g.resblock("x4", "x5", "res64_down2", 64, 9)
    U.append(Unit("downconv1", ("x5",), "x8", 64, 128, 7, 2, 3, bn="N128_down", relu=True))
    g.resblock("x8", "x9", "res128_down1", 128, 7)
    g.resblock("x9", "x10", "res128_down2", 128, 7)
    U.append(Unit("downconv2", ("x10",), "x13", 128, 256, 5, 2, 2, bn="N256_down", relu=True))
    g.resblock("x13", "x14", "res256_down1", 256, 5)
    g.resblock("x14", "x15", "res256_down2", 256, 5)
    U.append(Unit("downconv3", ("x15",), "x17r", 256, 512, 3, 2, 1, bn="N512_down", relu=True))
    prev = "x17r"
    for i in range(1, 7):
        nxt = "x%d" % (17 + i)
        g.resblock(prev, nxt, "res512_%d" % i, 512, 3)
        prev = nxt
    U.append(Unit("upconv0", ("x23",), "x27r", 512, 256, 3, 1, 1, transposed=True, up=2, bn="N256_up", relu=True))
    U.append(Unit("conv1x1_256", ("x27r", "x15"), "x27", 512, 256, 1))
    g.resblock("x27", "x28", "res256_up1", 256, 5)
    g.resblock("x28", "x29", "res256_up2", 256, 5)
    U.append(Unit("upconv1", ("x29",), "x33r", 256, 128, 5, 1, 2, transposed=True, up=2, bn="N128_up", relu=True))
    U.append(Unit("conv1x1_128", ("x33r", "x10"), "x33", 256, 128, 1))
    g.resblock("x33", "x34", "res128_up1", 128, 7)
    g.resblock("x34", "x35", "res128_up2", 128, 7)
    U.append(Unit("upconv2", ("x35",), "x39r", 128, 64, 7, 1, 3, transposed=True, up=2, bn="N64_up", relu=True))
    U.append(Unit("conv1x1_64", ("x39r", "x5"), "x39", 128, 64, 1))
    g.resblock("x39", "x40", "res64_up1", 64, 9)
    g.resblock("x40", "x41", "res64_up2", 64, 9)
    U.append(Unit("upconv3", ("x41",), "x44", 64, 1, 9, 1, 4, tanh=True))
    g.outputs = ("x5", "x10", "x15", "x23", "x29", "x35", "x41", "x44")
    return g


def graph_autoencoder_unet(cin=3) -> Graph:
    """AutoEncoder_Unet (AE_model_unet.py:385-470): ablation without residual blocks (never instantiated live)"""
    g = Graph("AutoEncoder_Unet", cin)
    g.convblock("in", "x1c", "downconv0", cin, 64, 9, 4)
    g.convblock("x1c", "x2c", "downconv1", 64, 128, 7, 3, 2)
    g.convblock("x2c", "x3c", "downconv2", 128, 256, 5, 2, 2)
    g.convblock("x3c", "x4c", "downconv3", 256, 512, 3, 1, 2)
    g.convblock("x4c", "x5", "downconv4", 512, 512, 3, 1, 2)
    prev = "x5"
    for i in range(1, 7):
        nxt = "x6" if i == 6 else "x6_%d" % i
        g.convblock(prev, nxt, "conv512_%d" % i, 512, 512, 3, 1)
        prev = nxt
    g.convblock("x6", "x7", "upconv0", 512, 512, 3, 1, up=1)
    g.convblock(("x7", "x4c"), "x8", "conv1x1_512", 1024, 512, 1, 0)
    g.convblock("x8", "x9", "upconv1", 512, 256, 3, 1, up=1)
    g.convblock(("x9", "x3c"), "x10", "conv1x1_256", 512, 256, 1, 0)
    g.convblock("x10", "x11", "upconv2", 256, 128, 5, 2, up=1)
    g.convblock(("x11", "x2c"), "x12", "conv1x1_128", 256, 128, 1, 0)
    g.convblock("x12", "x13", "upconv3", 128, 64, 7, 3, up=1)
    g.convblock(("x13", "x1c"), "x14", "conv1x1_64", 128, 64, 1, 0)
    g.units.append(Unit("upconv4", ("x14",), "x15", 64, 1, 9, 1, 4, tanh=True))
    g.outputs = ("x1c", "x2c", "x4c", "x6", "x8", "x12", "x14", "x15")
    return g


def graph_autoencoder_resnet(cin=3) -> Graph:
    """AutoEncoder_Resnet (AE_model_unet.py:592-682): AutoEncoder_2 without skips, stride-1 ConvTBlocks after x2 up"""
    g = Graph("AutoEncoder_Resnet", cin)
    g.convblock("in", "x1c", "downconv0", cin, 64, 9, 4)
    g.resblock("x1c", "x1", "res64_down1", 64, 9)
    g.convblock("x1", "x2c", "downconv1", 64, 128, 7, 3, 2)
    g.resblock("x2c", "x2", "res128_down1", 128, 7)
    g.convblock("x2", "x3c", "downconv2", 128, 256, 5, 2, 2)
    g.resblock("x3c", "x3", "res256_down1", 256, 5)
    g.convblock("x3", "x4c", "downconv3", 256, 512, 3, 1, 2)
    g.resblock("x4c", "x4a", "res512_down1", 512, 3)
    g.resblock("x4a", "x4", "res512_down2", 512, 3)
    g.convblock("x4", "x5", "downconv4", 512, 512, 3, 1, 2)
    prev = "x5"
    for i in range(1, 7):
        nxt = "x6" if i == 6 else "x6_%d" % i
        g.resblock(prev, nxt, "res512_%d" % i, 512, 3)
        prev = nxt
    g.convtblock("x6", "x7", "upconv0", 512, 512, 3, 1, 1, up=1)
    g.resblock("x7", "x8b", "res512_up1", 512, 3)
    g.resblock("x8b", "x8", "res512_up2", 512, 3)
    g.convtblock("x8", "x9", "upconv1", 512, 256, 3, 1, 1, up=1)
    g.resblock("x9", "x10", "res256_up1", 256, 5)
    g.convtblock("x10", "x11", "upconv2", 256, 128, 5, 2, 1, up=1)
    g.resblock("x11", "x12", "res128_up1", 128, 7)
    g.convtblock("x12", "x13", "upconv3", 128, 64, 7, 3, 1, up=1)
    g.resblock("x13", "x14", "res64_up1", 64, 9)
    g.units.append(Unit("upconv4", ("x14",), "x15", 64, 1, 9, 1, 4, tanh=True))
    g.outputs = ("x1", "x2", "x4", "x6", "x8", "x12", "x14", "x15")
    return g


GRAPHS = {
    "AutoEncoder": graph_autoencoder,
    "AutoEncoder_2": graph_autoencoder_2,
    "AutoEncoder_Unet": graph_autoencoder_unet,
    "AutoEncoder_DtoD": graph_autoencoder_dtod,
    "AutoEncoder_Resnet": graph_autoencoder_resnet,
}
