"""Shallow graphs that exercise every unit type of the engine (train-mode comparisons stay well conditioned)."""
import math
import zlib

import torch

from gdn_pytorch_b200.graph import Graph, Unit


def mini_rtod():
    g = Graph("mini_rtod", 3)
    g.convblock("in", "a1c", "downconv0", 3, 64, 9, 4)
    g.resblock("a1c", "a1", "res64_down1", 64, 9)
    g.convblock("a1", "a2c", "downconv1", 64, 128, 7, 3, 2)
    g.resblock("a2c", "a2", "res128_down1", 128, 7)
    g.convblock("a2", "a3", "upconv3", 128, 64, 7, 3, up=1)
    g.convblock(("a3", "a1c"), "a4a", "conv1x1_64", 128, 64, 1, 0)
    g.resblock("a4a", "a4", "res64_up1", 64, 9)
    g.units.append(Unit("upconv4", ("a4",), "out", 64, 1, 9, 1, 4, tanh=True))
    g.outputs = ("a1", "a2", "a4", "out")
    return g


def mini_dtod():
    g = Graph("mini_dtod", 1)
    g.convblock("in", "b1c", "downconv0", 1, 64, 9, 4)
    g.convblock("b1c", "b2c", "downconv1", 64, 128, 4, 1, 2)
    g.resblock("b2c", "b2", "res128_down1", 128, 7)
    g.convblock("b2", "b3c", "downconv2", 128, 256, 4, 1, 2)
    g.resblock("b3c", "b3", "res256_down1", 256, 5)
    g.convtblock("b3", "b4", "upconv2", 256, 128, 4, 1, 2)
    g.convtblock("b4", "b5", "upconv3", 128, 64, 4, 1, 2)
    g.resblock("b5", "b6", "res64_up1", 64, 9)
    g.units.append(Unit("upconv4", ("b6",), "out", 64, 1, 9, 1, 4, transposed=True, tanh=True))
    g.outputs = ("b2", "b3", "b6", "out")
    return g


def mini_deep512():
    """low-resolution 512-channel part: stride-2 k3, res512 blocks on a tiny map, upsample + k3, concat 1024->512"""
    g = Graph("mini_512", 3)
    g.convblock("in", "c0", "downconv0", 3, 64, 9, 4)
    g.convblock("c0", "c1", "downconv1", 64, 256, 5, 2, 2)
    g.convblock("c1", "c2", "downconv3", 256, 512, 3, 1, 2)
    g.resblock("c2", "c3", "res512_down1", 512, 3)
    g.convblock("c3", "c4", "downconv4", 512, 512, 3, 1, 2)
    g.resblock("c4", "c5", "res512_1", 512, 3)
    g.convblock("c5", "c6", "upconv0", 512, 512, 3, 1, up=1)
    g.convblock(("c6", "c2"), "c7", "conv1x1_512", 1024, 512, 1, 0)
    g.convblock("c7", "c8", "upconv1", 512, 256, 3, 1, up=1)
    g.convblock("c8", "c9", "upconv2", 256, 64, 5, 2, up=1)
    g.units.append(Unit("upconv4", ("c9",), "out", 64, 1, 9, 1, 4, tanh=True))
    g.outputs = ("c3", "c5", "c7", "out")
    return g


def synth_params(graph, seed=0, device="cpu"):
    """state_dict for a graph: conv weights U(+-1/sqrt(fan_in)), random BN affine / running stats"""
    sd = {}

    def gen(key):
        g = torch.Generator()
        g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        return g

    for u in graph.units:
        shp = (u.cin, u.cout, u.k, u.k) if u.transposed else (u.cout, u.cin, u.k, u.k)
        s = 1.0 / math.sqrt(u.cin * u.k * u.k)
        sd[u.conv + ".weight"] = ((torch.rand(shp, generator=gen(u.conv)) * 2 - 1) * s).to(device)
        if u.bn:
            c = u.cout
            sd[u.bn + ".weight"] = (torch.rand(c, generator=gen(u.bn + "w")) + 0.5).to(device)
            sd[u.bn + ".bias"] = ((torch.rand(c, generator=gen(u.bn + "b")) - 0.5) * 0.4).to(device)
            sd[u.bn + ".running_mean"] = ((torch.rand(c, generator=gen(u.bn + "m")) - 0.5) * 0.2).to(device)
            sd[u.bn + ".running_var"] = (torch.rand(c, generator=gen(u.bn + "v")) + 0.5).to(device)
    return sd
