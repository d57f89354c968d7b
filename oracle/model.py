"""Functional fp32 restatement of the reference autoencoders (TEST INFRASTRUCTURE).

Every function takes a plain ``state_dict`` (keys WITHOUT the ``module.`` prefix,
Appendix B of SURVEY.md) and computes what the reference ``nn.Module`` computes,
with torch.nn.functional on CPU (or CUDA with TF32 off).  Autograd works through
all of it, so backward parity uses ``torch.autograd.grad`` on these functions.

``bf16=True`` rounds exactly where the B200 path rounds (conv operands and stored
conv outputs), which gives a tight second oracle for the train-mode network whose
fp32 comparison is ill-conditioned at random init ("Tolerances").

Reference: /root/reference/src/AE_model_unet.py
  ResidualBlock :45-57   ConvBlock :60-77   ConvTBlock :79-94
  AutoEncoder :96-261    AutoEncoder_2 :263-382   AutoEncoder_DtoD :485-590
"""
import torch
import torch.nn.functional as F

EPS = 1e-5
MOMENTUM = 0.1


def _r(x, on):
    """round-trip through bf16 when emulating the device path"""
    return x.to(torch.bfloat16).to(torch.float32) if on else x


class Ctx:
    """carries mode flags and collects BN batch statistics (for running-stat parity)"""

    def __init__(self, sd, train, bf16=False, update_running=False):
        self.sd, self.train, self.bf16, self.update_running = sd, train, bf16, update_running

    def conv(self, x, key, stride=1, padding=0, reflect=0):
        w = self.sd[key + ".weight"]
        if reflect:
            x = F.pad(x, (reflect,) * 4, mode="reflect")  # nn.ReflectionPad2d, AE_model_unet.py:66
        y = F.conv2d(_r(x, self.bf16), _r(w, self.bf16), None, stride, padding)
        return y

    def convT(self, x, key, stride=1, padding=0):
        w = self.sd[key + ".weight"]  # (Cin, Cout, k, k), AE_model_unet.py:85
        return F.conv_transpose2d(_r(x, self.bf16), _r(w, self.bf16), None, stride, padding)

    def bn(self, x, key):
        sd = self.sd
        rm, rv = sd[key + ".running_mean"], sd[key + ".running_var"]
        if self.train:
            if self.bf16:
                # device path: statistics from the fp32 accumulators, normalisation of the bf16-stored value
                mean = x.mean((0, 2, 3))
                var = x.var((0, 2, 3), unbiased=False)
                xs = x.to(torch.float16).to(torch.float32)   # raw conv outputs are stored as fp16 on the device
                y = (xs - mean[None, :, None, None]) * torch.rsqrt(var + EPS)[None, :, None, None]
                y = y * sd[key + ".weight"][None, :, None, None] + sd[key + ".bias"][None, :, None, None]
                if self.update_running:
                    n = x.numel() // x.shape[1]
                    rm.mul_(1 - MOMENTUM).add_(MOMENTUM * mean.detach())
                    rv.mul_(1 - MOMENTUM).add_(MOMENTUM * var.detach() * n / (n - 1))
                return y
            if self.update_running:
                return F.batch_norm(x, rm, rv, sd[key + ".weight"], sd[key + ".bias"], True, MOMENTUM, EPS)
            return F.batch_norm(x, None, None, sd[key + ".weight"], sd[key + ".bias"], True, MOMENTUM, EPS)
        return F.batch_norm(x, rm, rv, sd[key + ".weight"], sd[key + ".bias"], False, MOMENTUM, EPS)

    # ---- blocks
    def res(self, x, name, k):
        """ResidualBlock: x + BN(conv(ReLU(BN(conv(x))))), zero padding k//2 (AE_model_unet.py:45-57)"""
        p = k // 2
        h = F.relu(self.bn(self.conv(x, name + ".main.0", 1, p), name + ".main.1"))
        h = self.bn(self.conv(h, name + ".main.3", 1, p), name + ".main.4")
        return x + h

    def cblock(self, x, name, k, pad, stride=1):
        """ConvBlock: ReLU(BN(conv(reflect_pad(x)))) (AE_model_unet.py:60-77)"""
        return F.relu(self.bn(self.conv(x, name + ".main.1", stride, 0, reflect=pad), name + ".main.2"))

    def tblock(self, x, name, k, pad, stride):
        """ConvTBlock: ReLU(BN(conv_transpose(x))) (AE_model_unet.py:79-94)"""
        return F.relu(self.bn(self.convT(x, name + ".main.0", stride, pad), name + ".main.1"))


def _finish(y, H, W):
    return torch.tanh(y).reshape(-1, 1, H, W)


def autoencoder_2(sd, x, istrain=False, train=False, bf16=False, update_running=False):
    """AutoEncoder_2.forward (AE_model_unet.py:312-368)."""
    c = Ctx(sd, train, bf16, update_running)
    H, W = x.shape[2], x.shape[3]
    x1c = c.cblock(x, "downconv0", 9, 4)
    x1 = c.res(x1c, "res64_down1", 9)
    x2c = c.cblock(x1, "downconv1", 7, 3, 2)
    x2 = c.res(x2c, "res128_down1", 7)
    x3c = c.cblock(x2, "downconv2", 5, 2, 2)
    x3 = c.res(x3c, "res256_down1", 5)
    x4c = c.cblock(x3, "downconv3", 3, 1, 2)
    x4 = c.res(c.res(x4c, "res512_down1", 3), "res512_down2", 3)
    x6 = c.cblock(x4, "downconv4", 3, 1, 2)
    for i in range(1, 7):
        x6 = c.res(x6, "res512_%d" % i, 3)

    def up(t):
        return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)

    x8 = c.cblock(torch.cat((c.cblock(up(x6), "upconv0", 3, 1), x4c), 1), "conv1x1_512", 1, 0)
    x8 = c.res(c.res(x8, "res512_up1", 3), "res512_up2", 3)
    x10 = c.cblock(torch.cat((c.cblock(up(x8), "upconv1", 3, 1), x3c), 1), "conv1x1_256", 1, 0)
    x10 = c.res(x10, "res256_up1", 5)
    x12 = c.cblock(torch.cat((c.cblock(up(x10), "upconv2", 5, 2), x2c), 1), "conv1x1_128", 1, 0)
    x12 = c.res(x12, "res128_up1", 7)
    x14 = c.cblock(torch.cat((c.cblock(up(x12), "upconv3", 7, 3), x1c), 1), "conv1x1_64", 1, 0)
    x14 = c.res(x14, "res64_up1", 9)
    x15 = _finish(c.conv(x14, "upconv4", 1, 4), H, W)
    return (x1, x2, x4, x6, x8, x12, x14, x15) if istrain is True else x15


def autoencoder_dtod(sd, x, istrain=False, train=False, bf16=False, update_running=False, encoder_only=False):
    """AutoEncoder_DtoD.forward (AE_model_unet.py:527-574).  ``encoder_only`` stops after x6 and returns
    (x1, x2, x4, x6): the only tensors the RtoD guidance loss consumes (trainer.py:700,703)."""
    c = Ctx(sd, train, bf16, update_running)
    H, W = x.shape[2], x.shape[3]
    x1 = c.res(c.cblock(x, "downconv0", 9, 4), "res64_down1", 9)
    x2 = c.res(c.cblock(x1, "downconv1", 4, 1, 2), "res128_down1", 7)
    x3 = c.res(c.cblock(x2, "downconv2", 4, 1, 2), "res256_down1", 5)
    x4 = c.res(c.res(c.cblock(x3, "downconv3", 4, 1, 2), "res512_down1", 3), "res512_down2", 3)
    x6 = c.cblock(x4, "downconv4", 4, 1, 2)
    for i in range(1, 7):
        x6 = c.res(x6, "res512_%d" % i, 3)
    if encoder_only:
        return x1, x2, x4, x6
    x8 = c.res(c.res(c.tblock(x6, "upconv0", 4, 1, 2), "res512_up1", 3), "res512_up2", 3)
    x10 = c.res(c.tblock(x8, "upconv1", 4, 1, 2), "res256_up1", 5)
    x12 = c.res(c.tblock(x10, "upconv2", 4, 1, 2), "res128_up1", 7)
    x14 = c.res(c.tblock(x12, "upconv3", 4, 1, 2), "res64_up1", 9)
    x15 = _finish(c.convT(x14, "upconv4", 1, 4), H, W)
    return (x1, x2, x4, x6, x8, x12, x14, x15) if istrain is True else x15


def autoencoder(sd, x, istrain=True, train=False, bf16=False, update_running=False):
    """AutoEncoder.forward (AE_model_unet.py:160-246), the RtoD_test / demo class.  Note :191-195: the
    ReLU after N512_down is in place, so res512_1 sees the ReLU'd tensor although the code passes x17."""
    c = Ctx(sd, train, bf16, update_running)
    H, W = x.shape[2], x.shape[3]

    def up(t):
        return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)

    x3 = F.relu(c.bn(c.conv(x, "downconv0", 1, 4), "N64_down"))
    x5 = c.res(c.res(x3, "res64_down1", 9), "res64_down2", 9)
    x8 = F.relu(c.bn(c.conv(x5, "downconv1", 2, 3), "N128_down"))
    x10 = c.res(c.res(x8, "res128_down1", 7), "res128_down2", 7)
    x13 = F.relu(c.bn(c.conv(x10, "downconv2", 2, 2), "N256_down"))
    x15 = c.res(c.res(x13, "res256_down1", 5), "res256_down2", 5)
    x23 = F.relu(c.bn(c.conv(x15, "downconv3", 2, 1), "N512_down"))
    for i in range(1, 7):
        x23 = c.res(x23, "res512_%d" % i, 3)
    x27 = F.relu(c.bn(c.convT(up(x23), "upconv0", 1, 1), "N256_up"))
    x27 = c.conv(torch.cat((x27, x15), 1), "conv1x1_256")
    x29 = c.res(c.res(x27, "res256_up1", 5), "res256_up2", 5)
    x33 = F.relu(c.bn(c.convT(up(x29), "upconv1", 1, 2), "N128_up"))
    x33 = c.conv(torch.cat((x33, x10), 1), "conv1x1_128")
    x35 = c.res(c.res(x33, "res128_up1", 7), "res128_up2", 7)
    x39 = F.relu(c.bn(c.convT(up(x35), "upconv2", 1, 3), "N64_up"))
    x39 = c.conv(torch.cat((x39, x5), 1), "conv1x1_64")
    x41 = c.res(c.res(x39, "res64_up1", 9), "res64_up2", 9)
    x44 = _finish(c.conv(x41, "upconv3", 1, 4), H, W)
    return (x5, x10, x15, x23, x29, x35, x41, x44) if istrain is True else x44


FORWARDS = {
    "AutoEncoder_2": autoencoder_2,
    "AutoEncoder_DtoD": autoencoder_dtod,
    "AutoEncoder": autoencoder,
}


def strip_module_prefix(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
