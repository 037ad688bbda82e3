"""Oracle: the whole MinkUNet forward as a function of a state_dict (CPU, torch).
Test infrastructure only.

Topology follows the reference ``models/minkunet.py:59-132`` (layers), ``:134-219`` (data flow)
and ``models/resnet.py:90-122`` (``_make_layer``: a 1x1 conv + BN ``downsample`` whenever
``inplanes != planes * expansion``); block internals follow
``MinkowskiEngine.modules.resnet_block.BasicBlock / Bottleneck`` [ME-upstream, SURVEY a14].
Parameters are looked up by their ME ``state_dict`` names, so the same dict drives both this
oracle and the CUDA-backed modules.
"""
from __future__ import annotations

import numpy as np
import torch

from . import conv as oc
from .coords import CoordLevels

ARCHS = {
    # name: (block, LAYERS, PLANES, INIT_DIM)   ref models/minkunet.py:528-591
    "MinkUNet14A": ("basic", (1,) * 8, (32, 64, 128, 256, 128, 128, 96, 96), 32),
    "MinkUNet18A": ("basic", (2,) * 8, (32, 64, 128, 256, 256, 128, 96, 96), 32),
    "MinkUNet34C": ("basic", (2, 3, 4, 6, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 96, 96), 32),
    "MinkUNet34RC": ("basic", (2, 3, 4, 6, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 96, 96), 32),
    "MinkUNet50": ("bottleneck", (2, 3, 4, 6, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 96, 96), 32),
}

DOWN_CONVS = ("conv1p1s2", "conv2p2s2", "conv3p4s2", "conv4p8s2")
UP_CONVS = ("convtr4p16s2", "convtr5p8s2", "convtr6p4s2", "convtr7p2s2")


class OracleMinkUNet:
    def __init__(self, params: dict, arch: str = "MinkUNet34C", training: bool = True, prefix: str = ""):
        self.p = params
        self.block, self.layers, self.planes, self.init_dim = ARCHS[arch]
        self.training = training
        self.prefix = prefix

    # -- primitives -----------------------------------------------------------------------
    def _w(self, name):
        return self.p[self.prefix + name]

    def _bn(self, x, name):
        return oc.batch_norm(x, self._w(name + ".bn.weight"), self._w(name + ".bn.bias"),
                             self._w(name + ".bn.running_mean"), self._w(name + ".bn.running_var"), self.training)

    def _block(self, x, name, nbr3):
        if self.block == "basic":
            out = oc.conv_table(x, nbr3, self._w(name + ".conv1.kernel"))
            out = torch.relu(self._bn(out, name + ".norm1"))
            out = oc.conv_table(out, nbr3, self._w(name + ".conv2.kernel"))
            out = self._bn(out, name + ".norm2")
        else:
            out = oc.conv_1x1(x, self._w(name + ".conv1.kernel"))
            out = torch.relu(self._bn(out, name + ".norm1"))
            out = oc.conv_table(out, nbr3, self._w(name + ".conv2.kernel"))
            out = torch.relu(self._bn(out, name + ".norm2"))
            out = oc.conv_1x1(out, self._w(name + ".conv3.kernel"))
            out = self._bn(out, name + ".norm3")
        res = x
        if (self.prefix + name + ".downsample.0.kernel") in self.p:
            res = self._bn(oc.conv_1x1(x, self._w(name + ".downsample.0.kernel")), name + ".downsample.1")
        return torch.relu(out + res)

    def _stage(self, x, name, n_blocks, nbr3):
        for i in range(n_blocks):
            x = self._block(x, f"{name}.{i}", nbr3)
        return x

    # -- network ---------------------------------------------------------------------------
    def features(self, coords: np.ndarray, feats: torch.Tensor, levels: CoordLevels | None = None):
        """Backbone up to block8 (``forward_no_logits``, ref models/minkunet.py:230-309).
        Returns ([N, PLANES[7]*expansion] features, bottleneck features at stride 16, levels)."""
        lv = levels or CoordLevels(coords)
        out = oc.conv_table(feats, lv.subm(0, 5), self._w("conv0p1s1.kernel"))
        skips = [torch.relu(self._bn(out, "bn0"))]
        x = skips[0]
        for i in range(4):                                     # encoder: levels 1..4
            x = oc.conv_table(x, lv.down(i), self._w(DOWN_CONVS[i] + ".kernel"))
            x = torch.relu(self._bn(x, f"bn{i + 1}"))
            x = self._stage(x, f"block{i + 1}", self.layers[i], lv.subm(i + 1, 3))
            skips.append(x)
        bottleneck = x
        for i in range(4):                                     # decoder: back to levels 3..0
            lvl = 3 - i
            x = oc.conv_table(x, lv.up(lvl), self._w(UP_CONVS[i] + ".kernel"))
            x = torch.relu(self._bn(x, f"bntr{i + 4}"))
            x = torch.cat((x, skips[lvl]), 1)
            x = self._stage(x, f"block{i + 5}", self.layers[i + 4], lv.subm(lvl, 3))
        return x, bottleneck, lv

    def head(self, feats96: torch.Tensor, name: str = "final"):
        b = self.p.get(self.prefix + name + ".bias")
        return oc.conv_1x1(feats96, self._w(name + ".kernel"), b)

    def forward(self, coords, feats, levels=None):
        f, _, lv = self.features(coords, feats, levels)
        return self.head(f), f, lv

    def forward_dummy(self, feats96):
        """ref models/minkunet.py:312-322: [final | max(final2)]."""
        y2 = self.head(feats96, "final2").max(dim=1, keepdim=True)[0]
        return torch.cat([self.head(feats96), y2], 1)

    def forward_novel(self, feats96):
        """ref models/minkunet.py:349-362: [final | final3 | max(final2)]."""
        y2 = self.head(feats96, "final2").max(dim=1, keepdim=True)[0]
        return torch.cat([self.head(feats96), self.head(feats96, "final3"), y2], 1)


def random_params(arch: str = "MinkUNet34C", in_channels: int = 1, n_classes: int = 17, seed: int = 1234, dtype=torch.float32) -> dict:
    """A random-init ``state_dict`` of ``arch`` under the reference's parameter names, built from the topology alone
    (ref models/minkunet.py:59-132, models/resnet.py:90-122) so that the CPU baseline never touches the product
    package.  Kernels are N(0, sqrt(2 / (kernel_volume * Cout)))-distributed (ME.utils.kaiming_normal_ with
    mode='fan_out', ref models/resnet.py:62-69), BN weight 1 / bias 0; float tensors other than running statistics
    require grad."""
    block, layers, planes, init_dim = ARCHS[arch]
    expansion = 4 if block == "bottleneck" else 1
    g = torch.Generator().manual_seed(seed)
    p = {}

    def conv(name, kv, cin, cout, bias=False):
        shape = (cin, cout) if kv == 1 else (kv, cin, cout)
        p[name + ".kernel"] = (torch.randn(shape, generator=g, dtype=dtype) * (2.0 / (kv * cout)) ** 0.5).requires_grad_(True)
        if bias:
            p[name + ".bias"] = torch.zeros((1, cout), dtype=dtype).requires_grad_(True)

    def bn(name, c):
        p[name + ".bn.weight"] = torch.ones(c, dtype=dtype).requires_grad_(True)
        p[name + ".bn.bias"] = torch.zeros(c, dtype=dtype).requires_grad_(True)
        p[name + ".bn.running_mean"] = torch.zeros(c, dtype=dtype)
        p[name + ".bn.running_var"] = torch.ones(c, dtype=dtype)
        p[name + ".bn.num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def stage(name, inplanes, width, n_blocks):
        for i in range(n_blocks):
            b = f"{name}.{i}"
            cin = inplanes if i == 0 else width * expansion
            if block == "basic":
                conv(b + ".conv1", 27, cin, width); bn(b + ".norm1", width)
                conv(b + ".conv2", 27, width, width); bn(b + ".norm2", width)
            else:
                conv(b + ".conv1", 1, cin, width); bn(b + ".norm1", width)
                conv(b + ".conv2", 27, width, width); bn(b + ".norm2", width)
                conv(b + ".conv3", 1, width, width * expansion); bn(b + ".norm3", width * expansion)
            if i == 0 and cin != width * expansion:
                conv(b + ".downsample.0", 1, cin, width * expansion); bn(b + ".downsample.1", width * expansion)
        return width * expansion

    conv("conv0p1s1", 125, in_channels, init_dim); bn("bn0", init_dim)
    inplanes, skip_width = init_dim, [init_dim]
    for i in range(4):
        conv(DOWN_CONVS[i], 8, inplanes, inplanes); bn(f"bn{i + 1}", inplanes)
        inplanes = stage(f"block{i + 1}", inplanes, planes[i], layers[i])
        skip_width.append(inplanes)
    for i in range(4):
        conv(UP_CONVS[i], 8, inplanes, planes[i + 4]); bn(f"bntr{i + 4}", planes[i + 4])
        inplanes = stage(f"block{i + 5}", planes[i + 4] + skip_width[3 - i], planes[i + 4], layers[i + 4])
    conv("final", 1, inplanes, n_classes, bias=True)
    return p
