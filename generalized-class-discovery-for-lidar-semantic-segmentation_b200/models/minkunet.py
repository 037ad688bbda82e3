"""Drop-in for the reference ``models/minkunet.py``: the MinkUNet family on the sm_100a kernels.

Same class names, constructors ``(in_channels, out_channels, D=3)``, class attributes
(``BLOCK / LAYERS / PLANES / INIT_DIM``), forward signatures and ``state_dict`` keys as the
reference (ref models/minkunet.py:44-591).  The topology is generated from a stage table instead of
being spelled out layer by layer; BN+ReLU pairs run as one fused kernel.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import MinkowskiEngine as ME
from MinkowskiEngine.modules.resnet_block import BasicBlock, Bottleneck

from gcdlss_b200.nn import conv_bn_act, run_trunk, trunk_plan

from models.resnet import ResNetBase

# (conv attribute, bn attribute, block attribute) per resolution change; "p<stride>" in the names
# is the tensor stride the layer reads (ref models/minkunet.py:62-121).
_ENCODER = (("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"), ("conv3p4s2", "bn3", "block3"), ("conv4p8s2", "bn4", "block4"))
_DECODER = (("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"), ("convtr6p4s2", "bntr6", "block7"),
            ("convtr7p2s2", "bntr7", "block8"))


class NormedLinear(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.weight = nn.Parameter(torch.Tensor(in_features, out_features))
        self.weight.data.uniform_(-1, 1).renorm_(2, 1, 1e-5).mul_(1e5)

    def forward(self, x):
        return 10 * F.normalize(x.features, dim=1).mm(F.normalize(self.weight, dim=0))


class _UNetTrunk(ResNetBase):
    """Shared constructor and data flow of MinkUNetBase / MinkUNetBaseRC."""
    BLOCK = None
    DILATIONS = (1, 1, 1, 1, 1, 1, 1, 1)
    LAYERS = (2, 2, 2, 2, 2, 2, 2, 2)
    PLANES = (32, 64, 128, 256, 256, 128, 96, 96)
    INIT_DIM = 32
    OUT_TENSOR_STRIDE = 1
    WITH_DROPOUT = False

    def __init__(self, in_channels, out_channels, D=3):
        ResNetBase.__init__(self, in_channels, out_channels, D)

    def network_initialization(self, in_channels, out_channels, D):
        exp = self.BLOCK.expansion
        self.inplanes = self.INIT_DIM
        self.conv0p1s1 = ME.MinkowskiConvolution(in_channels, self.inplanes, kernel_size=5, dimension=D)
        self.bn0 = ME.MinkowskiBatchNorm(self.inplanes)
        skip_planes = [self.INIT_DIM]
        for i, (conv, bn, block) in enumerate(_ENCODER):
            setattr(self, conv, ME.MinkowskiConvolution(self.inplanes, self.inplanes, kernel_size=2, stride=2, dimension=D))
            setattr(self, bn, ME.MinkowskiBatchNorm(self.inplanes))
            setattr(self, block, self._make_layer(self.BLOCK, self.PLANES[i], self.LAYERS[i]))
            skip_planes.append(self.inplanes)
        for i, (conv, bn, block) in enumerate(_DECODER):
            planes = self.PLANES[4 + i]
            setattr(self, conv, ME.MinkowskiConvolutionTranspose(self.inplanes, planes, kernel_size=2, stride=2, dimension=D))
            setattr(self, bn, ME.MinkowskiBatchNorm(planes))
            self.inplanes = planes + skip_planes[3 - i]
            setattr(self, block, self._make_layer(self.BLOCK, planes, self.LAYERS[4 + i]))
        self.final = ME.MinkowskiConvolution(self.PLANES[7] * exp, out_channels, kernel_size=1, bias=True, dimension=D)
        self.relu = ME.MinkowskiReLU(inplace=True)
        if self.WITH_DROPOUT:
            self.dropout = ME.MinkowskiDropout(p=0.5)

    def _trunk(self, x):
        """Returns the outputs of block1..block8 (index 0 = block1)."""
        cur = conv_bn_act(self.conv0p1s1, self.bn0, x)
        # training fast path: everything after the stem as ONE autograd node / two library calls (gcd_run_ops)
        plan = self.__dict__.get("_trunk_plan", False)
        if plan is False:
            plan = trunk_plan([(getattr(self, c), getattr(self, b), getattr(self, k)) for c, b, k in _ENCODER],
                              [(getattr(self, c), getattr(self, b), getattr(self, k)) for c, b, k in _DECODER])
            self.__dict__["_trunk_plan"] = plan
        if plan is not None:
            stages = run_trunk(plan, cur)
            if stages is not None:
                return stages
        skips, stages = [cur], []
        for conv, bn, block in _ENCODER:
            cur = conv_bn_act(getattr(self, conv), getattr(self, bn), cur)
            cur = getattr(self, block)(cur)
            skips.append(cur)
            stages.append(cur)
        for i, (conv, bn, block) in enumerate(_DECODER):
            cur = conv_bn_act(getattr(self, conv), getattr(self, bn), cur)
            cur = getattr(self, block)(ME.cat(cur, skips[3 - i]))
            stages.append(cur)
        return stages


class MinkUNetBaseRC(_UNetTrunk):
    """ref models/minkunet.py:44-374 (the variant the two Lightning modules instantiate)."""
    WITH_DROPOUT = True

    def forward(self, x, is_seg=True, layers=[], use_last=False, use_both=False, is_also=False):
        assert isinstance(layers, list), 'layers should be a list.'
        if len(layers) == 0:
            layers = [4]
        stages = self._trunk(x)
        out, bottleneck = stages[7], stages[3]
        taps = tuple(stages[i - 1] for i in range(1, 9) if i in layers)
        if is_seg:
            if is_also:
                return self.final(out), self.final(out).F
            if use_both:
                return self.final(out), out, bottleneck
            if use_last:
                return self.final(out), out
            return (self.final(out),) + taps
        if use_both:
            return out, out, bottleneck
        if use_last:
            return out, out
        return (out,) + taps

    def forward_no_logits(self, x, layers=[]):
        assert isinstance(layers, list), 'layers should be a list.'
        return self._trunk(x)[7]

    def _heads(self, feat, names):
        """The 1x1 classifier heads ``names`` (final / final2 / final3, ref models/minkunet.py:312-362 and the heads the
        Lightning module bolts on, ref modules/exp_merge_mean_teacher.py:128-153) applied to the same 96-channel features
        as ONE 96 -> sum(C_i) product instead of one per head: one forward, one dgrad and one wgrad launch for all of them
        (the features are read once each way).  Returns the per-head logits (fp32 [N, C_i])."""
        heads = [getattr(self, n) for n in names]
        if len(heads) == 1 or any(h.bias is None for h in heads) or any(h.kernel_volume != 1 for h in heads):
            return [h(feat).F for h in heads]
        from gcdlss_b200.functional import SparseConvFunction
        from gcdlss_b200.config import get_math_mode
        from gcdlss_b200 import ops
        widths = [h.out_channels for h in heads]
        total = sum(widths)
        pad = (-total) % 16                       # the tensor-core path wants a multiple of 16 output channels
        w = torch.cat([h.kernel for h in heads] + ([heads[0].kernel.new_zeros(heads[0].in_channels, pad)] if pad else []), 1)
        b = torch.cat([h.bias for h in heads] + ([heads[0].bias.new_zeros(1, pad)] if pad else []), 1)
        x = feat._F
        kmap = feat.coordinate_manager.kernel_map(feat.tensor_stride_int, 1, 1, False)
        if get_math_mode() == "bf16" and x.dtype != torch.bfloat16 and ops.tc_supported(x.shape[1], total + pad, 1):
            x = x.to(torch.bfloat16)
        out = SparseConvFunction.apply(x, w, b, kmap, torch.float32)
        return list(torch.split(out[:, :total], widths, dim=1))

    def _ncc_logits(self, feat, reduce):
        known, rc = self._heads(feat, ("final", "final2"))
        return torch.cat([known, reduce(rc)], dim=1)

    def forward_dummy(self, feat):
        return self._ncc_logits(feat, lambda t: torch.max(t, dim=1, keepdim=True)[0])

    def forward_dummy_mean(self, feat):
        return self._ncc_logits(feat, lambda t: torch.mean(t, dim=1, keepdim=True))

    def forward_dummy_sum(self, feat):
        return self._ncc_logits(feat, lambda t: torch.sum(t, dim=1, keepdim=True))

    def forward_novel(self, feat):
        known, novel, rc = self._heads(feat, ("final", "final3", "final2"))
        return torch.cat([known, novel, torch.max(rc, dim=1, keepdim=True)[0]], dim=1)

    def forward_dummy_sparse(self, x, is_seg=True):
        out = self.forward_no_logits(x)
        known = self.final(out)
        rc = torch.max(self.final2(out).F, dim=1, keepdim=True)[0]
        return ME.SparseTensor(torch.cat([known.F, rc], dim=1), coordinate_map_key=known.coordinate_map_key,
                               coordinate_manager=known.coordinate_manager)


class MinkUNetBase(_UNetTrunk):
    """ref models/minkunet.py:376-525."""

    def forward(self, x, return_feats=False):
        out = self._trunk(x)[7]
        return out if return_feats else self.final(out)


def _variant(name, base, **attrs):
    return type(name, (base,), dict(attrs, __module__=__name__))


MinkUNet14 = _variant("MinkUNet14", MinkUNetBase, BLOCK=BasicBlock, LAYERS=(1, 1, 1, 1, 1, 1, 1, 1))
MinkUNet18 = _variant("MinkUNet18", MinkUNetBase, BLOCK=BasicBlock, LAYERS=(2, 2, 2, 2, 2, 2, 2, 2))
MinkUNet34 = _variant("MinkUNet34", MinkUNetBase, BLOCK=BasicBlock, LAYERS=(2, 3, 4, 6, 2, 2, 2, 2))
MinkUNet50 = _variant("MinkUNet50", MinkUNetBase, BLOCK=Bottleneck, LAYERS=(2, 3, 4, 6, 2, 2, 2, 2))
MinkUNet101 = _variant("MinkUNet101", MinkUNetBase, BLOCK=Bottleneck, LAYERS=(2, 3, 4, 23, 2, 2, 2, 2))
MinkUNet14A = _variant("MinkUNet14A", MinkUNet14, PLANES=(32, 64, 128, 256, 128, 128, 96, 96))
MinkUNet14B = _variant("MinkUNet14B", MinkUNet14, PLANES=(32, 64, 128, 256, 128, 128, 128, 128))
MinkUNet14C = _variant("MinkUNet14C", MinkUNet14, PLANES=(32, 64, 128, 256, 192, 192, 128, 128))
MinkUNet14D = _variant("MinkUNet14D", MinkUNet14, PLANES=(32, 64, 128, 256, 384, 384, 384, 384))
MinkUNet18A = _variant("MinkUNet18A", MinkUNet18, PLANES=(32, 64, 128, 256, 256, 128, 96, 96))
MinkUNet18B = _variant("MinkUNet18B", MinkUNet18, PLANES=(32, 64, 128, 256, 128, 128, 128, 128))
MinkUNet18D = _variant("MinkUNet18D", MinkUNet18, PLANES=(32, 64, 128, 256, 384, 384, 384, 384))
MinkUNet34A = _variant("MinkUNet34A", MinkUNet34, PLANES=(32, 64, 128, 256, 256, 128, 64, 64))
MinkUNet34B = _variant("MinkUNet34B", MinkUNet34, PLANES=(32, 64, 128, 256, 256, 128, 64, 32))
MinkUNet34C = _variant("MinkUNet34C", MinkUNet34, PLANES=(32, 64, 128, 256, 256, 128, 96, 96))
MinkUNet34RC = _variant("MinkUNet34RC", MinkUNetBaseRC, BLOCK=BasicBlock, LAYERS=(2, 3, 4, 6, 2, 2, 2, 2))
