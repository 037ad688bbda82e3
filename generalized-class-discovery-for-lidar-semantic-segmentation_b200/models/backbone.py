"""Drop-in for ``MinkUNetBackbone`` of the reference's ``models/backbone.py`` (:47-254) with
``sparseconv_backend='minkowski'`` -- the mmdet3d-style wrapper of the same U-Net.

mmdet3d / mmcv are not installable here, so the two building blocks the reference imports from
``mmdet3d.models.layers.minkowski_engine_block`` are restated from their published definitions (parity unpinned):
``MinkowskiConvModule`` = ``self.net = nn.Sequential(conv[, norm][, act])`` and ``MinkowskiBasicBlock`` = ME's BasicBlock with
``dimension=3``.  State-dict keys therefore read ``conv_input.0.net.0.kernel``, ``encoder.0.1.conv1.kernel``,
``encoder.0.1.downsample.net.0.kernel`` ... as under mmdet3d.  The torchsparse / spconv back-ends of the reference class are
other libraries' code paths and are out of scope."""
from typing import List

import torch.nn as nn
from torch import Tensor

import MinkowskiEngine as ME
from MinkowskiEngine.modules.resnet_block import BasicBlock, Bottleneck

from gcdlss_b200.nn import conv_bn_act


class MinkowskiConvModule(nn.Module):
    """conv -> MinkowskiBatchNorm -> MinkowskiReLU (``act=False``: no activation, the residual branch)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, bias=False, transposed=False, act=True, **_ignored):
        super().__init__()
        conv = (ME.MinkowskiConvolutionTranspose if transposed else ME.MinkowskiConvolution)(
            in_channels, out_channels, kernel_size=kernel_size, stride=stride, dilation=dilation, bias=bias, dimension=3)
        layers = [conv, ME.MinkowskiBatchNorm(out_channels)]
        if act:
            layers.append(ME.MinkowskiReLU(inplace=True))
        self.net = nn.Sequential(*layers)
        self.act = act

    def forward(self, x):
        return conv_bn_act(self.net[0], self.net[1], x, relu=self.act)       # one fused unit (conv, batch statistics, apply + ReLU)


class MinkUNetBackbone(nn.Module):
    def __init__(self, in_channels: int = 4, base_channels: int = 32, num_stages: int = 4,
                 encoder_channels: List[int] = [32, 64, 128, 256], encoder_blocks: List[int] = [2, 2, 2, 2],
                 decoder_channels: List[int] = [256, 128, 96, 96], decoder_blocks: List[int] = [2, 2, 2, 2],
                 block_type: str = 'basic', sparseconv_backend: str = 'minkowski', init_cfg=None) -> None:
        super().__init__()
        assert num_stages == len(encoder_channels) == len(decoder_channels)
        if sparseconv_backend != 'minkowski':
            raise NotImplementedError("gcdlss_b200 provides the 'minkowski' back-end of MinkUNetBackbone; torchsparse / spconv are other "
                                      "libraries' code paths (SURVEY section 2: out of scope)")
        block = BasicBlock if block_type == 'basic' else Bottleneck
        self.num_stages, self.sparseconv_backend = num_stages, sparseconv_backend
        self.conv_input = nn.Sequential(MinkowskiConvModule(in_channels, base_channels, kernel_size=3),
                                        MinkowskiConvModule(base_channels, base_channels, kernel_size=3))
        enc_ch = [base_channels] + list(encoder_channels)          # the reference inserts in place; copies keep the defaults intact
        dec_ch = [enc_ch[-1]] + list(decoder_channels)
        self.encoder, self.decoder = nn.ModuleList(), nn.ModuleList()
        for i in range(num_stages):
            layers = [MinkowskiConvModule(enc_ch[i], enc_ch[i], kernel_size=2, stride=2)]
            for j in range(encoder_blocks[i]):
                if j == 0 and enc_ch[i] != enc_ch[i + 1]:
                    layers.append(block(enc_ch[i], enc_ch[i + 1], downsample=MinkowskiConvModule(enc_ch[i], enc_ch[i + 1], kernel_size=1, act=False), dimension=3))
                else:
                    layers.append(block(enc_ch[i + 1], enc_ch[i + 1], dimension=3))
            self.encoder.append(nn.Sequential(*layers))
            cat_ch = dec_ch[i + 1] + enc_ch[-2 - i]
            dlayers = [MinkowskiConvModule(dec_ch[i], dec_ch[i + 1], kernel_size=2, stride=2, transposed=True)]
            for j in range(decoder_blocks[i]):
                if j == 0:
                    dlayers.append(block(cat_ch, dec_ch[i + 1], downsample=MinkowskiConvModule(cat_ch, dec_ch[i + 1], kernel_size=1, act=False), dimension=3))
                else:
                    dlayers.append(block(dec_ch[i + 1], dec_ch[i + 1], dimension=3))
            self.decoder.append(nn.ModuleList([dlayers[0], nn.Sequential(*dlayers[1:])]))

    def forward(self, voxel_features: Tensor, coors: Tensor) -> Tensor:
        """voxel_features [N, C], coors [N, 4] handed to ``ME.SparseTensor`` as they are (batch index first, as the
        'minkunet' voxelizer emits them); returns the last decoder stage's features [N, decoder_channels[-1]]
        (ref models/backbone.py:209-254)."""
        x = ME.SparseTensor(voxel_features, coors)
        x = self.conv_input(x)
        laterals = [x]
        for encoder_layer in self.encoder:
            x = encoder_layer(x)
            laterals.append(x)
        laterals = laterals[:-1][::-1]
        for i, decoder_layer in enumerate(self.decoder):
            x = decoder_layer[0](x)
            x = ME.cat(x, laterals[i])
            x = decoder_layer[1](x)
        return x.F
