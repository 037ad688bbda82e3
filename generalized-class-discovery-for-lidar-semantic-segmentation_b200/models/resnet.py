"""Drop-in for the reference ``models/resnet.py``: ``ResNetBase`` supplies stage construction
(``_make_layer``) and weight initialisation to the MinkUNet family (ref models/resnet.py:29-122)."""
import torch.nn as nn

import MinkowskiEngine as ME


class ResNetBase(nn.Module):
    BLOCK = None
    LAYERS = ()
    INIT_DIM = 64
    PLANES = (64, 128, 256, 512)

    def __init__(self, in_channels, out_channels, D=3):
        nn.Module.__init__(self)
        self.D = D
        assert self.BLOCK is not None, "subclass must choose a residual block"
        self.network_initialization(in_channels, out_channels, D)
        self.weight_initialization()

    def network_initialization(self, in_channels, out_channels, D):
        raise NotImplementedError("the plain ResNet classifier (instance norm / pooling) is outside the MinkUNet hot path; "
                                  "subclasses such as MinkUNetBase override network_initialization")

    def weight_initialization(self):
        # kaiming(fan_out) on every MinkowskiConvolution (transposed convs keep their default init:
        # the isinstance test in ref models/resnet.py:83 does not match them); BN weight 1, bias 0.
        for m in self.modules():
            if isinstance(m, ME.MinkowskiConvolution):
                ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1, bn_momentum=0.1):
        out_planes = planes * block.expansion
        shortcut = None
        if stride != 1 or self.inplanes != out_planes:
            shortcut = nn.Sequential(
                ME.MinkowskiConvolution(self.inplanes, out_planes, kernel_size=1, stride=stride, dimension=self.D),
                ME.MinkowskiBatchNorm(out_planes),
            )
        stage = [block(self.inplanes, planes, stride=stride, dilation=dilation, downsample=shortcut, dimension=self.D)]
        self.inplanes = out_planes
        stage += [block(self.inplanes, planes, stride=1, dilation=dilation, dimension=self.D) for _ in range(1, blocks)]
        return nn.Sequential(*stage)
