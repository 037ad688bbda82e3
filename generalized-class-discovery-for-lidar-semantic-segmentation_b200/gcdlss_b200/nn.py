"""torch.nn.Modules with MinkowskiEngine-compatible names, constructors and state_dict keys
(SURVEY 8(b)), running on the sm_100a kernels.  Re-exported under ME names by the
``MinkowskiEngine`` shim package.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .config import get_math_mode
from .functional import (BasicBlockFunction, BatchNormActFunction, ConvBnActFunction, FusedBlockFunction, Im2colFunction, ReLUFunction,
                         SparseConvFunction, TrunkFunction, TrunkPlan, _BnSpec, fused_block_ok, packed_weights, shortcut_pair)
from .sparse_tensor import CoordinateMapKey, SparseTensor


def _scalar(v, name):
    if isinstance(v, (list, tuple)):
        if len(set(v)) != 1:
            raise NotImplementedError(f"anisotropic {name} is not on the MinkUNet path")
        v = v[0]
    return int(v)


class _ConvBase(nn.Module):
    TRANSPOSED = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        if dimension is None:
            raise ValueError("dimension must be given (3 for LiDAR voxels)")
        if dimension != 3:
            raise NotImplementedError("only 3-D sparse tensors are supported")
        if kernel_generator is not None or expand_coordinates:
            raise NotImplementedError("custom kernel generators / expand_coordinates are not on the MinkUNet path")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _scalar(kernel_size, "kernel_size")
        self.stride = _scalar(stride, "stride")
        self.dilation = _scalar(dilation, "dilation")
        if self.dilation != 1:
            raise NotImplementedError("dilation != 1 is not on the MinkUNet path")
        self.dimension = dimension
        self.kernel_volume = self.kernel_size ** dimension
        # ME layout: [K^3, Cin, Cout], offsets x-fastest; a 2-D [Cin, Cout] matrix when K^3 == 1
        shape = (self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape, dtype=torch.float32))
        self.bias = nn.Parameter(torch.empty(1, out_channels, dtype=torch.float32)) if bias else None
        self.reset_parameters()
        # persistent bf16 operand images for the tcgen05 path (see functional.PackedWeights)
        self._pk_fwd = self._pk_bwd = None
        self._pk_version, self._pk_ptr = -1, 0
        self._pk_tc = ops.tc_supported(in_channels, out_channels, self.kernel_volume)
        self._pk_mirror = self.stride == 1 and self.kernel_volume > 1     # stride-1 maps are self-transposed with mirrored offsets
        packed_weights.register(self)

    def reset_parameters(self):
        # ME default: U(-1/sqrt(fan), 1/sqrt(fan)), fan = (Cout if transposed else Cin) * K^3
        n = (self.out_channels if self.TRANSPOSED else self.in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def _out_stride(self, ts_in: int) -> int:
        if not self.TRANSPOSED:
            return ts_in * self.stride
        if ts_in % self.stride:
            raise ValueError("transposed convolution stride does not divide the tensor stride")
        return ts_in // self.stride

    def _out_dtype(self, feats: torch.Tensor, c_in: int, kv: int):
        if get_math_mode() == "bf16" and ops.tc_supported(c_in, self.out_channels, kv):
            return torch.bfloat16
        return torch.float32

    def forward(self, input: SparseTensor, coordinates=None) -> SparseTensor:
        if coordinates is not None:
            raise NotImplementedError("explicit output coordinates are not on the MinkUNet path")
        if not isinstance(input, SparseTensor):
            raise TypeError("expected a SparseTensor")
        mgr = input.coordinate_manager
        ts_in = input.tensor_stride_int
        ts_out = self._out_stride(ts_in)
        feats = input._F
        if feats.shape[1] != self.in_channels:
            raise ValueError(f"input has {feats.shape[1]} channels, layer expects {self.in_channels}")
        kmap = mgr.kernel_map(ts_in, self.kernel_size, self.stride, self.TRANSPOSED)
        bf16 = get_math_mode() == "bf16"
        thin = self.kernel_volume > 27 and self.kernel_volume * self.in_channels <= 256
        if thin:
            # 5x5x5 stem on a 1-channel input: explicit im2col, then a dense product (identity map)
            k_real = self.kernel_volume * self.in_channels
            k_pad = (k_real + 63) // 64 * 64
            col_dtype = torch.bfloat16 if (bf16 and ops.tc_supported(k_pad, self.out_channels, 1)) else torch.float32
            col = Im2colFunction.apply(feats, kmap, k_pad, col_dtype)
            w2 = self.kernel.reshape(k_real, self.out_channels)
            if k_pad != k_real:
                w2 = torch.cat([w2, w2.new_zeros(k_pad - k_real, self.out_channels)], 0)
            ident = mgr.kernel_map(ts_out, 1, 1, False)
            out = SparseConvFunction.apply(col, w2, self.bias, ident, self._out_dtype(col, k_pad, 1))
        elif (bf16 and self.kernel_volume == 1 and self.out_channels % 16 != 0 and feats.dtype == torch.bfloat16 and
              ops.tc_supported(self.in_channels, (self.out_channels + 15) // 16 * 16, 1)):
            # classifier head (ref models/minkunet.py:123-128, Cout = number of classes): zero-pad the output channels to
            # the next multiple of 16 so forward, dgrad and wgrad run on the tensor cores; fp32 logits, padding sliced off
            pad = (self.out_channels + 15) // 16 * 16 - self.out_channels
            w_pad = torch.nn.functional.pad(self.kernel, (0, pad))
            b_pad = torch.nn.functional.pad(self.bias, (0, pad)) if self.bias is not None else None
            out = SparseConvFunction.apply(feats, w_pad, b_pad, kmap, torch.float32)[:, :self.out_channels]
        else:
            out_dtype = self._out_dtype(feats, self.in_channels, self.kernel_volume)
            if bf16 and out_dtype == torch.bfloat16 and feats.dtype != torch.bfloat16:
                feats = feats.to(torch.bfloat16)
            out = SparseConvFunction.apply(feats, self.kernel, self.bias, kmap, out_dtype, self)
        return SparseTensor(out, coordinate_map_key=CoordinateMapKey(ts_out), coordinate_manager=mgr)

    def extra_repr(self):
        return (f"in={self.in_channels}, out={self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"dilation={self.dilation}, bias={self.bias is not None}")


class MinkowskiConvolution(_ConvBase):
    """ref use: models/minkunet.py:62-63, 67-68, 123-128; resnet_block convs."""
    TRANSPOSED = False


class MinkowskiConvolutionTranspose(_ConvBase):
    """ref use: models/minkunet.py:94-95, 101-102, 108-109, 115-116 (K=2, stride 2)."""
    TRANSPOSED = True


def _flush_batches_hook(module, prefix, keep_vars):
    module._flush_batches()


class MinkowskiBatchNorm(nn.Module):
    """``self.bn = nn.BatchNorm1d`` over all rows of F (state_dict keys ``<name>.bn.weight`` ...)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        if not (affine and track_running_stats):
            raise NotImplementedError("only affine batch norm with running statistics is on the MinkUNet path")
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)
        # num_batches_tracked is only read by checkpoints (and momentum=None): count on the host and write the
        # device scalar when somebody looks, instead of one tiny kernel per layer per step.
        self._pending_batches = 0
        self.register_state_dict_pre_hook(_flush_batches_hook)
        # a loaded checkpoint replaces the counter: batches counted before the load must not be added onto it
        self._register_load_state_dict_pre_hook(self._drop_pending_batches)

    def _drop_pending_batches(self, *_args):
        self._pending_batches = 0

    def _flush_batches(self):
        if self._pending_batches:
            self.bn.num_batches_tracked.add_(self._pending_batches)
            self._pending_batches = 0

    def fused_spec(self) -> _BnSpec:
        """Bookkeeping of one forward call (batch counter, momentum) for the fused block Functions."""
        bn = self.bn
        momentum = bn.momentum
        if bn.training:
            self._pending_batches += 1
            if momentum is None:
                self._flush_batches()
                momentum = 1.0 / float(bn.num_batches_tracked.item())
        return _BnSpec(self, float(momentum or 0.0))

    def forward(self, input: SparseTensor, relu: bool = False, residual: SparseTensor = None) -> SparseTensor:
        bn = self.bn
        training = bn.training
        momentum = bn.momentum
        if training:
            self._pending_batches += 1
            if momentum is None:
                self._flush_batches()
                momentum = 1.0 / float(bn.num_batches_tracked.item())
        x = input._F
        res = None
        if residual is not None:
            input._check_same_map(residual)
            res = residual._F if residual._F.dtype == x.dtype else residual._F.to(x.dtype)
        y = BatchNormActFunction.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, float(momentum or 0.0), bn.eps,
                                       relu, res)
        return input._like(y)


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, input: SparseTensor) -> SparseTensor:
        return input._like(ReLUFunction.apply(input._F))


class MinkowskiDropout(nn.Module):
    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p = p

    def forward(self, input: SparseTensor) -> SparseTensor:
        return input._like(torch.nn.functional.dropout(input._F, self.p, self.training))


class MinkowskiLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, input: SparseTensor) -> SparseTensor:
        return input._like(self.linear(input.F))


def conv_bn_act(conv, bn, x: SparseTensor, relu: bool = True) -> SparseTensor:
    """``relu(bn(conv(x)))`` as one autograd node when the layer allows it (no bias, not the im2col stem),
    otherwise the three modules in sequence.  Used by the drop-in MinkUNet trunk."""
    thin = conv.kernel_volume > 27
    if conv.bias is not None or thin or not isinstance(bn, MinkowskiBatchNorm):
        return bn(conv(x), relu=relu)
    mgr, ts_in = x.coordinate_manager, x.tensor_stride_int
    ts_out = conv._out_stride(ts_in)
    kmap = mgr.kernel_map(ts_in, conv.kernel_size, conv.stride, conv.TRANSPOSED)
    feats = x._F
    out_dtype = conv._out_dtype(feats, conv.in_channels, conv.kernel_volume)
    if out_dtype == torch.bfloat16 and feats.dtype != torch.bfloat16:
        feats = feats.to(torch.bfloat16)
    if fused_block_ok(feats, out_dtype, bn.bn.training):
        out = FusedBlockFunction.apply(feats, conv.kernel, bn.bn.weight, bn.bn.bias, None, None, None, None, None, None, kmap, None,
                                       (conv, None, None), (bn.fused_spec(), None, None), relu)
    else:
        out = ConvBnActFunction.apply(feats, conv.kernel, bn.bn.weight, bn.bn.bias, kmap, conv, bn.fused_spec(), relu, out_dtype)
    return SparseTensor(out, coordinate_map_key=CoordinateMapKey(ts_out), coordinate_manager=mgr)


def _shortcut_ok(ds) -> bool:
    pair = shortcut_pair(ds)
    return (pair is not None and isinstance(pair[0], MinkowskiConvolution) and isinstance(pair[1], MinkowskiBatchNorm) and
            pair[0].kernel_volume == 1 and pair[0].stride == 1 and pair[0].bias is None)


def _block_fusable(blk) -> bool:
    return (isinstance(blk, BasicBlock) and blk.conv1.stride == 1 and blk.conv1.bias is None and blk.conv2.bias is None and
            (blk.downsample is None or _shortcut_ok(blk.downsample)))


def trunk_plan(encoder, decoder):
    """TrunkPlan of a U-Net made of (conv, bn, nn.Sequential of BasicBlocks) stages, or None when some layer is outside what
    gcd_run_ops sequences (Bottleneck blocks, biases, channel counts the tensor-core path cannot take)."""
    stages = [(conv, bn, list(blocks)) for conv, bn, blocks in encoder + decoder]
    for conv, bn, blocks in stages:
        if conv.bias is not None or not isinstance(bn, MinkowskiBatchNorm) or conv.kernel_volume != 8 or not blocks:
            return None
        if not all(_block_fusable(b) for b in blocks):
            return None
        convs = [conv] + [c for b in blocks for c in ((b.conv1, b.conv2) + ((shortcut_pair(b.downsample)[0],) if b.downsample is not None else ()))]
        if not all(c._pk_tc for c in convs):
            return None
    return TrunkPlan(stages[:len(encoder)], stages[len(encoder):])


def run_trunk(plan, x: SparseTensor):
    """The eight stage outputs (SparseTensors) of a planned trunk from the stem's output, as one autograd node; None when the
    fast path does not apply right now (evaluation mode, instrumentation, mixed dtypes)."""
    feats = x._F
    training = all(bn.bn.training for _, bn in plan.units)
    bf16 = get_math_mode() == "bf16"
    if not fused_block_ok(feats, torch.bfloat16 if bf16 else torch.float32, training) or x.tensor_stride_int != 1:
        return None
    mgr = x.coordinate_manager
    outs = TrunkFunction.apply(plan, mgr, feats, *plan.parameters())
    strides = (2, 4, 8, 16, 8, 4, 2, 1)
    return [SparseTensor(o, coordinate_map_key=CoordinateMapKey(ts), coordinate_manager=mgr) for o, ts in zip(outs, strides)]


def cat(*tensors) -> SparseTensor:
    """ME.cat: channel concatenation of tensors on the same coordinate map (ref minkunet.py:178,188,198,208)."""
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    first = tensors[0]
    for t in tensors[1:]:
        first._check_same_map(t)
    feats = [t._F for t in tensors]
    if len({f.dtype for f in feats}) > 1:
        feats = [f.float() for f in feats]
    return first._like(torch.cat(feats, 1))


def _unsupported(name):
    class _Missing(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"ME.{name} is outside the voxelise + MinkUNet hot path (SURVEY section 8: out of scope)")
    _Missing.__name__ = name
    return _Missing


MinkowskiInstanceNorm = _unsupported("MinkowskiInstanceNorm")
MinkowskiMaxPooling = _unsupported("MinkowskiMaxPooling")
MinkowskiGlobalMaxPooling = _unsupported("MinkowskiGlobalMaxPooling")
MinkowskiGELU = _unsupported("MinkowskiGELU")


# -------------------------------------------------------------------------------------------------
# MinkowskiEngine.modules.resnet_block equivalents (attribute names fixed by the state_dict).
class BasicBlock(nn.Module):
    """conv3-norm-relu-conv3-norm, + identity or downsample(x), relu; BN+ReLU(+residual) run fused."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride, dilation=dilation, dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=dilation, dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x: SparseTensor) -> SparseTensor:
        ds = self.downsample
        fusable = self.conv1.stride == 1 and self.conv1.bias is None and self.conv2.bias is None and (ds is None or _shortcut_ok(ds))
        if not fusable:
            shortcut = x if ds is None else ds(x)
            y = self.norm1(self.conv1(x), relu=True)
            return self.norm2(self.conv2(y), relu=True, residual=shortcut)
        # one autograd node for the whole block (same kernels, far less Python / autograd work per layer)
        mgr, ts = x.coordinate_manager, x.tensor_stride_int
        kmap3 = mgr.kernel_map(ts, 3, 1, False)
        kmap1 = mgr.kernel_map(ts, 1, 1, False)
        feats = x._F
        out_dtype = self.conv1._out_dtype(feats, self.conv1.in_channels, 27)
        if out_dtype == torch.bfloat16 and feats.dtype != torch.bfloat16:
            feats = feats.to(torch.bfloat16)
        if ds is None:
            wd = gd = bd = hd = bnd = ds_bn = None
        else:
            ds_conv, ds_bn = shortcut_pair(ds)
            wd, gd, bd, hd, bnd = ds_conv.kernel, ds_bn.bn.weight, ds_bn.bn.bias, ds_conv, ds_bn.fused_spec()
        if fused_block_ok(feats, out_dtype, self.norm1.bn.training and self.norm2.bn.training and (ds is None or ds_bn.bn.training)):
            out = FusedBlockFunction.apply(feats, self.conv1.kernel, self.norm1.bn.weight, self.norm1.bn.bias, self.conv2.kernel,
                                           self.norm2.bn.weight, self.norm2.bn.bias, wd, gd, bd, kmap3, kmap1, (self.conv1, self.conv2, hd),
                                           (self.norm1.fused_spec(), self.norm2.fused_spec(), bnd), True)
            return x._like(out)
        out = BasicBlockFunction.apply(feats, self.conv1.kernel, self.norm1.bn.weight, self.norm1.bn.bias, self.conv2.kernel,
                                       self.norm2.bn.weight, self.norm2.bn.bias, wd, gd, bd, kmap3, kmap1, (self.conv1, self.conv2, hd),
                                       (self.norm1.fused_spec(), self.norm2.fused_spec(), bnd), out_dtype)
        return x._like(out)


class Bottleneck(nn.Module):
    """1x1 - 3x3x3 - 1x1 (x4 channels) residual block."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=1, dimension=dimension)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=stride, dilation=dilation, dimension=dimension)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv3 = MinkowskiConvolution(planes, planes * self.expansion, kernel_size=1, dimension=dimension)
        self.norm3 = MinkowskiBatchNorm(planes * self.expansion, momentum=bn_momentum)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x: SparseTensor) -> SparseTensor:
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.norm1(self.conv1(x), relu=True)
        y = self.norm2(self.conv2(y), relu=True)
        return self.norm3(self.conv3(y), relu=True, residual=shortcut)


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """ME.utils.kaiming_normal_ (ref models/resnet.py:84): fan of a [K^3, Cin, Cout] kernel is
    fan_in = Cin*K^3, fan_out = Cout*K^3; a 2-D [Cin, Cout] kernel is treated like nn.Linear's
    weight (fan_in = size(1), fan_out = size(0)) [ME-upstream, SURVEY 8(a) a9]."""
    if tensor.dim() == 3:
        kv, cin, cout = tensor.shape
        fan_in, fan_out = cin * kv, cout * kv
    elif tensor.dim() == 2:
        fan_in, fan_out = tensor.size(1), tensor.size(0)
    else:
        raise ValueError("kernel must be 2-D or 3-D")
    fan = fan_in if mode == "fan_in" else fan_out
    gain = nn.init.calculate_gain(nonlinearity, a)
    std = gain / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)
