"""autograd.Function wrappers: sparse convolution, fused BN(+ReLU)(+residual), ReLU, devoxelise.

Forward and backward both run the hand-written kernels behind the C ABI; torch only supplies the
autograd graph, device memory and the stream.
"""
from __future__ import annotations

import weakref

import torch

from . import ops
from ._cabi import MATH_BF16_TC, MATH_FP32_SIMT
from .config import get_math_mode

_pack_cache = {}


def _packed(weight: torch.Tensor, transpose: bool, mirror: bool):
    """bf16 operand image of a kernel parameter, re-packed only when the parameter changed.

    Entries are validated by identity (weak reference) and version counter: a freed parameter's
    address can be reused by a new one, so neither ``data_ptr`` nor ``id`` alone is a safe key."""
    key = (id(weight), transpose, mirror)
    hit = _pack_cache.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2].device == weight.device:
        return hit[2]
    packed = ops.pack_weights(_as3d(weight.detach()), transpose, mirror)
    if len(_pack_cache) > 1024:
        for k in [k for k, v in _pack_cache.items() if v[0]() is None]:
            del _pack_cache[k]
    _pack_cache[key] = (weakref.ref(weight), weight._version, packed)
    return packed


def _as3d(weight: torch.Tensor) -> torch.Tensor:
    return weight if weight.dim() == 3 else weight.unsqueeze(0)


def _use_tc(feats: torch.Tensor, c_in: int, c_out: int, kv: int) -> bool:
    return get_math_mode() == "bf16" and feats.dtype == torch.bfloat16 and ops.tc_supported(c_in, c_out, kv)


class SparseConvFunction(torch.autograd.Function):
    """out[o] = sum_k feats[nbr[k,o]] @ W[k] (+ bias).  ``kmap`` is a coords.KernelMap."""

    @staticmethod
    def forward(ctx, feats, weight, bias, kmap, out_dtype):
        w3 = _as3d(weight.detach())
        kv, c_in, c_out = w3.shape
        tc = _use_tc(feats, c_in, c_out, kv)
        b = bias.detach().reshape(-1) if bias is not None else None
        out = ops.conv_forward(feats.detach(), kmap.nbr, w3, kmap.n_out, bias=b, out_dtype=out_dtype,
                               math_mode=MATH_BF16_TC if tc else MATH_FP32_SIMT,
                               w_packed=_packed(weight, False, False) if tc else None)
        ctx.kmap, ctx.has_bias, ctx.w_shape = kmap, bias is not None, weight.shape
        ctx.save_for_backward(feats, weight)
        return out

    @staticmethod
    def backward(ctx, gout):
        feats, weight = ctx.saved_tensors
        kmap = ctx.kmap
        w3 = _as3d(weight.detach())
        kv, c_in, c_out = w3.shape
        gout = gout.contiguous()
        gfeats = gw = gb = None
        if ctx.needs_input_grad[0]:
            g = gout if gout.dtype == feats.dtype else gout.to(feats.dtype)
            tc = _use_tc(g, c_out, c_in, kv)
            gfeats = ops.conv_forward(g, kmap.back_nbr, w3, kmap.n_in, transpose_w=True, mirror=kmap.back_mirror,
                                      out_dtype=feats.dtype, math_mode=MATH_BF16_TC if tc else MATH_FP32_SIMT,
                                      w_packed=_packed(weight, True, kmap.back_mirror) if tc else None)
        if ctx.needs_input_grad[1]:
            gw3 = torch.zeros(w3.shape, dtype=torch.float32, device=w3.device)
            gb = torch.zeros(c_out, dtype=torch.float32, device=w3.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
            g = gout
            f = feats.detach()
            tc = get_math_mode() == "bf16" and f.dtype == torch.bfloat16 and ops.tc_supported(c_in, c_out, kv) and c_out <= 256
            if tc and g.dtype != torch.bfloat16:
                g = g.to(torch.bfloat16)
            ops.conv_wgrad(f, g, kmap.pairs, kv, gw3, dbias=gb, math_mode=MATH_BF16_TC if tc else MATH_FP32_SIMT)
            gw = gw3.reshape(ctx.w_shape)
            if gb is not None:
                gb = gb.reshape(1, -1)
        elif ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gout.float().sum(0, keepdim=True)
        return gfeats, gw, gb, None, None


class Im2colFunction(torch.autograd.Function):
    """Explicit im2col of a thin input (the Cin=1 5x5x5 stem): [n_out, ld] = gather by the table."""

    @staticmethod
    def forward(ctx, feats, kmap, ld_out, out_dtype):
        ctx.kmap, ctx.c_in, ctx.in_dtype = kmap, feats.shape[1], feats.dtype
        return ops.im2col(feats.detach(), kmap.nbr, ld_out, out_dtype)

    @staticmethod
    def backward(ctx, gcol):
        # col2im: gin[i, c] = sum_k gcol[nbr_back[k, i], (kv-1-k)*c_in + c]; the stem input never needs it in
        # the reference (raw remission features), so it is expressed with the generic forward kernel.
        kmap, c_in = ctx.kmap, ctx.c_in
        kv = kmap.kv
        gcol = gcol.float().contiguous()
        eye = torch.eye(c_in, dtype=torch.float32, device=gcol.device)
        gin = torch.zeros((kmap.n_in, c_in), dtype=torch.float32, device=gcol.device)
        back = kmap.back_nbr
        for k in range(kv):
            src = gcol[:, (kv - 1 - k) * c_in:(kv - k) * c_in] if kmap.back_mirror else gcol[:, k * c_in:(k + 1) * c_in]
            gin += ops.conv_forward(src.contiguous(), back[k:k + 1].contiguous(), eye.unsqueeze(0), kmap.n_in)
        return gin.to(ctx.in_dtype), None, None, None


class BatchNormActFunction(torch.autograd.Function):
    """y = act(batch_norm(x) + residual) with nn.BatchNorm1d semantics over the rows of x."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, relu, residual):
        y, mean, invstd = ops.bn_forward(x.detach(), gamma.detach(), beta.detach(), running_mean, running_var, training, momentum, eps,
                                         relu, residual.detach() if residual is not None else None)
        ctx.relu, ctx.training, ctx.has_res = relu, training, residual is not None
        ctx.save_for_backward(x, y if relu else None, mean, invstd, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, invstd, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dx, dres, dgamma, dbeta = ops.bn_backward(dy, x.detach(), y, mean, invstd, gamma.detach(), ctx.relu, ctx.training,
                                                  ctx.has_res and ctx.needs_input_grad[9])
        return dx, dgamma, dbeta, None, None, None, None, None, None, dres


class ReLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.relu(x.detach())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.relu_backward(dy.to(y.dtype), y)


class DevoxelizeFunction(torch.autograd.Function):
    """Voxel -> point gather ``feats[inverse_map]``; backward is the segmented sum over each voxel's points."""

    @staticmethod
    def forward(ctx, voxel_feats, inverse_map):
        ctx.n_vox, ctx.dtype = voxel_feats.shape[0], voxel_feats.dtype
        ctx.save_for_backward(inverse_map)
        return ops.rows_gather(voxel_feats.detach(), inverse_map)

    @staticmethod
    def backward(ctx, gpoints):
        (inverse_map,) = ctx.saved_tensors
        seg_off, order = ops.csr_build(inverse_map, ctx.n_vox)
        g = ops.segment_reduce(gpoints.contiguous(), seg_off, order, ctx.n_vox, 0)
        return g.to(ctx.dtype), None


def devoxelize(voxel_feats: torch.Tensor, inverse_map: torch.Tensor) -> torch.Tensor:
    """Point features from voxel features (ref models/decoder.py:416-424, exp_merge_mean_teacher.py:2845)."""
    return DevoxelizeFunction.apply(voxel_feats, inverse_map.to(voxel_feats.device, torch.int64))


def voxelize_reduce(point_feats: torch.Tensor, inverse_map: torch.Tensor, n_voxels: int, mode: str = "mean") -> torch.Tensor:
    """Point -> voxel reduce (mmcv DynamicScatter of ref models/encoder.py:121-164); forward only."""
    seg_off, order = ops.csr_build(inverse_map.to(point_feats.device, torch.int64), n_voxels)
    return ops.segment_reduce(point_feats, seg_off, order, n_voxels, {"sum": 0, "mean": 1, "max": 2}[mode])
