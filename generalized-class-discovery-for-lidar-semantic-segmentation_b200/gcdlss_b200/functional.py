"""autograd.Function wrappers: sparse convolution, fused BN(+ReLU)(+residual), ReLU, devoxelise.

Forward and backward both run the hand-written kernels behind the C ABI; torch only supplies the
autograd graph, device memory and the stream.
"""
from __future__ import annotations

import weakref

import torch

import ctypes as C

from . import ops
from ._cabi import MATH_BF16_TC, MATH_FP32_SIMT, BlockArgs, call
from .config import get_math_mode, nvtx_range

class PackedWeights:
    """bf16 tcgen05 operand images (forward and dgrad orientation) of every convolution module, kept in
    persistent device buffers and refreshed by ONE batched launch when parameters changed (i.e. once per
    optimiser / EMA step) instead of two small launches per layer."""

    def __init__(self):
        self.modules = weakref.WeakSet()
        self._sig = None
        self._table = None

    def register(self, module):
        self.modules.add(module)

    @staticmethod
    def _stale(m, dev):
        w = m.kernel
        return w.device == dev and (m._pk_version != w._version or m._pk_ptr != w.data_ptr())

    def ensure(self, module):
        w = module.kernel
        if module._pk_version == w._version and module._pk_ptr == w.data_ptr():
            return
        dev = w.device
        self.modules.add(module)            # deep-copied modules (EMA teacher) never ran __init__
        stale = [m for m in self.modules if m._pk_tc and self._stale(m, dev)]
        if module not in stale:
            stale.append(module)
        recs, blocks = [], 0
        for m in stale:
            kv, cin, cout = m.kernel_volume, m.in_channels, m.out_channels
            if m._pk_fwd is None or m._pk_fwd.device != dev:
                nbytes = int(ops.lib().gcd_conv_packed_weight_bytes(kv, cin, cout))
                m._pk_fwd = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                m._pk_bwd = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            for dst, transpose, kdim, ndim in ((m._pk_fwd, 0, cin, cout), (m._pk_bwd, 1, cout, cin)):
                recs.append((m.kernel.data_ptr(), dst.data_ptr(), kv, cin, cout, transpose, int(m._pk_mirror and transpose), blocks))
                blocks += (kv * ((kdim + 63) // 64) * ndim * 64 + 255) // 256
        sig = tuple(r[:2] for r in recs)
        if sig != self._sig or self._table is None or self._table.device != dev:
            import numpy as np
            host = np.zeros(len(recs), dtype=np.dtype([("w", "<u8"), ("dst", "<u8"), ("kv", "<i4"), ("c_in", "<i4"), ("c_out", "<i4"),
                                                         ("transpose", "<i4"), ("mirror", "<i4"), ("block_start", "<i4")]))
            for i, r in enumerate(recs):
                host[i] = r
            self._table = torch.from_numpy(host.view(np.uint8).copy()).to(dev)
            self._sig = sig
        ops.pack_weights_batched(self._table, len(recs), blocks)
        for m in stale:
            m._pk_version, m._pk_ptr = m.kernel._version, m.kernel.data_ptr()


packed_weights = PackedWeights()


def _as3d(weight: torch.Tensor) -> torch.Tensor:
    return weight if weight.dim() == 3 else weight.unsqueeze(0)


def _use_tc(feats: torch.Tensor, c_in: int, c_out: int, kv: int) -> bool:
    return get_math_mode() == "bf16" and feats.dtype == torch.bfloat16 and ops.tc_supported(c_in, c_out, kv)


# ---- convolution building blocks (no autograd): shared by the per-op and the fused block Functions ----------------
def conv_fwd_impl(feats, weight, bias, kmap, out_dtype, holder):
    w3 = _as3d(weight.detach())
    kv, c_in, c_out = w3.shape
    tc = _use_tc(feats, c_in, c_out, kv)
    packed = None
    if tc:
        if holder is not None:
            packed_weights.ensure(holder)
            packed = holder._pk_fwd
        else:
            packed = ops.pack_weights(w3, False, False)
    nbr, rows, masks = kmap.tc_table() if tc else (kmap.nbr, None, None)
    return ops.conv_forward(feats, nbr, w3, kmap.n_out, bias=bias.detach().reshape(-1) if bias is not None else None,
                            out_dtype=out_dtype, math_mode=MATH_BF16_TC if tc else MATH_FP32_SIMT, w_packed=packed, out_rows=rows,
                            tile_masks=masks)


def conv_dgrad_impl(gout, weight, kmap, in_dtype, holder):
    w3 = _as3d(weight.detach())
    kv, c_in, c_out = w3.shape
    g = gout if gout.dtype == in_dtype else gout.to(in_dtype)
    tc = _use_tc(g, c_out, c_in, kv)
    packed = None
    if tc:
        if holder is not None and holder._pk_mirror == kmap.back_mirror:
            packed_weights.ensure(holder)
            packed = holder._pk_bwd
        else:
            packed = ops.pack_weights(w3, True, kmap.back_mirror)
    back, rows, masks = kmap.tc_back_table() if tc else (kmap.back_nbr, None, None)
    return ops.conv_forward(g, back, w3, kmap.n_in, transpose_w=True, mirror=kmap.back_mirror, out_dtype=in_dtype,
                            math_mode=MATH_BF16_TC if tc else MATH_FP32_SIMT, w_packed=packed, out_rows=rows, tile_masks=masks)


def conv_wgrad_impl(feats, gout, weight, kmap, want_bias_grad=False):
    """Returns (grad_weight shaped like weight, grad_bias [1, Cout] or None)."""
    w3 = _as3d(weight.detach())
    kv, c_in, c_out = w3.shape
    gw3 = ops.zeros_f32.take(w3.numel(), w3.device).view(w3.shape)
    gb = ops.zeros_f32.take(c_out, w3.device) if want_bias_grad else None
    tc = get_math_mode() == "bf16" and feats.dtype == torch.bfloat16 and ops.tc_supported(c_in, c_out, kv) and c_out <= 256
    g = gout.to(torch.bfloat16) if (tc and gout.dtype != torch.bfloat16) else gout
    ops.conv_wgrad(feats, g, kmap.pairs, kv, gw3, dbias=gb, math_mode=MATH_BF16_TC if tc else MATH_FP32_SIMT)
    return gw3.reshape(weight.shape), (gb.reshape(1, -1) if gb is not None else None)


class SparseConvFunction(torch.autograd.Function):
    """out[o] = sum_k feats[nbr[k,o]] @ W[k] (+ bias).  ``kmap`` is a coords.KernelMap."""

    @staticmethod
    def forward(ctx, feats, weight, bias, kmap, out_dtype, holder=None):
        out = conv_fwd_impl(feats.detach(), weight, bias, kmap, out_dtype, holder)
        ctx.kmap, ctx.has_bias, ctx.holder = kmap, bias is not None, holder
        ctx.save_for_backward(feats, weight)
        return out

    @staticmethod
    def backward(ctx, gout):
        feats, weight = ctx.saved_tensors
        kmap = ctx.kmap
        gout = gout.contiguous()
        gfeats = gw = gb = None
        if ctx.needs_input_grad[0]:
            gfeats = conv_dgrad_impl(gout, weight, kmap, feats.dtype, ctx.holder)
        if ctx.needs_input_grad[1]:
            gw, gb = conv_wgrad_impl(feats.detach(), gout, weight, kmap, ctx.has_bias and ctx.needs_input_grad[2])
        elif ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gout.float().sum(0, keepdim=True)
        return gfeats, gw, gb, None, None, None


class _BnSpec:
    """The pieces of a MinkowskiBatchNorm a fused Function needs (parameters go through apply() separately)."""
    __slots__ = ("running_mean", "running_var", "training", "momentum", "eps")

    def __init__(self, bn_module, momentum):
        bn = bn_module.bn
        self.running_mean, self.running_var = bn.running_mean, bn.running_var
        self.training, self.momentum, self.eps = bn.training, momentum, bn.eps


class ConvBnActFunction(torch.autograd.Function):
    """conv -> batch norm -> optional ReLU as ONE autograd node (the conv->bn->relu triples of MinkUNet's trunk:
    ref models/minkunet.py:140-147 etc.).  Same kernels as the separate modules, a third of the Python/autograd work."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, kmap, holder, bn: _BnSpec, relu: bool, out_dtype):
        y = conv_fwd_impl(x.detach(), w, None, kmap, out_dtype, holder)
        a, mean, invstd = ops.bn_forward(y, gamma.detach(), beta.detach(), bn.running_mean, bn.running_var, bn.training, bn.momentum,
                                         bn.eps, relu, None)
        ctx.kmap, ctx.holder, ctx.relu, ctx.training = kmap, holder, relu, bn.training
        ctx.save_for_backward(x, w, gamma, y, a if relu else None, mean, invstd)
        return a

    @staticmethod
    def backward(ctx, ga):
        x, w, gamma, y, a, mean, invstd = ctx.saved_tensors
        ga = ga.contiguous()
        if ga.dtype != y.dtype:
            ga = ga.to(y.dtype)
        dy, _, dgamma, dbeta = ops.bn_backward(ga, y, a, mean, invstd, gamma.detach(), ctx.relu, ctx.training, False)
        dx = conv_dgrad_impl(dy, w, ctx.kmap, x.dtype, ctx.holder) if ctx.needs_input_grad[0] else None
        dw, _ = conv_wgrad_impl(x.detach(), dy, w, ctx.kmap)
        return dx, dw, dgamma, dbeta, None, None, None, None, None


class BasicBlockFunction(torch.autograd.Function):
    """A whole residual block (conv3-bn-relu-conv3-bn (+ 1x1 conv-bn shortcut) + add + relu) as one autograd node.

    Inputs: x, then (w1, g1, b1, w2, g2, b2) and optionally (wd, gd, bd) of the shortcut; non-tensor arguments carry the
    kernel maps, weight-image holders and batch-norm buffers."""

    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, wd, gd, bd, kmap3, kmap1, holders, bns, out_dtype):
        h1, h2, hd = holders
        bn1, bn2, bnd = bns
        xd = x.detach()
        y1 = conv_fwd_impl(xd, w1, None, kmap3, out_dtype, h1)
        a1, m1, i1 = ops.bn_forward(y1, g1.detach(), b1.detach(), bn1.running_mean, bn1.running_var, bn1.training, bn1.momentum, bn1.eps,
                                    True, None)
        y2 = conv_fwd_impl(a1, w2, None, kmap3, out_dtype, h2)
        if wd is not None:
            yd = conv_fwd_impl(xd, wd, None, kmap1, out_dtype, hd)
            res, md, idd = ops.bn_forward(yd, gd.detach(), bd.detach(), bnd.running_mean, bnd.running_var, bnd.training, bnd.momentum,
                                          bnd.eps, False, None)
        else:
            yd = md = idd = None
            res = xd if xd.dtype == y2.dtype else xd.to(y2.dtype)
        out, m2, i2 = ops.bn_forward(y2, g2.detach(), b2.detach(), bn2.running_mean, bn2.running_var, bn2.training, bn2.momentum, bn2.eps,
                                     True, res)
        ctx.kmap3, ctx.kmap1, ctx.holders, ctx.training = kmap3, kmap1, holders, bn1.training
        ctx.save_for_backward(x, w1, g1, w2, g2, wd, gd, y1, a1, y2, out, yd, m1, i1, m2, i2, md, idd)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, w1, g1, w2, g2, wd, gd, y1, a1, y2, out, yd, m1, i1, m2, i2, md, idd = ctx.saved_tensors
        h1, h2, hd = ctx.holders
        tr = ctx.training
        gout = gout.contiguous()
        if gout.dtype != out.dtype:
            gout = gout.to(out.dtype)
        dy2, dres, dg2, db2 = ops.bn_backward(gout, y2, out, m2, i2, g2.detach(), True, tr, True)
        da1 = conv_dgrad_impl(dy2, w2, ctx.kmap3, a1.dtype, h2)
        dw2, _ = conv_wgrad_impl(a1, dy2, w2, ctx.kmap3)
        dy1, _, dg1, db1 = ops.bn_backward(da1, y1, a1, m1, i1, g1.detach(), True, tr, False)
        need_dx = ctx.needs_input_grad[0]
        dx = conv_dgrad_impl(dy1, w1, ctx.kmap3, x.dtype, h1) if need_dx else None
        dw1, _ = conv_wgrad_impl(x.detach(), dy1, w1, ctx.kmap3)
        dwd = dgd = dbd = None
        if wd is not None:
            dyd, _, dgd, dbd = ops.bn_backward(dres, yd, None, md, idd, gd.detach(), False, tr, False)
            if need_dx:
                dx = dx + conv_dgrad_impl(dyd, wd, ctx.kmap1, x.dtype, hd)
            dwd, _ = conv_wgrad_impl(x.detach(), dyd, wd, ctx.kmap1)
        elif need_dx:
            dx = dx + (dres if dres.dtype == dx.dtype else dres.to(dx.dtype))
        return dx, dw1, dg1, db1, dw2, dg2, db2, dwd, dgd, dbd, None, None, None, None, None


# ---- whole blocks sequenced in C (gcd_block_forward / gcd_block_backward) ----------------------------------------
def fused_block_ok(x: torch.Tensor, out_dtype, training: bool) -> bool:
    """The C-sequenced block path: training mode, one activation dtype, no per-launch instrumentation."""
    return (training and _FUSED_C and x.dtype == out_dtype and not ops.kernel_timer.enabled and not ops.kernel_timer.capture)


_FUSED_C = __import__("os").environ.get("GCDLSS_FUSED_BLOCKS", "1") != "0"


def _fill_unit(u, w, gamma, beta, bn, kmap, holder, tc, with_grad):
    w3 = _as3d(w)
    kv, c_in, c_out = w3.shape
    nbr, rows, masks = kmap.tc_table() if tc else (kmap.nbr, None, None)
    u.nbr = nbr.data_ptr() if nbr is not None else None
    u.out_rows = rows.data_ptr() if rows is not None else None
    u.tile_masks = masks.data_ptr() if masks is not None else None
    u.n_in, u.n_out = kmap.n_in, kmap.n_out
    u.kv, u.c_in, u.c_out = kv, c_in, c_out
    u.w = w.data_ptr()
    if tc:
        if holder._pk_mirror != kmap.back_mirror:
            raise RuntimeError("packed dgrad image orientation does not match the kernel map")
        packed_weights.ensure(holder)
        u.w_packed_fwd = holder._pk_fwd.data_ptr()
    u.gamma, u.beta = gamma.data_ptr(), beta.data_ptr()
    u.running_mean, u.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
    u.eps, u.momentum = bn.eps, bn.momentum
    if with_grad:
        back, back_rows, back_masks = kmap.tc_back_table() if tc else (kmap.back_nbr, None, None)
        u.back_nbr = back.data_ptr() if back is not None else None
        u.back_out_rows = back_rows.data_ptr() if back_rows is not None else None
        u.back_tile_masks = back_masks.data_ptr() if back_masks is not None else None
        u.back_mirror = int(kmap.back_mirror)
        if tc:
            u.w_packed_bwd = holder._pk_bwd.data_ptr()
        pairs = kmap.pairs
        if pairs is not None:
            u.pair_in, u.pair_out, u.pair_off = pairs[0].data_ptr(), pairs[1].data_ptr(), pairs[2].data_ptr()
            u.n_pairs = pairs[0].shape[0]
    return kv * c_in * c_out


class FusedBlockFunction(torch.autograd.Function):
    """conv-bn(-relu) or a whole residual block, forward and backward each ONE call into the C library
    (gcd_block_forward / gcd_block_backward sequence the same kernels the per-op Functions launch).

    ``w2 is None``: single unit, output relu?(bn(conv(x))).  Otherwise the BasicBlock with identity (``wd is None``)
    or 1x1-conv + bn shortcut."""

    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, wd, gd, bd, kmap, kmap1, holders, bns, relu1):
        xd = ops._rowmajor(x.detach())
        dev, dt = xd.device, xd.dtype
        has_u2, has_ud = w2 is not None, wd is not None
        c_in, c = w1.shape[-2], w1.shape[-1]
        n_out = kmap.n_out
        with_grad = any(ctx.needs_input_grad)      # grad mode itself is off inside Function.forward
        units = 1 + int(has_u2) + int(has_ud)
        slab = torch.empty((2 * units, n_out, c), dtype=dt, device=dev)
        moments = torch.empty((2 * units, c), dtype=torch.float32, device=dev)
        stats = ops.zeros_f64.take(2 * units * c, dev)
        a = BlockArgs()
        a.has_u2, a.has_ud, a.relu1 = int(has_u2), int(has_ud), int(relu1)
        a.dtype = ops._dtype_code(xd)
        a.x, a.ld_x = xd.data_ptr(), ops._ld(xd)
        esz = slab.element_size() * n_out * c
        base, mbase, sbase = slab.data_ptr(), moments.data_ptr(), stats.data_ptr()
        h1, h2, hd = holders
        bn1, bn2, bnd = bns
        tc = dt == torch.bfloat16 and get_math_mode() == "bf16"
        sizes = [_fill_unit(a.u1, w1, g1, b1, bn1, kmap, h1, tc and h1._pk_tc, with_grad)]
        a.u1.stats, a.u1.mean, a.u1.invstd = sbase, mbase, mbase + 4 * c
        a.y1, a.a1 = base, base + esz
        out = slab[1]
        if has_u2:
            sizes.append(_fill_unit(a.u2, w2, g2, b2, bn2, kmap, h2, tc and h2._pk_tc, with_grad))
            a.u2.stats, a.u2.mean, a.u2.invstd = sbase + 16 * c, mbase + 8 * c, mbase + 12 * c
            a.y2, a.out = base + 2 * esz, base + 3 * esz
            out = slab[3]
            if has_ud:
                sizes.append(_fill_unit(a.ud, wd, gd, bd, bnd, kmap1, hd, tc and hd._pk_tc, with_grad))
                a.ud.stats, a.ud.mean, a.ud.invstd = sbase + 32 * c, mbase + 16 * c, mbase + 20 * c
                a.yd, a.rd = base + 4 * esz, base + 5 * esz
        call("gcd_block_forward", C.byref(a), ops._stream())
        ops._count(a.launches)
        ctx.args, ctx.sizes, ctx.c, ctx.c_in = a, sizes, c, c_in
        ctx.w_shapes = (w1.shape, w2.shape if has_u2 else None, wd.shape if has_ud else None)
        ctx.keep = (xd, slab, moments, kmap, kmap1, holders)      # everything the struct points into
        ctx.x_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, gout):
        a, c, c_in = ctx.args, ctx.c, ctx.c_in
        xd, slab = ctx.keep[0], ctx.keep[1]
        dev, dt = slab.device, slab.dtype
        gout = ops._rowmajor(gout)
        if gout.dtype != dt:
            gout = gout.to(dt)
        if gout.stride(0) != c and gout.shape[0] > 1:
            gout = gout.contiguous()
        has_u2, has_ud = bool(a.has_u2), bool(a.has_ud)
        n_out, n_in = a.u1.n_out, a.u1.n_in
        units = 1 + int(has_u2) + int(has_ud)
        need_dx = ctx.needs_input_grad[0]
        gslab = torch.empty((3 if has_u2 else 1, n_out, c), dtype=dt, device=dev)
        esz = gslab.element_size() * n_out * c
        gb = gslab.data_ptr()
        dx = torch.empty((n_in, c_in), dtype=dt, device=dev) if need_dx else None
        dxd = torch.empty((n_in, c_in), dtype=dt, device=dev) if (need_dx and has_ud) else None
        sums = ops.zeros_f64.take(2 * units * c, dev)
        sizes = ctx.sizes
        grads = ops.zeros_f32.take(sum(sizes) + 2 * units * c, dev)
        gp, sp = grads.data_ptr(), sums.data_ptr()
        a.gout = gout.data_ptr()
        a.need_dx = int(need_dx)
        a.dx = dx.data_ptr() if dx is not None else None
        a.dxd = dxd.data_ptr() if dxd is not None else None
        outs, off = [], 0
        for i, (u, size) in enumerate(zip((a.u1, a.u2, a.ud), sizes)):
            u.sums = sp + 16 * c * i
            u.dw, u.dgamma, u.dbeta = gp + 4 * off, gp + 4 * (off + size), gp + 4 * (off + size + c)
            outs.append((grads[off:off + size].view(ctx.w_shapes[i]), grads[off + size:off + size + c], grads[off + size + c:off + size + 2 * c]))
            off += size + 2 * c
        if has_u2:
            a.dy2, a.dres, a.da1 = gb, gb + esz, gb + 2 * esz
            a.dy1 = a.da1             # the first unit's BN backward runs in place on the dgrad of the second
            a.dyd = a.dres            # likewise the shortcut's on the residual gradient
        else:
            a.dy1 = gb
        call("gcd_block_backward", C.byref(a), ops._stream())
        ops._count(a.launches)
        if dx is not None and dx.dtype != ctx.x_dtype:
            dx = dx.to(ctx.x_dtype)
        none3 = (None, None, None)
        g1 = outs[0]
        g2 = outs[1] if has_u2 else none3
        g3 = outs[2] if has_ud else none3
        return (dx,) + g1 + g2 + g3 + (None, None, None, None, None)


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


# ---- the whole U-Net trunk as ONE autograd node (gcd_run_ops) ------------------------------------------------------
def shortcut_pair(ds):
    """(1x1 convolution, batch norm) of a residual block's ``downsample`` module: ``nn.Sequential(conv, bn)`` as ResNetBase
    builds it (ref models/resnet.py:100-113), or a conv module that carries that Sequential as ``.net`` (mmdet3d's
    MinkowskiConvModule without activation, ref models/backbone.py:160-166).  None when it is anything else."""
    seq = getattr(ds, "net", ds)
    if isinstance(seq, torch.nn.Sequential) and len(seq) == 2:
        return seq[0], seq[1]
    return None


class TrunkPlan:
    """Static description of a MinkUNet trunk after the stem (ref models/minkunet.py:149-217): four encoder stages
    (stride-2 conv-bn-relu + BasicBlocks) and four decoder stages (transposed conv-bn-relu, concatenation with the
    encoder's output of the same resolution, BasicBlocks).  Built once per model from its modules; per step only
    pointers and row counts change."""

    def __init__(self, encoder, decoder):
        # encoder / decoder: lists of (conv, bn, [BasicBlock, ...]) in forward order
        self.encoder, self.decoder = encoder, decoder
        self.units = []          # (conv module, MinkowskiBatchNorm module) of every conv-bn unit, in parameter order
        for conv, bn, blocks in encoder + decoder:
            self.units.append((conv, bn))
            for blk in blocks:
                self.units.append((blk.conv1, blk.norm1))
                self.units.append((blk.conv2, blk.norm2))
                if blk.downsample is not None:
                    self.units.append(shortcut_pair(blk.downsample))
        self.n_blocks = sum(1 + len(b) for _, _, b in encoder + decoder)
        self.sum_c = sum(conv.out_channels for conv, _ in self.units)
        self.sum_w = sum(conv.kernel.numel() for conv, _ in self.units)

    def parameters(self):
        out = []
        for conv, bn in self.units:
            out += [conv.kernel, bn.bn.weight, bn.bn.bias]
        return out


class _Bump:
    """Bump allocator over one device buffer (all blocks of a pass share one allocation instead of ~5 tensors each)."""

    def __init__(self, nbytes, device, zero=False):
        self.buf = (torch.zeros if zero else torch.empty)(max(nbytes, 256), dtype=torch.uint8, device=device)
        self.base, self.off, self.cap = self.buf.data_ptr(), 0, max(nbytes, 256)

    def take(self, nbytes):
        off = self.off
        self.off = (off + nbytes + 255) & ~255
        if self.off > self.cap:
            raise RuntimeError("internal error: trunk arena too small")
        return self.base + off, off

    def view(self, off, n, c, dtype):
        return self.buf[off:off + n * c * dtype.itemsize].view(dtype).view(n, c)


def _al(nbytes):
    return (nbytes + 255) & ~255


class TrunkFunction(torch.autograd.Function):
    """Everything between the stem and the classifier head, forward and backward each ONE call into the library
    (gcd_run_ops over the gcd_block_* sequences of every stage plus the column copies that stand for ME.cat).
    Inputs: the stem's activation and the (kernel, gamma, beta) of every unit in ``plan.units`` order; outputs: the eight
    stage outputs (block1 .. block8) the reference's forward variants hand out (ref models/minkunet.py:134-228)."""

    @staticmethod
    def forward(ctx, plan, mgr, x, *params):
        from ._cabi import OP_BLOCK_FORWARD, OP_COPY_COLS, Op
        xd = ops._rowmajor(x.detach())
        dev, dt = xd.device, xd.dtype
        esz = xd.element_size()
        tc = dt == torch.bfloat16 and get_math_mode() == "bf16"
        with_grad = any(ctx.needs_input_grad)
        n_lv = [mgr.get_map(1 << l).n for l in range(5)]
        # ---- arena sizes: every unit writes y and a ([n_out, c] each), every decoder stage one concatenated input
        act_bytes, lvl = 0, 0
        for conv, bn, blocks in plan.encoder:
            lvl += 1
            act_bytes += 2 * _al(n_lv[lvl] * conv.out_channels * esz)
            for blk in blocks:
                act_bytes += (6 if blk.downsample is not None else 4) * _al(n_lv[lvl] * blk.conv1.out_channels * esz)
        for conv, bn, blocks in plan.decoder:
            lvl -= 1
            act_bytes += 2 * _al(n_lv[lvl] * conv.out_channels * esz) + _al(n_lv[lvl] * blocks[0].conv1.in_channels * esz)
            for blk in blocks:
                act_bytes += (6 if blk.downsample is not None else 4) * _al(n_lv[lvl] * blk.conv1.out_channels * esz)
        act = _Bump(act_bytes, dev)
        moments = torch.empty(2 * plan.sum_c, dtype=torch.float32, device=dev)
        stats = ops.zeros_f64.take(2 * plan.sum_c, dev)
        mptr, sptr = moments.data_ptr(), stats.data_ptr()
        dtc = ops._dtype_code(xd)
        blocks_c = (BlockArgs * plan.n_blocks)()
        prog = (Op * (plan.n_blocks + 8))()
        n_ops = bi = ci = pi = 0             # ops, blocks, channel offset into moments / stats, parameter index
        sizes = []                            # weight elements per unit, in plan.units order

        def unit(u, conv, bn, kmap, c_off):
            w, g, b = params[pi_box[0]], params[pi_box[0] + 1], params[pi_box[0] + 2]
            pi_box[0] += 3
            sizes.append(_fill_unit(u, w, g, b, bn.fused_spec(), kmap, conv, tc and conv._pk_tc, with_grad))
            c = conv.out_channels
            u.stats, u.mean, u.invstd = sptr + 16 * c_off, mptr + 8 * c_off, mptr + 8 * c_off + 4 * c
            return c

        pi_box = [0]

        def single(conv, bn, kmap, x_ptr, ld_x, n_out):
            nonlocal bi, ci, n_ops
            a = blocks_c[bi]
            a.has_u2 = a.has_ud = 0
            a.relu1, a.dtype, a.x, a.ld_x = 1, dtc, x_ptr, ld_x
            c = unit(a.u1, conv, bn, kmap, ci)
            ci += c
            nb = n_out * c * esz
            a.y1, _ = act.take(nb)
            a.a1, off = act.take(nb)
            prog[n_ops].op, prog[n_ops].block = OP_BLOCK_FORWARD, C.pointer(a)
            n_ops += 1
            bi += 1
            return a.a1, off, c

        def residual(blk, kmap3, kmap1, x_ptr, ld_x, n):
            nonlocal bi, ci, n_ops
            a = blocks_c[bi]
            ds = blk.downsample
            a.has_u2, a.has_ud, a.relu1, a.dtype, a.x, a.ld_x = 1, int(ds is not None), 1, dtc, x_ptr, ld_x
            c = unit(a.u1, blk.conv1, blk.norm1, kmap3, ci)
            unit(a.u2, blk.conv2, blk.norm2, kmap3, ci + c)
            ci += 2 * c
            nb = n * c * esz
            a.y1, _ = act.take(nb)
            a.a1, _ = act.take(nb)
            a.y2, _ = act.take(nb)
            a.out, off = act.take(nb)
            if ds is not None:
                unit(a.ud, *shortcut_pair(ds), kmap1, ci)
                ci += c
                a.yd, _ = act.take(nb)
                a.rd, _ = act.take(nb)
            prog[n_ops].op, prog[n_ops].block = OP_BLOCK_FORWARD, C.pointer(a)
            n_ops += 1
            bi += 1
            return a.out, off, c

        def copy_cols(dst, ld_dst, src, ld_src, n, c):
            nonlocal n_ops
            o = prog[n_ops]
            o.op, o.dtype, o.dst, o.ld_dst, o.src, o.ld_src, o.n, o.c = OP_COPY_COLS, dtc, dst, ld_dst, src, ld_src, n, c
            n_ops += 1

        cur, ld, c_cur = xd.data_ptr(), ops._ld(xd), xd.shape[1]
        skips = [(cur, ld, c_cur)]
        outs = []                                                     # (offset, n, c) of block1 .. block8 outputs
        for i, (conv, bn, blks) in enumerate(plan.encoder):
            ts = 1 << i
            n = n_lv[i + 1]
            cur, off, c_cur = single(conv, bn, mgr.kernel_map(ts, 2, 2, False), cur, ld, n)
            ld = c_cur
            k3, k1 = mgr.kernel_map(2 * ts, 3, 1, False), mgr.kernel_map(2 * ts, 1, 1, False)
            for blk in blks:
                cur, off, c_cur = residual(blk, k3, k1, cur, ld, n)
                ld = c_cur
            outs.append((off, n, c_cur))
            skips.append((cur, ld, c_cur))
        for i, (conv, bn, blks) in enumerate(plan.decoder):
            ts = 16 >> i
            n = n_lv[3 - i]
            up, _, c_up = single(conv, bn, mgr.kernel_map(ts, 2, 2, True), cur, ld, n)
            s_ptr, s_ld, s_c = skips[3 - i]
            c_cat = c_up + s_c
            cat_ptr, _ = act.take(n * c_cat * esz)
            copy_cols(cat_ptr, c_cat, up, c_up, n, c_up)                  # ME.cat(up, skip), ref models/minkunet.py:178-208
            copy_cols(cat_ptr + c_up * esz, c_cat, s_ptr, s_ld, n, s_c)
            cur, ld = cat_ptr, c_cat
            k3, k1 = mgr.kernel_map(ts // 2, 3, 1, False), mgr.kernel_map(ts // 2, 1, 1, False)
            for blk in blks:
                cur, off, c_cur = residual(blk, k3, k1, cur, ld, n)
                ld = c_cur
            outs.append((off, n, c_cur))
        with nvtx_range("gcd:trunk forward"):
            ops.run_ops(prog, n_ops, backward=False)
        ctx.plan, ctx.blocks_c, ctx.sizes, ctx.n_lv, ctx.x_dtype, ctx.x_shape = plan, blocks_c, sizes, n_lv, x.dtype, tuple(xd.shape)
        # everything the structs point into: input, arena, statistics, the kernel maps (tables, pair lists) of every level
        ctx.keep = (xd, act, moments, [mgr.kernel_map(*k) for k in list(mgr._kmaps)])
        ctx.set_materialize_grads(False)
        return tuple(act.view(off, n, c, dt) for off, n, c in outs)

    @staticmethod
    def backward(ctx, *gouts):
        from ._cabi import OP_ADD_COLS, OP_BLOCK_BACKWARD, OP_RECORD_EVENT, Op
        plan, blocks_c, sizes, n_lv = ctx.plan, ctx.blocks_c, ctx.sizes, ctx.n_lv
        xd, act = ctx.keep[0], ctx.keep[1]
        dev, dt = xd.device, xd.dtype
        esz = xd.element_size()
        dtc = ops._dtype_code(xd)
        need_dx = ctx.needs_input_grad[2]
        # ---- gradient arena: per block its three (one) gradient slabs and its input gradient(s)
        gbytes = 0
        for b in blocks_c:
            nb = _al(b.u1.n_out * b.u1.c_out * esz)
            gbytes += (3 if b.has_u2 else 1) * nb + (2 if b.has_ud else 1) * _al(b.u1.n_in * b.u1.c_in * esz)
        garena = _Bump(gbytes, dev)
        # Parameter gradients.  Default: a zero arena whose slices are handed to autograd.  With a gradient sink (a data-parallel
        # reducer, ddp.GradBucketReducer) the kernels accumulate straight into the reducer's flat all-reduce buckets (zeroed by
        # its reset()), nothing is returned to autograd for those parameters, and the reducer is told per network stage -- with
        # an event the library records behind the stage's kernels -- which gradients are final, so the all-reduce of a bucket
        # starts while the rest of the backward pass is still running.
        sink = ops.grad_sink() if (dev.type == "cuda" or ops.SINK_ON_HOST) else None
        sink_ptrs = None
        if sink is not None:
            sink_ptrs = [sink.grad_ptr(p) for p in plan.parameters()]
            if any(q is None for q in sink_ptrs):
                sink, sink_ptrs = None, None
        grads = ops.zeros_f32.take(plan.sum_w + 2 * plan.sum_c, dev) if sink is None else None
        sums = ops.zeros_f64.take(2 * plan.sum_c, dev)
        gp, sp = (grads.data_ptr() if grads is not None else 0), sums.data_ptr()
        prog = (Op * (plan.n_blocks + 32))()
        n_ops = 0
        # parameter-gradient slots, in plan.units order (weights first, then [dgamma | dbeta] of the unit)
        unit_slots, woff, coff = [], 0, 0
        for size, (conv, _) in zip(sizes, plan.units):
            c = conv.out_channels
            unit_slots.append((woff, size, plan.sum_w + 2 * coff, c, coff))
            woff += size
            coff += c
        # units of block j, in plan order
        ui = 0
        block_units = []
        for b in blocks_c:
            k = 1 + int(b.has_u2) + int(b.has_ud)
            block_units.append(list(range(ui, ui + k)))
            ui += k

        def add_cols(dst, ld_dst, src, ld_src, n, c):
            nonlocal n_ops
            o = prog[n_ops]
            o.op, o.dtype, o.dst, o.ld_dst, o.src, o.ld_src, o.n, o.c = OP_ADD_COLS, dtc, dst, ld_dst, src, ld_src, n, c
            n_ops += 1

        def ext_grad(g, n, c):
            """External gradient of a stage output as a dense buffer of the activation dtype (or None)."""
            if g is None:
                return None
            g = ops._rowmajor(g)
            if g.dtype != dt:
                g = g.to(dt)
            if g.shape[0] > 1 and g.stride(0) != c:
                g = g.contiguous()
            keep.append(g)
            return g.data_ptr()

        keep = []
        stage_done = []                        # (event, unit indices) per stage, in the order the stages finish
        # block index ranges per stage (forward order): encoder stage i = [single, blocks...], decoder likewise
        stages, bi = [], 0
        for conv, bn, blks in plan.encoder + plan.decoder:
            stages.append((bi, bi + 1 + len(blks)))
            bi += 1 + len(blks)
        # gradient flowing into the output of block j (dense [n_out, c_out]); filled while walking backwards
        g_in = [None] * plan.n_blocks
        ld_in = [0] * plan.n_blocks
        skip_extra = {}                        # last block of encoder stage s -> (ptr, ld, c_off) of the decoder's gradient slice
        for s in range(7, -1, -1):
            first, last = stages[s]
            n_out, c_out = blocks_c[last - 1].u1.n_out, blocks_c[last - 1].u1.c_out
            # gradient of the stage output: from the consumer inside the trunk (already in g_in) plus the caller's
            eg = ext_grad(gouts[s], n_out, c_out)
            if g_in[last - 1] is None:
                if eg is None:
                    raise RuntimeError("internal error: a trunk stage received no gradient")
                g_in[last - 1], ld_in[last - 1] = eg, 0
            elif eg is not None:
                add_cols(g_in[last - 1], ld_in[last - 1] or c_out, eg, c_out, n_out, c_out)
            if s in skip_extra:
                ptr, ld_src = skip_extra[s]
                add_cols(g_in[last - 1], ld_in[last - 1] or c_out, ptr, ld_src, n_out, c_out)
            for j in range(last - 1, first - 1, -1):
                b = blocks_c[j]
                n_o, c_o, n_i, c_i = b.u1.n_out, b.u1.c_out, b.u1.n_in, b.u1.c_in
                nb = n_o * c_o * esz
                b.gout, b.ld_gout = g_in[j], ld_in[j]
                stem_input = s == 0 and j == first
                b.need_dx = int(need_dx or not stem_input)
                if b.has_u2:
                    b.dy2, _ = garena.take(nb)
                    b.dres, _ = garena.take(nb)
                    b.da1, _ = garena.take(nb)
                    b.dy1, b.dyd = b.da1, b.dres       # in place: the first unit's BN backward on the second's dgrad, the shortcut's on the residual gradient
                else:
                    b.dy1, _ = garena.take(nb)
                dx_off = None
                if b.need_dx:
                    b.dx, dx_off = garena.take(n_i * c_i * esz)
                    if b.has_ud:
                        b.dxd, _ = garena.take(n_i * c_i * esz)
                for u, k in zip((b.u1, b.u2, b.ud), block_units[j]):
                    wo, size, go, c, co = unit_slots[k]
                    u.sums = sp + 16 * co
                    if sink_ptrs is not None:
                        u.dw, u.dgamma, u.dbeta = sink_ptrs[3 * k], sink_ptrs[3 * k + 1], sink_ptrs[3 * k + 2]
                    else:
                        u.dw, u.dgamma, u.dbeta = gp + 4 * wo, gp + 4 * go, gp + 4 * (go + c)
                prog[n_ops].op, prog[n_ops].block = OP_BLOCK_BACKWARD, C.pointer(b)
                n_ops += 1
                if j > first:
                    g_in[j - 1], ld_in[j - 1] = b.dx, 0
                elif s >= 4:
                    # first block of a decoder stage's residual chain is at first + 1; `first` is the transposed conv whose
                    # input is the previous stage's output
                    g_in[stages[s - 1][1] - 1], ld_in[stages[s - 1][1] - 1] = b.dx, 0
                elif s > 0:
                    g_in[stages[s - 1][1] - 1], ld_in[stages[s - 1][1] - 1] = b.dx, 0
                else:
                    dx_first = (dx_off, n_i, c_i) if b.need_dx else None
                # the block right after a decoder stage's transposed conv reads [up | skip]: split its input gradient
                if s >= 4 and j == first + 1:
                    c_up = blocks_c[first].u1.c_out
                    g_in[first], ld_in[first] = b.dx, c_i                      # left columns, strided
                    skip_extra[7 - s - 1] = (b.dx + c_up * esz, c_i)           # right columns -> the encoder stage of that resolution
            if sink is not None:
                ev = sink.stage_event(7 - s)
                o = prog[n_ops]
                o.op, o.dst = OP_RECORD_EVENT, ev.cuda_event
                n_ops += 1
                stage_done.append((ev, [k for j in range(first, last) for k in block_units[j]]))
        # the skip of tensor stride 1 is the trunk's input (the stem activation): its decoder slice joins dx of the first block
        if need_dx and -1 in skip_extra:
            ptr, ld_src = skip_extra[-1]
            add_cols(blocks_c[0].dx, 0 or blocks_c[0].u1.c_in, ptr, ld_src, blocks_c[0].u1.n_in, blocks_c[0].u1.c_in)
        with nvtx_range("gcd:trunk backward"):
            ops.run_ops(prog, n_ops, backward=True)
        dx = None
        if need_dx:
            dx = garena.view(dx_first[0], dx_first[1], dx_first[2], dt)
            if dx.dtype != ctx.x_dtype:
                dx = dx.to(ctx.x_dtype)
        out = [None, None, dx]
        if sink is not None:
            params = plan.parameters()
            for ev, units in stage_done:
                sink.mark_ready([params[3 * k + i] for k in units for i in range(3)], ev)
            return tuple(out + [None] * (3 * len(plan.units)))
        for (wo, size, go, c, _), (conv, _) in zip(unit_slots, plan.units):
            out += [grads[wo:wo + size].view(conv.kernel.shape), grads[go:go + c], grads[go + c:go + 2 * c]]
        return tuple(out)
