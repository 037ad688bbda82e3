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
from .config import get_math_mode

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
