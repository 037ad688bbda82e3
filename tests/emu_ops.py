"""Host emulation of gcd_run_ops / gcd_block_forward / gcd_block_backward (csrc/block.cu) for CPU tests of the Python
plumbing: interprets the very ctypes structures the product hands to the library, reading and writing the caller's
buffers through their raw addresses, with the arithmetic done in float64 torch (neighbour-table gather, batch statistics,
pair-list weight gradient).  Test infrastructure: follows the contract of include/gcdlss_b200.h, not the kernels' code."""
import ctypes as C

import numpy as np
import torch

_NP = {0: np.float32, 1: np.uint16}          # gcd_dtype -> storage


def _mat(ptr, n, c, ld, dtype_code):
    """[n, c] view (row pitch ld elements) of caller memory as a writable numpy array of the storage type."""
    if not ptr or n == 0:
        return None
    esz = 4 if dtype_code == 0 else 2
    buf = (C.c_char * (((n - 1) * ld + c) * esz)).from_address(ptr)
    flat = np.frombuffer(buf, dtype=_NP[dtype_code])
    return np.lib.stride_tricks.as_strided(flat, shape=(n, c), strides=(ld * esz, esz))


def _load(ptr, n, c, ld, dtype_code):
    m = _mat(ptr, n, c, ld, dtype_code)
    t = torch.from_numpy(np.ascontiguousarray(m))
    return (t.view(torch.bfloat16) if dtype_code == 1 else t).double()


def _store(ptr, n, c, ld, dtype_code, value):
    m = _mat(ptr, n, c, ld, dtype_code)
    v = value.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16) if dtype_code == 1 else value.float().numpy()
    m[...] = v


def _vec(ptr, n, np_dtype):
    return np.frombuffer((C.c_char * (n * np.dtype(np_dtype).itemsize)).from_address(ptr), dtype=np_dtype) if ptr else None


def _table(ptr, kv, n_out):
    return _vec(ptr, kv * n_out, np.int32).reshape(kv, n_out) if ptr else None


def _conv(u, x, nbr_ptr, n_out, w, mirror, transpose, out_rows_ptr):
    kv = u.kv
    c_out = w.shape[1] if transpose else w.shape[2]
    cols = torch.zeros((n_out, c_out), dtype=torch.float64)
    nbr = _table(nbr_ptr, kv, n_out)
    for k in range(kv):
        wk = w[kv - 1 - k if mirror else k]
        b = wk.t() if transpose else wk
        if nbr is None:
            cols += x @ b
        else:
            idx = torch.from_numpy(nbr[k].astype(np.int64))
            o = torch.nonzero(idx >= 0).reshape(-1)
            cols.index_add_(0, o, x[idx[o]] @ b)
    if out_rows_ptr:
        rows = torch.from_numpy(_vec(out_rows_ptr, n_out, np.int32).astype(np.int64))
        cols = torch.zeros_like(cols).index_copy_(0, rows, cols)
    return cols


def _weights(u):
    return torch.from_numpy(_vec(u.w, u.kv * u.c_in * u.c_out, np.float32).reshape(u.kv, u.c_in, u.c_out).copy()).double()


def _bn_forward(u, y, res, relu):
    n, c = y.shape
    stats = _vec(u.stats, 2 * c, np.float64)
    stats[:c] += y.sum(0).numpy()
    stats[c:] += (y * y).sum(0).numpy()
    mean = torch.from_numpy(stats[:c] / n)
    var = torch.from_numpy(stats[c:] / n) - mean * mean
    invstd = torch.rsqrt(var + u.eps)
    _vec(u.mean, c, np.float32)[:] = mean.float().numpy()
    _vec(u.invstd, c, np.float32)[:] = invstd.float().numpy()
    rm, rv = _vec(u.running_mean, c, np.float32), _vec(u.running_var, c, np.float32)
    rm[:] = (1 - u.momentum) * rm + u.momentum * mean.float().numpy()
    rv[:] = (1 - u.momentum) * rv + u.momentum * (var * n / max(n - 1, 1)).float().numpy()
    gamma, beta = torch.from_numpy(_vec(u.gamma, c, np.float32).copy()).double(), torch.from_numpy(_vec(u.beta, c, np.float32).copy()).double()
    out = (y - mean) * invstd * gamma + beta
    if res is not None:
        out = out + res
    return torch.relu(out) if relu else out


def block_forward(b):
    dt = b.dtype
    u1 = b.u1
    x = _load(b.x, u1.n_in, u1.c_in, b.ld_x, dt)
    y1 = _conv(u1, x, u1.nbr, u1.n_out, _weights(u1), False, False, u1.out_rows)
    _store(b.y1, u1.n_out, u1.c_out, u1.c_out, dt, y1)
    y1 = _load(b.y1, u1.n_out, u1.c_out, u1.c_out, dt)                 # as stored (bf16 rounding)
    a1 = _bn_forward(u1, y1, None, 1 if b.has_u2 else b.relu1)
    _store(b.a1, u1.n_out, u1.c_out, u1.c_out, dt, a1)
    if not b.has_u2:
        return
    a1 = _load(b.a1, u1.n_out, u1.c_out, u1.c_out, dt)
    u2 = b.u2
    y2 = _conv(u2, a1, u2.nbr, u2.n_out, _weights(u2), False, False, u2.out_rows)
    _store(b.y2, u2.n_out, u2.c_out, u2.c_out, dt, y2)
    y2 = _load(b.y2, u2.n_out, u2.c_out, u2.c_out, dt)
    res = x
    if b.has_ud:
        ud = b.ud
        yd = _conv(ud, x, ud.nbr, ud.n_out, _weights(ud), False, False, ud.out_rows)
        _store(b.yd, ud.n_out, ud.c_out, ud.c_out, dt, yd)
        rd = _bn_forward(ud, _load(b.yd, ud.n_out, ud.c_out, ud.c_out, dt), None, 0)
        _store(b.rd, ud.n_out, ud.c_out, ud.c_out, dt, rd)
        res = _load(b.rd, ud.n_out, ud.c_out, ud.c_out, dt)
    out = _bn_forward(u2, y2, res, 1)
    _store(b.out, u2.n_out, u2.c_out, u2.c_out, dt, out)


def _bn_backward(u, dy, x, y, relu):
    n, c = x.shape
    g = dy * (y > 0) if relu else dy
    mean = torch.from_numpy(_vec(u.mean, c, np.float32).copy()).double()
    invstd = torch.from_numpy(_vec(u.invstd, c, np.float32).copy()).double()
    gamma = torch.from_numpy(_vec(u.gamma, c, np.float32).copy()).double()
    xhat = (x - mean) * invstd
    sg, sgx = g.sum(0), (g * xhat).sum(0)
    _vec(u.dgamma, c, np.float32)[:] += sgx.float().numpy()
    _vec(u.dbeta, c, np.float32)[:] += sg.float().numpy()
    return gamma * invstd * (g - sg / n - xhat * sgx / n), g


def _wgrad(u, x, dy):
    kv = u.kv
    dw = _vec(u.dw, kv * u.c_in * u.c_out, np.float32).reshape(kv, u.c_in, u.c_out)
    if not u.pair_in:
        dw[0] += (x.t() @ dy).float().numpy()
        return
    off = _vec(u.pair_off, kv + 1, np.int32)
    pi, po = _vec(u.pair_in, int(off[kv]), np.int32).astype(np.int64), _vec(u.pair_out, int(off[kv]), np.int32).astype(np.int64)
    for k in range(kv):
        a, e = int(off[k]), int(off[k + 1])
        if e > a:
            dw[k] += (x[pi[a:e]].t() @ dy[po[a:e]]).float().numpy()


def _dgrad(u, dy):
    return _conv(u, dy, u.back_nbr, u.n_in, _weights(u), bool(u.back_mirror), True, u.back_out_rows)


def block_backward(b):
    dt = b.dtype
    u1, u2, ud = b.u1, b.u2, b.ud
    n, c = u1.n_out, u1.c_out
    x = _load(b.x, u1.n_in, u1.c_in, b.ld_x, dt)
    gout = _load(b.gout, n, c, b.ld_gout or c, dt)
    g1 = gout
    dres = None
    if b.has_u2:
        dy2, dres = _bn_backward(u2, gout, _load(b.y2, n, c, c, dt), _load(b.out, n, c, c, dt), True)
        _store(b.dy2, n, c, c, dt, dy2)
        _store(b.dres, n, c, c, dt, dres)
        dy2, dres = _load(b.dy2, n, c, c, dt), _load(b.dres, n, c, c, dt)
        g1 = _dgrad(u2, dy2)
        _store(b.da1, n, c, c, dt, g1)
        g1 = _load(b.da1, n, c, c, dt)
        _wgrad(u2, _load(b.a1, n, c, c, dt), dy2)
    relu1 = 1 if b.has_u2 else b.relu1
    dy1, _ = _bn_backward(u1, g1, _load(b.y1, n, c, c, dt), _load(b.a1, n, c, c, dt) if relu1 else None, bool(relu1))
    _store(b.dy1, n, c, c, dt, dy1)
    dy1 = _load(b.dy1, n, c, c, dt)
    dx = None
    if b.need_dx:
        dx = _dgrad(u1, dy1)
        _store(b.dx, u1.n_in, u1.c_in, u1.c_in, dt, dx)
    _wgrad(u1, x, dy1)
    if b.has_u2:
        if b.has_ud:
            dyd, _ = _bn_backward(ud, dres, _load(b.yd, n, c, c, dt), None, False)
            _store(b.dyd, n, c, c, dt, dyd)
            dyd = _load(b.dyd, n, c, c, dt)
            if b.need_dx:
                _store(b.dxd, u1.n_in, u1.c_in, u1.c_in, dt, _dgrad(ud, dyd))
                tot = _load(b.dx, u1.n_in, u1.c_in, u1.c_in, dt) + _load(b.dxd, u1.n_in, u1.c_in, u1.c_in, dt)
                _store(b.dx, u1.n_in, u1.c_in, u1.c_in, dt, tot)
            _wgrad(ud, x, dyd)
        elif b.need_dx:
            tot = _load(b.dx, u1.n_in, u1.c_in, u1.c_in, dt) + dres
            _store(b.dx, u1.n_in, u1.c_in, u1.c_in, dt, tot)


def run_ops(prog, n_ops, stream, launches):
    for i in range(n_ops):
        o = prog[i]
        if o.op == 0:
            block_forward(o.block.contents)
        elif o.op == 1:
            block_backward(o.block.contents)
        elif o.op in (2, 3):
            src = _load(o.src, o.n, o.c, o.ld_src, o.dtype)
            if o.op == 3:
                src = src + _load(o.dst, o.n, o.c, o.ld_dst, o.dtype)
            _store(o.dst, o.n, o.c, o.ld_dst, o.dtype, src)
        elif o.op == 4:                     # GCD_OP_RECORD_EVENT: nothing to order on the host; the test reads the order from this log
            recorded_events.append(int(o.dst))
        else:
            raise ValueError(f"unknown op {o.op}")
    return 0


recorded_events = []
