"""Tensor-level wrappers over the C ABI: every function takes CUDA torch tensors (used only as
device memory + the current stream) and enqueues hand-written sm_100a kernels.  Nothing here
computes with torch ops."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from ._cabi import BF16, F32, MATH_BF16_TC, MATH_FP32_SIMT, ConvArgs, WgradArgs, call, lib

# Count of kernel-launching C-ABI calls, for bench.py's "gpu_launches" claim.
launch_counter = {"calls": 0}


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_device = torch._C._cuda_getDevice


def _stream():
    """cudaStream_t of torch's current stream (raw pointer value; cheap enough to call per launch)."""
    return _raw_stream(_cur_device())


class ZeroArena:
    """Hands out zero-filled scratch (per-channel statistics, weight-gradient accumulators) carved from
    large pre-zeroed chunks, so a training step issues one memset per chunk instead of one per layer.
    A slice is used once; a chunk is released when its last slice dies.  ``new_period`` (called when a new
    coordinate manager, i.e. a new batch, appears) starts a fresh chunk sized to the previous period's use,
    which keeps the allocation sequence identical from step to step (no allocator growth in steady state)."""

    def __init__(self, dtype, min_chunk):
        self.dtype, self.min_chunk, self.chunk = dtype, min_chunk, min_chunk
        self.buf, self.off, self.used = None, 0, 0

    def new_period(self):
        if self.used:
            self.chunk = max(self.min_chunk, (int(self.used * 1.02) + 4095) & ~4095)
        self.buf, self.off, self.used = None, 0, 0

    def take(self, n: int, device) -> torch.Tensor:
        n_al = (n + 3) & ~3
        if self.buf is None or self.buf.device != device or self.off + n_al > self.buf.numel():
            self.buf = torch.zeros(max(self.chunk, n_al), dtype=self.dtype, device=device)
            self.off = 0
        out = self.buf[self.off:self.off + n]
        self.off += n_al
        self.used += n_al
        return out


zeros_f64 = ZeroArena(torch.float64, 1 << 14)
zeros_f32 = ZeroArena(torch.float32, 1 << 16)


def new_batch():
    """Marks a batch boundary for the scratch arenas (called by CoordinateManager.__init__)."""
    zeros_f64.new_period()
    zeros_f32.new_period()


def set_option(option: int, value: int) -> None:
    """Process-wide tuning option of the library (``_cabi.OPT_*``; gcd_set_option)."""
    call("gcd_set_option", int(option), int(value))


def get_option(option: int) -> int:
    return int(lib().gcd_get_option(int(option)))


class _ExecContexts:
    """Execution contexts of the library (gcd_exec_create: a second stream for the weight gradients of a backward pass,
    the events that order it, the counters of the dynamic tile schedule), one per (device, stream) the trunk runs on.
    Created on first use, destroyed with the process (a context may only be used from one host thread at a time: the
    autograd engine runs the backward of one device on one thread, the forward on the caller's)."""

    def __init__(self):
        self._ctx = {}

    def get(self, stream_ptr: int, backward: bool) -> C.c_void_p:
        key = (_cur_device(), int(stream_ptr), backward)
        h = self._ctx.get(key)
        if h is None:
            out = C.c_void_p(0)
            call("gcd_exec_create", C.byref(out))
            h = self._ctx[key] = out
        return h

    def close(self):
        for h in self._ctx.values():
            lib().gcd_exec_destroy(h)
        self._ctx.clear()


exec_contexts = _ExecContexts()

# Gradient sink: a data-parallel reducer (ddp.GradBucketReducer) registers itself here; the trunk's backward pass then lets
# the kernels accumulate parameter gradients straight into the reducer's buckets (functional.TrunkFunction.backward).
_grad_sink = {"ref": None}
SINK_ON_HOST = False        # CPU plumbing tests only (tests/test_trunk_plumbing.py): let the sink see host tensors


def set_grad_sink(sink) -> None:
    import weakref
    _grad_sink["ref"] = weakref.ref(sink) if sink is not None else None


def grad_sink():
    r = _grad_sink["ref"]
    return r() if r is not None else None


def run_ops(prog, n_ops: int, backward: bool) -> int:
    """gcd_run_ops_exec on the current stream with that stream's execution context; returns the number of kernels launched."""
    launches = C.c_int32(0)
    st = _stream()
    call("gcd_run_ops_exec", exec_contexts.get(st, backward), prog, n_ops, st, C.byref(launches))
    _count(launches.value)
    return launches.value


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported feature dtype {t.dtype}")


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gcdlss_b200 ops need CUDA tensors (there is no CPU fallback)")


def _count(n=1):
    launch_counter["calls"] += n


class KernelTimer:
    """Optional instrumentation of the convolution launches for bench.py's live roofline.

    ``enabled``: CUDA-event brackets around every launch inside the timed region (on the launching stream).  When the
    step is host-bound those intervals also contain the idle gap before the launch reaches the GPU, so bench.py
    additionally uses ``capture``: the launches of one step are remembered (a closure that re-issues the identical
    call on the same tensors) and replayed back to back after the timed region to measure the kernels alone."""

    def __init__(self):
        self.enabled = False
        self.capture = False
        self.records = []        # (kind, nbr_ptr, n_out, kv, c_in, c_out, start_event, end_event)
        self.captured = []       # (kind, nbr_ptr, n_out, kv, c_in, c_out, replay_fn)

    def bracket(self, kind, nbr, n_out, kv, c_in, c_out):
        if not self.enabled:
            return None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.records.append((kind, nbr.data_ptr() if nbr is not None else 0, n_out, kv, c_in, c_out, e0, e1))
        return e1

    def remember(self, kind, nbr, n_out, kv, c_in, c_out, fn):
        if self.capture:
            self.captured.append((kind, nbr.data_ptr() if nbr is not None else 0, n_out, kv, c_in, c_out, fn))


kernel_timer = KernelTimer()


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    return t if (t.dim() == 2 and t.stride(1) == 1) or t.numel() == 0 else t.contiguous()


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else t.shape[1]


# ------------------------------------------------------------------------------------ quantise
def quantize(points: torch.Tensor, q: float, dims: int, round_mode: int) -> torch.Tensor:
    """int32 [n, dims] = round_mode(points[:, :dims] / q); points fp32 or fp64, row-major."""
    _require_cuda(points)
    points = _rowmajor(points)
    n = points.shape[0]
    out = torch.empty((n, dims), dtype=torch.int32, device=points.device)
    if points.dtype == torch.float32:
        call("gcd_quantize_f32", _ptr(points), _ld(points), n, dims, float(q), round_mode, _ptr(out), _stream())
    elif points.dtype == torch.float64:
        call("gcd_quantize_f64", _ptr(points), _ld(points), n, dims, float(q), round_mode, _ptr(out), _stream())
    else:
        raise TypeError("points must be float32 or float64")
    _count()
    return out


def affine_f64(points: torch.Tensor, matrix) -> torch.Tensor:
    """float64 [n, 3] = [points | 1] @ matrix.T[:, :3] for float32 CUDA points [n, >= 3] and a 4x4 (or 3x4) float64 host matrix
    (the dataset's rigid / voxelisation transform, ref utils/dataset_remission.py:821-833)."""
    import numpy as np
    _require_cuda(points)
    if points.dtype != torch.float32:
        raise TypeError("points must be float32 (the datasets read float32 scans)")
    points = _rowmajor(points)
    m = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64)[:3, :4])
    out = torch.empty((points.shape[0], 3), dtype=torch.float64, device=points.device)
    call("gcd_affine_f64", _ptr(points), _ld(points), points.shape[0], m.ctypes.data_as(C.c_void_p), _ptr(out), _stream())
    _count()
    return out


def shift_to_min(coords: torch.Tensor) -> torch.Tensor:
    """coords -= coords.min(0), in place (ref models/voxelizer.py:276)."""
    n, dims = coords.shape
    mins = torch.full((4,), 2**31 - 1, dtype=torch.int32, device=coords.device)
    call("gcd_colmin_i32", _ptr(coords), n, dims, _ptr(mins), _stream())
    call("gcd_sub_cols_i32", _ptr(coords), n, dims, _ptr(mins), _stream())
    _count(2)
    return coords


class HashTable:
    """Device memory of one open-addressing table (keys uint64 as int64 storage, vals int32)."""

    def __init__(self, n: int, device):
        self.cap = int(lib().gcd_hash_capacity(n))
        self.keys = torch.empty(self.cap, dtype=torch.int64, device=device)
        self.vals = torch.empty(self.cap, dtype=torch.int32, device=device)


def _status_check(status: torch.Tensor, what: str):
    s = int(status.item())
    if s:
        msgs = []
        if s & _cabi.DEV_KEY_RANGE:
            msgs.append("a coordinate does not fit the 64-bit key (|coord| < 2^17, 0 <= batch < 1023)")
        if s & _cabi.DEV_DUPLICATE:
            msgs.append("duplicate coordinates (quantise the input first)")
        if s & _cabi.DEV_TABLE_FULL:
            msgs.append("hash table full")
        raise RuntimeError(f"{what}: " + "; ".join(msgs))


def unique_rows(icoords: torch.Tensor, order: int = 0):
    """Unique rows of int32 [n, 3|4].  Returns (unique_idx [M] int64, inverse [n] int64, table)."""
    _require_cuda(icoords)
    icoords = icoords.contiguous()
    n, dims = icoords.shape
    dev = icoords.device
    table = HashTable(n, dev)
    ws_bytes = int(lib().gcd_unique_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    unique_idx = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    inverse = torch.empty(n, dtype=torch.int64, device=dev)
    meta = torch.zeros(2, dtype=torch.int32, device=dev)  # [m, status]
    call("gcd_unique_rows", _ptr(icoords), n, dims, order, _ptr(table.keys), _ptr(table.vals), table.cap, _ptr(unique_idx),
         _ptr(inverse), _ptr(meta[0:1]), _ptr(ws), ws_bytes, _ptr(meta[1:2]), _stream())
    _count(7 if order == 0 else 40)
    m, status = meta.tolist()
    if status:
        _status_check(meta[1:2], "sparse_quantize")
    return unique_idx[:m], inverse, table


# ------------------------------------------------------------------------------------ coordinate maps
def hash_build(coords: torch.Tensor, status: torch.Tensor) -> HashTable:
    n = coords.shape[0]
    table = HashTable(n, coords.device)
    call("gcd_hash_build", _ptr(coords), n, _ptr(table.keys), _ptr(table.vals), table.cap, _ptr(status), _stream())
    _count(2)
    return table


def coords_stride2(coords: torch.Tensor, ts: int, status: torch.Tensor):
    """Returns (coarse coords [M,4], parent [n], code [n], coarse table)."""
    n = coords.shape[0]
    dev = coords.device
    table = HashTable(n, dev)
    ws_bytes = int(lib().gcd_stride2_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    coarse = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    parent = torch.empty(n, dtype=torch.int32, device=dev)
    code = torch.empty(n, dtype=torch.int32, device=dev)
    m_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    call("gcd_coords_stride2", _ptr(coords), n, ts, _ptr(table.keys), _ptr(table.vals), table.cap, _ptr(coarse), _ptr(parent),
         _ptr(code), _ptr(m_dev), _ptr(ws), ws_bytes, _ptr(status), _stream())
    _count(8)
    m = int(m_dev.item())
    return coarse[:m], parent, code, table


def kmap_subm(coords: torch.Tensor, table: HashTable, kernel_size: int, ts: int) -> torch.Tensor:
    n = coords.shape[0]
    kv = kernel_size ** 3
    nbr = torch.empty((kv, n), dtype=torch.int32, device=coords.device)
    call("gcd_kmap_subm", _ptr(coords), n, _ptr(table.keys), _ptr(table.vals), table.cap, kernel_size, ts, _ptr(nbr), _stream())
    _count()
    return nbr


class RunTable:
    """Device memory of one run table: ``cap`` 32-byte slots (csrc/runtable.cuh), int64 storage so the base is 32-byte aligned."""

    def __init__(self, n: int, device):
        self.cap = int(lib().gcd_hash_capacity(n))
        words = int(lib().gcd_runtable_slot_bytes()) // 8
        self.slots = torch.empty((self.cap, words), dtype=torch.int64, device=device)


def runtable_build(coords: torch.Tensor, ts: int, status: torch.Tensor) -> RunTable:
    """Run table of the unique int32 [n, 4] coordinates of one map at tensor stride ``ts``."""
    n = coords.shape[0]
    table = RunTable(n, coords.device)
    call("gcd_runtable_build", _ptr(coords), n, ts, _ptr(table.slots), table.cap, _ptr(status), _stream())
    _count(2)
    return table


def kmap_subm_runs(coords: torch.Tensor, table: RunTable, kernel_size: int, ts: int) -> torch.Tensor:
    """Same result as :func:`kmap_subm`, searched in a run table."""
    n = coords.shape[0]
    kv = kernel_size ** 3
    nbr = torch.empty((kv, n), dtype=torch.int32, device=coords.device)
    call("gcd_kmap_subm_runs", _ptr(coords), n, _ptr(table.slots), table.cap, kernel_size, ts, _ptr(nbr), _stream())
    _count()
    return nbr


def kmap_tile_sort(nbr: torch.Tensor):
    """Tile-sorted copy of a 3x3x3 / 2x2x2 table for the tcgen05 convolution:
    (nbr_sorted [kv, n] = nbr[:, rows], rows [n] int32, tile_masks [ceil(n / 128)] int32: offsets with a hit per 128-column tile)."""
    kv, n = nbr.shape
    dev = nbr.device
    nbr_sorted = torch.empty_like(nbr)
    rows = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    masks = torch.empty(max((n + 127) // 128, 1), dtype=torch.int32, device=dev)
    ws_bytes = int(lib().gcd_tile_sort_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("gcd_kmap_tile_sort", _ptr(nbr), n, kv, _ptr(nbr_sorted), _ptr(rows), _ptr(masks), _ptr(ws), ws_bytes, _stream())
    _count(23 if kv == 27 else 13)
    return nbr_sorted, rows[:n], masks


def kmap_down2(parent, code, n_coarse: int) -> torch.Tensor:
    nbr = torch.empty((8, n_coarse), dtype=torch.int32, device=parent.device)
    call("gcd_kmap_down2", _ptr(parent), _ptr(code), parent.shape[0], n_coarse, _ptr(nbr), _stream())
    _count(2)
    return nbr


def kmap_up2(parent, code) -> torch.Tensor:
    nbr = torch.empty((8, parent.shape[0]), dtype=torch.int32, device=parent.device)
    call("gcd_kmap_up2", _ptr(parent), _ptr(code), parent.shape[0], _ptr(nbr), _stream())
    _count()
    return nbr


def pairs_from_table(nbr: torch.Tensor):
    """(pair_in, pair_out, pair_off) of a [kv, n_out] table; arrays sized for the worst case."""
    kv, n_out = nbr.shape
    dev = nbr.device
    cap = max(kv * n_out, 1)
    pair_in = torch.empty(cap, dtype=torch.int32, device=dev)
    pair_out = torch.empty(cap, dtype=torch.int32, device=dev)
    pair_off = torch.empty(kv + 1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib().gcd_pairs_workspace_bytes(n_out, kv))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("gcd_pairs_from_table", _ptr(nbr), n_out, kv, _ptr(pair_in), _ptr(pair_out), _ptr(pair_off), _ptr(ws), ws_bytes, _stream())
    _count(5)
    return pair_in, pair_out, pair_off


# ------------------------------------------------------------------------------------ convolution
def tc_supported(c_in: int, c_out: int, kv: int) -> bool:
    return bool(lib().gcd_conv_tc_supported(c_in, c_out, kv))


def pack_weights(w: torch.Tensor, transpose: bool, mirror: bool) -> torch.Tensor:
    """bf16 operand image of fp32 kernel [kv, c_in, c_out] for the tcgen05 path."""
    kv, c_in, c_out = w.shape
    nbytes = int(lib().gcd_conv_packed_weight_bytes(kv, c_in, c_out))
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    call("gcd_conv_pack_weights", _ptr(w), kv, c_in, c_out, int(transpose), int(mirror), _ptr(packed), _stream())
    _count()
    return packed


def pack_weights_batched(desc_table_dev: torch.Tensor, n_descs: int, total_blocks: int):
    """One launch re-packing many kernels; desc_table_dev is a device uint8 tensor of gcd_pack_desc records."""
    call("gcd_conv_pack_weights_batched", desc_table_dev.data_ptr(), n_descs, total_blocks, _stream())
    _count()


def conv_forward(inp, nbr, w3, n_out: int, *, transpose_w=False, mirror=False, bias=None, out_dtype=None, math_mode=MATH_FP32_SIMT,
                 w_packed=None, stats=None, out_rows=None, tile_masks=None, sched=None):
    """out[o] = sum_k inp[nbr[k, o]] @ B_k, B_k = w3[k] (or w3[wsel(k)]^T when transpose_w).

    inp [n_in, c_in'] row-major; nbr [kv, n_out] int32 or None (identity); w3 fp32 [kv, c_in, c_out].
    ``out_rows`` (tcgen05 path only): ``nbr`` is a tile-sorted table (:func:`kmap_tile_sort`) and column i of it is
    output row out_rows[i]; ``tile_masks``: its per-tile offset masks; ``sched``: two zero int32 on the device, the
    kernel then claims its row tiles dynamically (gcd_conv_args.sched) and leaves the counters zero again.
    """
    _require_cuda(inp, w3)
    inp = _rowmajor(inp)
    kv, c_in, c_out = w3.shape
    k_dim, n_dim = (c_out, c_in) if transpose_w else (c_in, c_out)
    if inp.shape[1] != k_dim:
        raise ValueError(f"feature width {inp.shape[1]} does not match the kernel ({k_dim})")
    out_dtype = out_dtype or inp.dtype
    out = torch.empty((n_out, n_dim), dtype=out_dtype, device=inp.device)
    a = ConvArgs()
    a.inp, a.ld_in, a.n_in = inp.data_ptr(), _ld(inp), inp.shape[0]
    a.nbr = nbr.data_ptr() if nbr is not None else None
    a.kv, a.n_out, a.c_in, a.c_out = kv, n_out, k_dim, n_dim
    a.w = w3.data_ptr()
    a.w_packed = w_packed.data_ptr() if w_packed is not None else None
    if transpose_w:
        a.w_stride_k, a.w_stride_c, a.w_stride_n = c_in * c_out, 1, c_out
    else:
        a.w_stride_k, a.w_stride_c, a.w_stride_n = c_in * c_out, c_out, 1
    a.mirror = int(mirror)
    a.bias = bias.data_ptr() if bias is not None else None
    a.out, a.ld_out = out.data_ptr(), n_dim
    a.in_dtype, a.out_dtype = _dtype_code(inp), _dtype_code(out)
    a.stats = stats.data_ptr() if stats is not None else None
    a.math_mode = math_mode
    a.out_rows = out_rows.data_ptr() if out_rows is not None else None
    a.tile_masks = tile_masks.data_ptr() if tile_masks is not None else None
    a.sched = sched.data_ptr() if sched is not None else None
    kind = "conv_tc" if math_mode == MATH_BF16_TC else "conv_simt"
    end = kernel_timer.bracket(kind, nbr, n_out, kv, k_dim, n_dim)
    call("gcd_conv_forward", C.byref(a), _stream())
    if end is not None:
        end.record()
    if kernel_timer.capture:
        keep = (inp, nbr, w3, w_packed, bias, out, out_rows, tile_masks)        # keeps the operands alive for the replay
        kernel_timer.remember(kind, nbr, n_out, kv, k_dim, n_dim, lambda a=a, keep=keep: call("gcd_conv_forward", C.byref(a), _stream()))
    _count()
    return out


def conv_wgrad(inp, gout, pairs, kv: int, dw: torch.Tensor, dbias=None, math_mode=MATH_FP32_SIMT, sched=None):
    """dw[k] += inp[pair_in]^T gout[pair_out] (accumulating); pairs = (pair_in, pair_out, pair_off) or None (identity)."""
    _require_cuda(inp, gout, dw)
    inp, gout = _rowmajor(inp), _rowmajor(gout)
    a = WgradArgs()
    a.inp, a.ld_in = inp.data_ptr(), _ld(inp)
    a.gout, a.ld_gout = gout.data_ptr(), _ld(gout)
    if pairs is not None:
        a.pair_in, a.pair_out, a.pair_off = pairs[0].data_ptr(), pairs[1].data_ptr(), pairs[2].data_ptr()
        a.n_pairs = pairs[0].shape[0]
    else:
        a.pair_in = a.pair_out = a.pair_off = None
        a.n_pairs = inp.shape[0]
    a.kv, a.c_in, a.c_out = kv, inp.shape[1], gout.shape[1]
    a.dw = dw.data_ptr()
    a.dbias = dbias.data_ptr() if dbias is not None else None
    a.n_out = gout.shape[0]
    a.in_dtype, a.gout_dtype = _dtype_code(inp), _dtype_code(gout)
    a.math_mode = math_mode
    a.sched = sched.data_ptr() if sched is not None else None
    kind = "wgrad_tc" if math_mode == MATH_BF16_TC else "wgrad_simt"
    end = kernel_timer.bracket(kind, pairs[0] if pairs is not None else None, gout.shape[0], kv, inp.shape[1], gout.shape[1])
    call("gcd_conv_wgrad", C.byref(a), _stream())
    if end is not None:
        end.record()
    if kernel_timer.capture:
        keep = (inp, gout, pairs, dw, dbias)
        kernel_timer.remember(kind, pairs[0] if pairs is not None else None, gout.shape[0], kv, inp.shape[1], gout.shape[1],
                              lambda a=a, keep=keep: call("gcd_conv_wgrad", C.byref(a), _stream()))
    _count(2 if dbias is not None else 1)


def im2col(inp, nbr, ld_out: int, out_dtype) -> torch.Tensor:
    inp = _rowmajor(inp)
    kv, n_out = nbr.shape
    out = torch.empty((n_out, ld_out), dtype=out_dtype, device=inp.device)
    call("gcd_im2col", _ptr(inp), _ld(inp), inp.shape[1], _ptr(nbr), kv, n_out, _ptr(out), ld_out, _dtype_code(inp), _dtype_code(out), _stream())
    _count()
    return out


# ------------------------------------------------------------------------------------ batch norm
def bn_forward(x, gamma, beta, running_mean, running_var, training: bool, momentum: float, eps: float, relu: bool, residual=None,
               stats=None):
    """Returns (y, mean, invstd); mean/invstd are the statistics used (batch or running).
    ``stats``: optional fp64 [2c] per-channel sum / sum of squares already produced by the conv epilogue."""
    _require_cuda(x)
    x = _rowmajor(x)
    n, c = x.shape
    dev = x.device
    st = _stream()
    y = torch.empty_like(x)
    if residual is not None:
        residual = _rowmajor(residual)
    res_ptr, res_ld = (residual.data_ptr(), _ld(residual)) if residual is not None else (None, 0)
    if training:
        buf = torch.empty((2, c), dtype=torch.float32, device=dev)
        mean, invstd = buf[0], buf[1]
        if stats is None:
            stats = zeros_f64.take(2 * c, dev)
            call("gcd_bn_stats", x.data_ptr(), _ld(x), n, c, _dtype_code(x), stats.data_ptr(), st)
            _count()
        call("gcd_bn_apply_train", x.data_ptr(), _ld(x), n, c, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps), float(momentum),
             running_mean.data_ptr(), running_var.data_ptr(), mean.data_ptr(), invstd.data_ptr(), res_ptr, res_ld, int(relu), y.data_ptr(),
             _ld(y), _dtype_code(x), st)
        _count()
        return y, mean, invstd
    buf = torch.empty((2, c), dtype=torch.float32, device=dev)
    scale, shift = buf[0], buf[1]
    call("gcd_bn_fold_eval", c, gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(), running_var.data_ptr(), float(eps),
         scale.data_ptr(), shift.data_ptr(), st)
    call("gcd_bn_apply", x.data_ptr(), _ld(x), n, c, scale.data_ptr(), shift.data_ptr(), res_ptr, res_ld, int(relu), y.data_ptr(), _ld(y),
         _dtype_code(x), st)
    _count(2)
    return y, running_mean, torch.rsqrt(running_var + eps)


def bn_backward(dy, x, y, mean, invstd, gamma, relu: bool, training: bool, need_dres: bool):
    """Returns (dx, dres or None, dgamma, dbeta)."""
    dy = _rowmajor(dy)
    n, c = x.shape
    dev = x.device
    sums = zeros_f64.take(2 * c, dev)
    st = _stream()
    yp, ldy = (y.data_ptr(), _ld(y)) if y is not None else (None, 0)
    call("gcd_bn_backward_reduce", dy.data_ptr(), _ld(dy), x.data_ptr(), _ld(x), yp, ldy, n, c, mean.data_ptr(), invstd.data_ptr(),
         int(relu), _dtype_code(x), sums.data_ptr(), st)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if need_dres else None
    dgb = zeros_f32.take(2 * c, dev)
    dgamma, dbeta = dgb[:c], dgb[c:]
    call("gcd_bn_backward_apply", dy.data_ptr(), _ld(dy), x.data_ptr(), _ld(x), yp, ldy, n, c, mean.data_ptr(), invstd.data_ptr(),
         gamma.data_ptr(), sums.data_ptr(), int(relu), int(training), dx.data_ptr(), _ld(dx), dres.data_ptr() if dres is not None else None,
         _ld(dres) if dres is not None else 0, dgamma.data_ptr(), dbeta.data_ptr(), _dtype_code(x), st)
    _count(2)
    return dx, dres, dgamma, dbeta


def relu(x):
    x = x.contiguous()
    y = torch.empty_like(x)
    call("gcd_relu", _ptr(x), _ptr(y), x.numel(), _dtype_code(x), _stream())
    _count()
    return y


def relu_backward(dy, y):
    dy = dy.contiguous()
    dx = torch.empty_like(dy)
    call("gcd_relu_backward", _ptr(dy), _ptr(y), _ptr(dx), dy.numel(), _dtype_code(dy), _stream())
    _count()
    return dx


# ------------------------------------------------------------------------------------ voxel <-> point
def rows_gather(x, idx):
    _require_cuda(x, idx)
    x = _rowmajor(x.float())
    idx = idx.contiguous()
    out = torch.empty((idx.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    call("gcd_rows_gather", _ptr(x), _ld(x), _ptr(idx), idx.shape[0], x.shape[1], _ptr(out), x.shape[1], _stream())
    _count()
    return out


def csr_build(idx, n_segments: int):
    """(seg_off [n_segments+1] int32, order [n_points] int32): points grouped by segment, ascending."""
    idx = idx.contiguous()
    n = idx.shape[0]
    dev = idx.device
    seg_off = torch.empty(n_segments + 1, dtype=torch.int32, device=dev)
    order = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    ws_bytes = int(lib().gcd_csr_workspace_bytes(n, n_segments))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("gcd_csr_build", _ptr(idx), n, n_segments, _ptr(seg_off), _ptr(order), _ptr(ws), ws_bytes, _stream())
    _count(12)
    return seg_off, order[:n]


def segment_reduce(x, seg_off, order, n_segments: int, mode: int):
    x = _rowmajor(x.float())
    out = torch.empty((n_segments, x.shape[1]), dtype=torch.float32, device=x.device)
    call("gcd_segment_reduce", _ptr(x), _ld(x), _ptr(seg_off), _ptr(order), n_segments, x.shape[1], mode, _ptr(out), x.shape[1], _stream())
    _count()
    return out


# ------------------------------------------------------------------------------------ loss side
def consistency_rows(logits_s: torch.Tensor, logits_t: torch.Tensor, threshold: float = 0.0, want_grad: bool = False):
    """One pass over two fp32 logit matrices [n, c]: (sq_err [n], max_prob [n], label [n] int64, grad [n, c] or None);
    see gcd_consistency_rows in include/gcdlss_b200.h."""
    _require_cuda(logits_s, logits_t)
    if logits_s.shape != logits_t.shape or logits_s.dim() != 2:
        raise ValueError("student and teacher logits must be [n, c] of the same shape")
    ls, lt = _rowmajor(logits_s.float()), _rowmajor(logits_t.float())
    n, c = ls.shape
    dev = ls.device
    sq_err = torch.empty(n, dtype=torch.float32, device=dev)
    max_prob = torch.empty(n, dtype=torch.float32, device=dev)
    label = torch.empty(n, dtype=torch.int64, device=dev)
    grad = torch.empty((n, c), dtype=torch.float32, device=dev) if want_grad else None
    call("gcd_consistency_rows", ls.data_ptr(), _ld(ls), lt.data_ptr(), _ld(lt), n, c, float(threshold), sq_err.data_ptr(), max_prob.data_ptr(),
         label.data_ptr(), grad.data_ptr() if grad is not None else None, c, _stream())
    _count()
    return sq_err, max_prob, label, grad
