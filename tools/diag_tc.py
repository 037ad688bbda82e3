"""GPU diagnostic: isolated timing of the tcgen05 convolution kernels on representative MinkUNet layers of one kitti_b4
batch, forward through the scan-order table, the tile-sorted table and the tile-sorted table with per-tile offset masks,
plus wgrad.  With the PROFILE=1 build (GCDLSS_LIB_PATH=.../libgcdlss_sm100a_profile.so) the per-role cycle counters of the
forward kernel are printed for each variant (max over CTAs)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import torch

import bench
import gcdlss_b200
import MinkowskiEngine as ME
from gcdlss_b200 import _cabi, ops

gcdlss_b200.set_math_mode("bf16")
gcdlss_b200.set_tile_sort(True, min_rows=1)
dev = torch.device("cuda:0")
hb = bench.make_host_batches("kitti", 4, None, 17, 0, 1)
bc, f, l = bench.quantize_batch_on_gpu(hb[0], 0.05, dev)
st = ME.SparseTensor(features=f, coordinates=bc)
mgr = st.coordinate_manager
profile = hasattr(_cabi.lib(), "gcd_debug_set_buffer")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def counters(fn):
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    _cabi.lib().gcd_debug_set_buffer(C.c_void_p(dbg.data_ptr()))
    fn()
    torch.cuda.synchronize()
    _cabi.lib().gcd_debug_set_buffer(None)
    d = dbg.view(148, 16).double()
    d = d[d[:, 0] > 0]
    m = d.max(0).values
    tiles, iters = max(float(d[:, 7].mean()), 1.0), max(float(d[:, 3].mean()), 1.0)
    return (f"tiles/CTA={tiles:.1f} stages/owner={iters:.0f} | producer0 total={m[0]:.0f} wait_table={m[1]:.0f} wait_slot={m[2]:.0f} "
            f"busy/stage={(m[0] - m[1] - m[2]) / iters:.0f} | mma total={m[4]:.0f} wait_full={m[5]:.0f} wait_acc={m[6]:.0f} | "
            f"table total={m[8]:.0f} wait_free={m[9]:.0f} busy/tile={(m[8] - m[9]) / tiles:.0f} | epilogue total={m[10]:.0f} wait_acc_full={m[11]:.0f} "
            f"busy/tile={(m[10] - m[11]) / tiles:.0f}")


layers = ((96, 96, 1, 3), (128, 96, 1, 3), (32, 32, 2, 3), (64, 64, 4, 3), (128, 128, 8, 3), (256, 256, 16, 3), (384, 256, 8, 3), (32, 32, 1, 2), (64, 64, 2, 2))
for (cin, cout, ts, ks) in layers:
    km = mgr.kernel_map(ts, ks, 1 if ks == 3 else 2, False)
    n_in, n = km.n_in, km.n_out
    kv = km.kv
    x = torch.randn(n_in, cin, device=dev).to(torch.bfloat16)
    g = torch.randn(n, cout, device=dev).to(torch.bfloat16)
    w = torch.randn(kv, cin, cout, device=dev) * 0.05
    packed = ops.pack_weights(w, False, False)
    pairs = km.num_pairs()
    flops = 2.0 * pairs * cin * cout
    table, rows, masks = km.tc_table()
    variants = {"scan": lambda: ops.conv_forward(x, km.nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed),
                "sorted": lambda: ops.conv_forward(x, table, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed, out_rows=rows),
                "sorted+masks": lambda: ops.conv_forward(x, table, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed, out_rows=rows,
                                                         tile_masks=masks)}
    dw = torch.zeros_like(w)
    us_w = timeit(lambda: ops.conv_wgrad(x, g, km.pairs, kv, dw, math_mode=1))
    print(f"{cin}->{cout} ts{ts} K{ks}: n={n} pairs={pairs} density={pairs / (kv * n):.2f} | wgrad {us_w:.1f} us {flops / us_w / 1e6:.1f} TFLOP/s")
    for warps in (8, 16):
        _cabi.lib().gcd_set_option(_cabi.OPT_TC_WARPS, warps)
        for name, fn in variants.items():
            if warps == 16 and name == "scan":
                continue
            us = timeit(fn)
            print(f"    fwd {name:13s} warps={warps:2d} {us:7.1f} us  {flops / us / 1e6:6.1f} TFLOP/s alg")
            if profile:
                print("        " + counters(fn))
    _cabi.lib().gcd_set_option(_cabi.OPT_TC_WARPS, 8)
