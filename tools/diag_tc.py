"""GPU diagnostic: isolated timing of the tcgen05 conv kernels on representative MinkUNet layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import numpy as np, torch
import gcdlss_b200, MinkowskiEngine as ME
from gcdlss_b200 import ops, _cabi
import bench

gcdlss_b200.set_math_mode("bf16")
dev = torch.device("cuda:0")
hb = bench.make_host_batches("kitti", 4, None, 17, 0, 1)
bc, f, l = bench.quantize_batch_on_gpu(hb[0], 0.05, dev)
st = ME.SparseTensor(features=f, coordinates=bc)
mgr = st.coordinate_manager

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

for (cin, cout, ts) in ((96, 96, 1), (128, 96, 1), (32, 32, 2), (64, 64, 4), (128, 128, 8), (256, 256, 16), (384, 256, 8)):
    km = mgr.kernel_map(ts, 3, 1, False)
    n = km.n_out
    x = torch.randn(n, cin, device=dev).to(torch.bfloat16)
    g = torch.randn(n, cout, device=dev).to(torch.bfloat16)
    w = torch.randn(27, cin, cout, device=dev) * 0.05
    packed = ops.pack_weights(w, False, False)
    pairs = km.num_pairs()
    flops = 2.0 * pairs * cin * cout
    us_f = timeit(lambda: ops.conv_forward(x, km.nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed))
    dw = torch.zeros_like(w)
    us_w = timeit(lambda: ops.conv_wgrad(x, g, km.pairs, 27, dw, math_mode=1))
    import ctypes as C
    from gcdlss_b200 import _cabi
    if hasattr(_cabi.lib(), "gcd_debug_set_buffer"):
        dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
        _cabi.lib().gcd_debug_set_buffer(C.c_void_p(dbg.data_ptr()))
        ops.conv_forward(x, km.nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed)
        torch.cuda.synchronize()
        _cabi.lib().gcd_debug_set_buffer(None)
        d = dbg.view(148, 8).double(); d = d[d[:, 0] > 0].max(0).values
        print(f"   [profile build] producer: total={d[0]:.0f} table={d[1]:.0f} wait_empty={d[2]:.0f} iters={d[3]:.0f} busy/iter={(d[0]-d[1]-d[2])/max(d[3],1):.0f}"
              f" | mma: total={d[4]:.0f} wait_full={d[5]:.0f} wait_acc={d[6]:.0f} busy/iter={(d[4]-d[5]-d[6])/max(d[3],1):.0f}")
    print(f"{cin}->{cout} ts{ts}: n={n} pairs={pairs} density={pairs/(27*n):.2f} | fwd {us_f:.1f} us {flops/us_f/1e6:.1f} TFLOP/s alg "
          f"({27*n*cin*cout*2/us_f/1e6:.0f} dense-equiv) | wgrad {us_w:.1f} us {flops/us_w/1e6:.1f} TFLOP/s")
