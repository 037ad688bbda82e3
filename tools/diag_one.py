"""Runs one tcgen05 forward layer a few times (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import torch
import gcdlss_b200, MinkowskiEngine as ME
from gcdlss_b200 import ops
import bench
gcdlss_b200.set_math_mode("bf16")
dev = torch.device("cuda:0")
hb = bench.make_host_batches("kitti", 4, None, 17, 0, 1)
bc, f, l = bench.quantize_batch_on_gpu(hb[0], 0.05, dev)
st = ME.SparseTensor(features=f, coordinates=bc)
cin, cout, ts = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (96, 96, 1)))
km = st.coordinate_manager.kernel_map(ts, 3, 1, False)
n = km.n_out
x = torch.randn(n, cin, device=dev).to(torch.bfloat16)
w = torch.randn(27, cin, cout, device=dev) * 0.05
packed = ops.pack_weights(w, False, False)
for _ in range(4):
    ops.conv_forward(x, km.nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed)
g = torch.randn(n, cout, device=dev).to(torch.bfloat16)
dw = torch.zeros_like(w)
pairs = km.pairs
for _ in range(3):
    ops.conv_wgrad(x, g, pairs, 27, dw, math_mode=1)
torch.cuda.synchronize()
print("ok")
