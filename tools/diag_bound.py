"""Which resource bounds conv_fwd_tc? Same launch with three neighbour tables: real, one valid row per (tile, offset), all rows valid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import torch
import gcdlss_b200
from gcdlss_b200 import ops

gcdlss_b200.set_math_mode("bf16")
dev = torch.device("cuda:0")

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

for (cin, cout, n) in ((96, 96, 175581), (32, 32, 90524), (256, 256, 3944), (256, 256, 148 * 128 * 2)):
    x = torch.randn(n, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(27, cin, cout, device=dev) * 0.05
    packed = ops.pack_weights(w, False, False)
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = torch.randint(0, n, (27, n), device=dev, dtype=torch.int32, generator=g)
    sparse1 = torch.full((27, n), -1, device=dev, dtype=torch.int32)
    sparse1[:, ::128] = rnd[:, ::128]
    p25 = torch.where(torch.rand((27, n), device=dev, generator=g) < 0.25, rnd, torch.full_like(rnd, -1))
    tiles = (n + 127) // 128
    for name, nbr in (("one valid row per tile+offset", sparse1), ("25% valid, random rows", p25), ("all valid, random rows", rnd)):
        us = timeit(lambda: ops.conv_forward(x, nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed))
        iters = -(-tiles // 148) * 27 * ((cin + 63) // 64)
        print(f"{cin}->{cout} n={n} ({tiles} tiles) {name:32s}: {us:7.1f} us  ~{us * 1.92e3 / iters:6.0f} cycles per stage iteration (ceil(tiles/148) tiles per CTA)")
