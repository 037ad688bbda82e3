"""Per-tile fixed cost F and per-iteration cost c of conv_fwd_tc: same rows, kernel volumes 27 / 9 / 3 / 1 (all rows valid)."""
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

for (cin, cout, n) in ((96, 96, 148 * 128 * 8), (32, 32, 148 * 128 * 8), (64, 64, 148 * 128 * 8), (128, 128, 148 * 128 * 4), (256, 256, 148 * 128 * 2)):
    x = torch.randn(n, cin, device=dev).to(torch.bfloat16)
    g = torch.Generator(device=dev).manual_seed(0)
    res = []
    for kv in (27, 9, 3, 1):
        w = torch.randn(kv, cin, cout, device=dev) * 0.05
        packed = ops.pack_weights(w, False, False)
        nbr = torch.randint(0, n, (kv, n), device=dev, dtype=torch.int32, generator=g)
        us = timeit(lambda: ops.conv_forward(x, nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed))
        tiles_per_cta = n // 128 // 148
        res.append((kv, us * 1.92e3 / tiles_per_cta))
    nq = (cin + 63) // 64
    c = (res[0][1] - res[1][1]) / (18 * nq)
    F = res[0][1] - 27 * nq * c
    print(f"{cin}->{cout}: cycles per tile at kv=27/9/3/1: " + " ".join(f"{v:8.0f}" for _, v in res) + f"  => per-iteration c = {c:.0f}, per-tile fixed F = {F:.0f}")
