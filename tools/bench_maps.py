"""HBM-side kernels of the hot path on BASELINE config 5 (dense ~1.2 M-point scans): quantise + dedup, coordinate
hash, stride-2 maps, kernel maps, pair lists, BN passes, devoxelise.  Prints achieved GB/s against the algorithmic
bytes of SURVEY 8(d) and the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import numpy as np, torch
import gcdlss_b200
from gcdlss_b200 import ops, synth, _cabi
from gcdlss_b200.coords import CoordinateManager

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(_paths.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(_paths.ROOT, "MEASURED_PEAKS.json")) else 6650.0

def timeit(fn, reps=20, flush=None):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None: flush.zero_()           # evict L2 between iterations (buffer > 126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
xyz, feat = synth.make_dense_scan(0, sweeps=10)
pts = torch.from_numpy(xyz).to(dev)
n = pts.shape[0]
rows = []
def report(name, seconds, nbytes):
    gbs = nbytes / seconds / 1e9
    rows.append((name, seconds * 1e6, nbytes / 1e6, gbs, gbs / peak))
    print(f"{name:38s} {seconds*1e6:9.1f} us  {nbytes/1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs/peak*100:5.1f}% of {peak:.0f}")

# quantise (12 B in, 12 B out per point)
t = timeit(lambda: ops.quantize(pts, 0.05, 3, 0), flush=flush)
report("quantize_kernel (floor, fp32)", t, n * 24)
ic = ops.quantize(pts, 0.05, 3, 0)
# dedup: N*(12 B coords + 8 B inverse) + M*(8 B unique idx) + table cap*12 B
t = timeit(lambda: ops.unique_rows(ic, 0), reps=10, flush=flush)
um, inv, table = ops.unique_rows(ic, 0)
m = um.shape[0]
report("unique_rows (hash insert+scan+emit)", t, n * (12 + 8) + m * 8 + table.cap * 12)
vox = ic.index_select(0, um)
bc = torch.cat([torch.zeros((m, 1), dtype=torch.int32, device=dev), vox], 1).contiguous()
status = torch.zeros(1, dtype=torch.int32, device=dev)
t = timeit(lambda: ops.hash_build(bc, status), flush=flush)
tb = ops.hash_build(bc, status)
report("hash_build (clear + insert)", t, m * 16 + tb.cap * 12)
t = timeit(lambda: ops.coords_stride2(bc, 1, status), reps=10, flush=flush)
coarse, parent, code, tb2 = ops.coords_stride2(bc, 1, status)
report("coords_stride2 (insert+scan+emit)", t, m * 16 + m * 8 + coarse.shape[0] * 16 + tb2.cap * 12)
for K in (3, 5):
    t = timeit(lambda: ops.kmap_subm(bc, tb, K, 1), reps=10, flush=flush)
    report(f"kmap_subm_kernel<{K}>", t, m * 16 + K**3 * m * 8 + K**3 * m * 4)
# warp-cooperative probing of the same point-wise table (GCD_OPT_KMAP_COOP: four lanes per voxel, one sector per probe)
ops.set_option(_cabi.OPT_KMAP_COOP, 1)
for K in (3, 5):
    t = timeit(lambda: ops.kmap_subm(bc, tb, K, 1), reps=10, flush=flush)
    report(f"kmap_subm_coop_kernel<{K}> (warp-coop)", t, m * 16 + K**3 * m * 8 + K**3 * m * 4)
ops.set_option(_cabi.OPT_KMAP_COOP, 0)
# the same tables searched in a run table (csrc/runtable.cuh, GCDLSS_KMAP=runs); same algorithmic bytes
t = timeit(lambda: ops.runtable_build(bc, 1, status), flush=flush)
rt = ops.runtable_build(bc, 1, status)
report("runtable_build (clear + insert)", t, m * 16 + rt.cap * 32)
for K in (3, 5):
    t = timeit(lambda: ops.kmap_subm_runs(bc, rt, K, 1), reps=10, flush=flush)
    report(f"kmap_runs_kernel<{K}>", t, m * 16 + K**3 * m * 8 + K**3 * m * 4)
    assert torch.equal(ops.kmap_subm_runs(bc, rt, K, 1), ops.kmap_subm(bc, tb, K, 1))
nbr = ops.kmap_subm(bc, tb, 3, 1)
t = timeit(lambda: ops.pairs_from_table(nbr), reps=10, flush=flush)
p = int(ops.pairs_from_table(nbr)[2][-1])
report("pairs_from_table (flag+scan+emit)", t, 27 * m * 4 + p * 8)
# BN passes and devox on a [m, 96] bf16 / fp32 tensor
for dt, s in ((torch.bfloat16, 2), (torch.float32, 4)):
    x = torch.randn(m, 96, device=dev).to(dt)
    g = torch.ones(96, device=dev); b = torch.zeros(96, device=dev); rm = torch.zeros(96, device=dev); rv = torch.ones(96, device=dev)
    st = torch.zeros(192, dtype=torch.float64, device=dev)
    t = timeit(lambda: _cabi.call("gcd_bn_stats", x.data_ptr(), 96, m, 96, 1 if s == 2 else 0, st.data_ptr(), ops._stream()), flush=flush)
    report(f"bn_stats_kernel ({'bf16' if s == 2 else 'fp32'}, C=96)", t, m * 96 * s)
    t = timeit(lambda: ops.bn_forward(x, g, b, rm, rv, True, 0.1, 1e-5, True, None, stats=st), flush=flush)
    report(f"bn_apply_train_kernel ({'bf16' if s == 2 else 'fp32'})", t, 2 * m * 96 * s)
    y, mean, invstd = ops.bn_forward(x, g, b, rm, rv, True, 0.1, 1e-5, True, None, stats=st)
    dy = torch.randn(m, 96, device=dev).to(dt)
    t = timeit(lambda: ops.bn_backward(dy, x, y, mean, invstd, g, True, True, False), flush=flush)
    report(f"bn_backward reduce+apply ({'bf16' if s == 2 else 'fp32'})", t, 7 * m * 96 * s)
vf = torch.randn(m, 96, device=dev)
t = timeit(lambda: ops.rows_gather(vf, inv), flush=flush)
report("rows_gather (devoxelise, C=96 fp32)", t, n * (8 + 96 * 4) + m * 96 * 4)
print(f"points={n} voxels={m} pairs3={p}")
json.dump([dict(kernel=r[0], us=r[1], mb=r[2], gbs=r[3], frac=r[4]) for r in rows], open("gpurun_out/maps_roofline.json", "w"), indent=1)
