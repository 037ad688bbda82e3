"""One launch of every kernel the round-2 roofline numbers quote, between cudaProfilerStart/Stop, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_full python tools/ncu_targets.py
Inputs are one kitti_b4 batch (conv / batch-norm / tile-sort targets) and one dense ~1.2 M-point scan (hash / kernel-map /
pair-list / gather targets), i.e. the shapes bench.py and tools/bench_maps.py time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import torch

import bench
import gcdlss_b200
import MinkowskiEngine as ME
from gcdlss_b200 import _cabi, ops, synth

gcdlss_b200.set_math_mode("bf16")
gcdlss_b200.set_tile_sort(True)
dev = torch.device("cuda:0")
hb = bench.make_host_batches("kitti", 4, None, 17, 0, 1)
bc, f, l = bench.quantize_batch_on_gpu(hb[0], 0.05, dev)
mgr = ME.SparseTensor(features=f, coordinates=bc).coordinate_manager
targets = []


def conv_targets(cin, cout, ts):
    km = mgr.kernel_map(ts, 3, 1, False)
    n = km.n_out
    x = torch.randn(n, cin, device=dev).to(torch.bfloat16)
    g = torch.randn(n, cout, device=dev).to(torch.bfloat16)
    w = torch.randn(27, cin, cout, device=dev) * 0.05
    packed = ops.pack_weights(w, False, False)
    table, rows, masks = km.tc_table()
    dw = torch.zeros_like(w)
    pairs = km.pairs
    targets.append(lambda: ops.conv_forward(x, table, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=packed, out_rows=rows, tile_masks=masks))
    targets.append(lambda: ops.conv_wgrad(x, g, pairs, 27, dw, math_mode=1))


conv_targets(96, 96, 1)
conv_targets(256, 256, 16)
conv_targets(32, 32, 2)
# batch norm on the largest activation of the step ([n, 96] bf16): the fused two-phase forms run inside gcd_block_*; here the
# stand-alone passes (same device code, two launches) plus one fused block forward / backward
n1 = mgr.get_map(1).n
xb = torch.randn(n1, 96, device=dev).to(torch.bfloat16)
gam, bet, rm, rv = torch.ones(96, device=dev), torch.zeros(96, device=dev), torch.zeros(96, device=dev), torch.ones(96, device=dev)
yb, mean, invstd = ops.bn_forward(xb, gam, bet, rm, rv, True, 0.1, 1e-5, True, None)
dyb = torch.randn(n1, 96, device=dev).to(torch.bfloat16)
targets.append(lambda: ops.bn_forward(xb, gam, bet, rm, rv, True, 0.1, 1e-5, True, None))
targets.append(lambda: ops.bn_backward(dyb, xb, yb, mean, invstd, gam, True, True, False))
blk = ME.MinkowskiConvolution(96, 96, kernel_size=3, dimension=3).to(dev)
bnm = ME.MinkowskiBatchNorm(96).to(dev).train()
from gcdlss_b200.nn import conv_bn_act
from gcdlss_b200.sparse_tensor import CoordinateMapKey, SparseTensor


def fused_unit():
    xin = xb.clone().requires_grad_(True)
    out = conv_bn_act(blk, bnm, SparseTensor(xin, coordinate_map_key=CoordinateMapKey(1), coordinate_manager=mgr))
    out._F.float().sum().backward()


targets.append(fused_unit)
# tile sort of the largest table
nbr1 = mgr.kernel_map(1, 3, 1, False).nbr
targets.append(lambda: ops.kmap_tile_sort(nbr1))
# hash / kernel maps / pair lists / gather on a dense scan
xyz, feat = synth.make_dense_scan(0, sweeps=10)
pts = torch.from_numpy(xyz).to(dev)
ic = ops.quantize(pts, 0.05, 3, 0)
um, inv, _ = ops.unique_rows(ic, 0)
dbc = torch.cat([torch.zeros((um.shape[0], 1), dtype=torch.int32, device=dev), ic.index_select(0, um)], 1).contiguous()
status = torch.zeros(1, dtype=torch.int32, device=dev)
rt = ops.runtable_build(dbc, 1, status)
tb = ops.hash_build(dbc, status)
dn = ops.kmap_subm_runs(dbc, rt, 3, 1)
vf = torch.randn(um.shape[0], 96, device=dev)
targets += [lambda: ops.quantize(pts, 0.05, 3, 0), lambda: ops.unique_rows(ic, 0), lambda: ops.runtable_build(dbc, 1, status),
            lambda: ops.kmap_subm_runs(dbc, rt, 3, 1), lambda: ops.kmap_subm_runs(dbc, rt, 5, 1), lambda: ops.kmap_subm(dbc, tb, 3, 1),
            lambda: ops.coords_stride2(dbc, 1, status), lambda: ops.pairs_from_table(dn), lambda: ops.rows_gather(vf, inv)]


def coop():
    ops.set_option(_cabi.OPT_KMAP_COOP, 1)
    ops.kmap_subm(dbc, tb, 3, 1)
    ops.set_option(_cabi.OPT_KMAP_COOP, 0)


targets.append(coop)
which = os.environ.get("NCU_TARGETS", "all")      # "conv": the tensor-core kernels + batch norm + tile sort; "side": hashing / maps / gather
n_conv = 10
if which == "conv":
    targets = targets[:n_conv]
elif which == "side":
    targets = targets[n_conv:]
elif which == "final":       # end-of-round refresh: the 96->96 forward / wgrad, the 256-channel deep layer, one fused conv-BN unit forward + backward
    targets = [targets[0], targets[1], targets[2], targets[8]]
for t in targets:            # warm-up: module loading, allocator
    t()
    t()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for t in targets:
    t()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ncu targets done:", len(targets))
