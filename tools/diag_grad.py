"""GPU diagnostic: per-parameter gradient error of MinkUNet14A vs the oracle, plus composite checks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import numpy as np, torch
import gcdlss_b200, MinkowskiEngine as ME
from models import minkunet as mu
from oracle import quantize as oq, conv as oc, coords as ocd
from oracle.minkunet import OracleMinkUNet
from gcdlss_b200 import synth

def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))

gcdlss_b200.set_math_mode("fp32")
torch.manual_seed(1234)
coords, feats = [], []
for i in range(2):
    xyz, f = synth.make_scan("kitti", i, n_points=6000)
    c, um, _ = oq.sparse_quantize_me(xyz, 0.05)
    coords.append(c); feats.append(f[um])
bc = oq.batched_coordinates(coords); feats = np.concatenate(feats)
model = mu.MinkUNet14A(1, 17).cuda().train()
params = {k: (v.detach().cpu().double() if v.is_floating_point() else v.cpu().clone()) for k, v in model.state_dict().items()}
for k in params:
    if params[k].is_floating_point() and "running" not in k: params[k].requires_grad_(True)
labels = torch.from_numpy(np.random.default_rng(0).integers(0, 17, bc.shape[0]))
st = ME.SparseTensor(features=torch.from_numpy(feats).cuda(), coordinates=torch.from_numpy(bc).cuda())
logits = model(st).F
torch.nn.functional.cross_entropy(logits, labels.cuda()).backward()
lo, _, lv = OracleMinkUNet(params, "MinkUNet14A", training=True).forward(bc, torch.from_numpy(feats).double())
torch.nn.functional.cross_entropy(lo, labels).backward()
errs = sorted(((rel(p.grad, params[n].grad), n) for n, p in model.named_parameters()), reverse=True)
print("ordered by model definition:")
for n, p in model.named_parameters():
    print(f"  {n:40s} {rel(p.grad, params[n].grad):.3e}  |g|max={float(params[n].grad.abs().max()):.3e}")

# composite: BasicBlock alone
print("--- BasicBlock (identity shortcut) and with downsample, vs oracle")
lvl = ocd.CoordLevels(bc)
for cin, cout in ((32, 32), (48, 32)):
    torch.manual_seed(0)
    ds = None
    if cin != cout:
        ds = torch.nn.Sequential(ME.MinkowskiConvolution(cin, cout, kernel_size=1, dimension=3), ME.MinkowskiBatchNorm(cout))
    blk = ME.modules.resnet_block.BasicBlock(cin, cout, downsample=ds, dimension=3).cuda().train()
    x = torch.randn(bc.shape[0], cin).cuda().requires_grad_(True)
    g = torch.randn(bc.shape[0], cout).cuda()
    s = ME.SparseTensor(features=x, coordinates=torch.from_numpy(bc).cuda())
    y = blk(s).F
    (y * g).sum().backward()
    P = {k: (v.detach().cpu().double() if v.is_floating_point() else v.cpu().clone()) for k, v in blk.state_dict().items()}
    for k in P:
        if P[k].is_floating_point() and "running" not in k: P[k].requires_grad_(True)
    P = {"b." + k: v for k, v in P.items()}
    om = OracleMinkUNet(P, "MinkUNet14A", training=True)
    xo = x.detach().cpu().double().requires_grad_(True)
    yo = om._block(xo, "b", lvl.subm(0, 3))
    (yo * g.cpu().double()).sum().backward()
    print(f"  block {cin}->{cout}: y {rel(y.detach(), yo.detach()):.2e} dx {rel(x.grad, xo.grad):.2e}",
          {n: f"{rel(p.grad, P['b.' + n].grad):.1e}" for n, p in blk.named_parameters()})
