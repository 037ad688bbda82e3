"""GPU diagnostic: re-compute every backward op of one MinkUNet step with plain torch fp64 ops on the
actual tensors the kernels received, and report the ops whose result disagrees."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _paths  # noqa
import numpy as np, torch
import gcdlss_b200, MinkowskiEngine as ME
from gcdlss_b200 import ops, synth, functional
from models import minkunet as mu
from oracle import quantize as oq

def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))

log = []
real_bn_bwd, real_conv_fwd, real_wgrad = ops.bn_backward, ops.conv_forward, ops.conv_wgrad

def bn_backward(dy, x, y, mean, invstd, gamma, relu, training, need_dres):
    dx, dres, dgamma, dbeta = real_bn_bwd(dy, x, y, mean, invstd, gamma, relu, training, need_dres)
    g = dy.double()
    if relu: g = g * (y.double() > 0)
    xhat = (x.double() - mean.double()) * invstd.double()
    n = x.shape[0]
    sg, sgx = g.sum(0), (g * xhat).sum(0)
    rdx = gamma.double() * invstd.double() * (g - sg / n - xhat * sgx / n)
    log.append(("bn_bwd", tuple(x.shape), relu, dy.is_contiguous(), dy.stride(), rel(dx, rdx), rel(dgamma, sgx), rel(dbeta, sg),
                rel(dres, g) if dres is not None else 0.0))
    return dx, dres, dgamma, dbeta

def conv_forward(inp, nbr, w3, n_out, *, transpose_w=False, mirror=False, bias=None, out_dtype=None, math_mode=0, w_packed=None, stats=None, out_rows=None):
    out = real_conv_fwd(inp, nbr, w3, n_out, transpose_w=transpose_w, mirror=mirror, bias=bias, out_dtype=out_dtype, math_mode=math_mode,
                        w_packed=w_packed, stats=stats, out_rows=out_rows)
    kv = w3.shape[0]
    ref = torch.zeros((n_out, out.shape[1]), dtype=torch.float64, device=inp.device)
    x = inp.double()
    for k in range(kv):
        wk = w3[kv - 1 - k if mirror else k].double()
        B = wk.t() if transpose_w else wk
        if nbr is None:
            ref += x @ B
        else:
            idx = nbr[k].long()
            o = torch.nonzero(idx >= 0).reshape(-1)
            ref.index_add_(0, o, x[idx[o]] @ B)
    if out_rows is not None: ref = torch.zeros_like(ref).index_copy_(0, out_rows.long(), ref)
    if bias is not None: ref += bias.double()
    log.append(("dgrad" if transpose_w else "fwd", tuple(inp.shape), tuple(w3.shape), inp.is_contiguous(), inp.stride(), rel(out, ref)))
    return out

def conv_wgrad(inp, gout, pairs, kv, dw, dbias=None, math_mode=0):
    before = dw.clone()
    real_wgrad(inp, gout, pairs, kv, dw, dbias=dbias, math_mode=math_mode)
    ref = torch.zeros_like(dw, dtype=torch.float64)
    if pairs is None:
        ref[0] = inp.double().t() @ gout.double()
    else:
        pi, po, off = pairs
        off = off.tolist()
        for k in range(kv):
            a, b = off[k], off[k + 1]
            ref[k] = inp.double()[pi[a:b].long()].t() @ gout.double()[po[a:b].long()]
    log.append(("wgrad", tuple(inp.shape), tuple(gout.shape), inp.is_contiguous(), gout.is_contiguous(), gout.stride(), rel(dw - before, ref)))

ops.bn_backward, ops.conv_forward, ops.conv_wgrad = bn_backward, conv_forward, conv_wgrad

gcdlss_b200.set_math_mode("fp32")
torch.manual_seed(1234)
coords, feats = [], []
for i in range(2):
    xyz, f = synth.make_scan("kitti", i, n_points=6000)
    c, um, _ = oq.sparse_quantize_me(xyz, 0.05)
    coords.append(c); feats.append(f[um])
bc = oq.batched_coordinates(coords); feats = np.concatenate(feats)
model = mu.MinkUNet14A(1, 17).cuda().train()
labels = torch.from_numpy(np.random.default_rng(0).integers(0, 17, bc.shape[0])).cuda()
st = ME.SparseTensor(features=torch.from_numpy(feats).cuda(), coordinates=torch.from_numpy(bc).cuda())
logits = model(st).F
n_fwd = len(log)
torch.nn.functional.cross_entropy(logits, labels).backward()
torch.cuda.synchronize()
print("forward ops:", n_fwd, "worst", max(l[-1] for l in log[:n_fwd]))
for l in log[n_fwd:]:
    worst = max(v for v in l if isinstance(v, float))
    print(("BAD " if worst > 1e-4 else "ok  "), l)
