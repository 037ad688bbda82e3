"""CUDA sparse convolution / BN / devoxelise vs the oracle (tolerances in gpu_util.py)."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from conftest import small_cloud
from gpu_util import TOL_BF16, TOL_FP32, rel_err
from oracle import conv as oc
from oracle import coords as ocd

pytestmark = pytest.mark.gpu


@pytest.fixture()
def math_mode():
    import gcdlss_b200
    prev = gcdlss_b200.get_math_mode()
    yield gcdlss_b200.set_math_mode
    gcdlss_b200.set_math_mode(prev)


def _scene(seed=0, n=6000, spread=0.8):
    bc = np.concatenate([small_cloud(seed, n, spread, batch=0), small_cloud(seed + 1, n // 2, spread, batch=1)])
    return bc, ocd.CoordLevels(bc)


def _run_layer(layer, bc, x_np, g_np, level_stride=1):
    import MinkowskiEngine as ME
    from gcdlss_b200.sparse_tensor import CoordinateMapKey
    base = ME.SparseTensor(features=torch.zeros(bc.shape[0], 1).cuda(), coordinates=torch.from_numpy(bc).cuda())
    mgr = base.coordinate_manager
    mgr.get_map(level_stride)
    x = torch.from_numpy(x_np).float().cuda().requires_grad_(True)
    st = ME.SparseTensor(x, coordinate_map_key=CoordinateMapKey(level_stride), coordinate_manager=mgr)
    out = layer(st)
    y = out.F
    (y * torch.from_numpy(g_np).float().cuda()).sum().backward()
    return y.detach().cpu(), x.grad.cpu(), out


CASES = [
    # kind, K, stride, Cin, Cout, bias, level_stride
    ("conv", 3, 1, 32, 64, False, 1),
    ("conv", 3, 1, 96, 96, False, 2),
    ("conv", 3, 1, 384, 256, False, 4),
    ("conv", 5, 1, 1, 32, False, 1),
    ("conv", 5, 1, 4, 32, False, 1),
    ("conv", 2, 2, 32, 32, False, 1),
    ("conv", 2, 2, 128, 128, False, 4),
    ("convtr", 2, 2, 256, 128, False, 4),
    ("convtr", 2, 2, 96, 96, False, 2),
    ("conv", 1, 1, 96, 17, True, 1),
    ("conv", 1, 1, 64, 128, False, 2),
]


def _oracle_table(lv, kind, K, stride, level):
    if K == 1:
        n = lv.coords[level].shape[0]
        return np.arange(n, dtype=np.int32)[:, None], n
    if stride == 1:
        return lv.subm(level, K), lv.coords[level].shape[0]
    if kind == "conv":
        return lv.down(level), lv.coords[level + 1].shape[0]
    return lv.up(level - 1), lv.coords[level - 1].shape[0]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}K{c[1]}s{c[2]}_{c[3]}to{c[4]}")
def test_layer_forward_backward(cuda, math_mode, mode, case):
    import MinkowskiEngine as ME
    kind, K, stride, cin, cout, bias, ts = case
    math_mode(mode)
    tol = TOL_FP32 if mode == "fp32" else TOL_BF16
    bc, lv = _scene()
    level = int(np.log2(ts))
    table, n_out = _oracle_table(lv, kind, K, stride, level)
    n_in = lv.coords[level].shape[0]
    rng = np.random.default_rng(1)
    x = rng.normal(0, 1, (n_in, cin)).astype(np.float32)
    g = rng.normal(0, 1, (n_out, cout)).astype(np.float32)
    torch.manual_seed(0)
    cls = ME.MinkowskiConvolution if kind == "conv" else ME.MinkowskiConvolutionTranspose
    layer = cls(cin, cout, kernel_size=K, stride=stride, bias=bias, dimension=3).cuda()
    y, dx, out = _run_layer(layer, bc, x, g, ts)
    assert out.tensor_stride_int == (ts * stride if kind == "conv" else ts // stride)
    # oracle in fp64
    xo = torch.from_numpy(x).double().requires_grad_(True)
    wo = layer.kernel.detach().cpu().double().reshape(K ** 3, cin, cout).requires_grad_(True)
    bo = layer.bias.detach().cpu().double().requires_grad_(True) if bias else None
    yo = oc.conv_table(xo, table, wo, bo)
    (yo * torch.from_numpy(g).double()).sum().backward()
    errs = {"y": rel_err(y, yo.detach()), "dx": rel_err(dx, xo.grad), "dw": rel_err(layer.kernel.grad.reshape(K ** 3, cin, cout), wo.grad)}
    if bias:
        errs["db"] = rel_err(layer.bias.grad, bo.grad)
    print(mode, case, errs)
    for k, v in errs.items():
        assert v < tol, (k, v)


# c = 96: the 16-byte-wide kernels (what a MinkUNet runs); c = 20: the 8-byte "vec" generation in bf16 (20 % 8 != 0), still wide in fp32;
# c = 17: the element-wise generation in both (the fallback chain of csrc/bn.cu, bn.cu:gcd_bn_*)
@pytest.mark.parametrize("c", [96, 20, 17])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("training,relu,residual", [(True, True, True), (True, False, False), (False, True, False), (True, True, False)])
def test_batchnorm_fused(cuda, math_mode, mode, training, relu, residual, c):
    import MinkowskiEngine as ME
    math_mode(mode)
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    tol = TOL_FP32 if mode == "fp32" else TOL_BF16
    bc, _ = _scene(3, 2000)
    n = bc.shape[0]
    rng = np.random.default_rng(2)
    x_np = (rng.normal(0.3, 2.0, (n, c))).astype(np.float32)
    r_np = rng.normal(0, 1, (n, c)).astype(np.float32)
    g_np = rng.normal(0, 1, (n, c)).astype(np.float32)
    bn = ME.MinkowskiBatchNorm(c).cuda()
    with torch.no_grad():
        bn.bn.weight.uniform_(0.5, 1.5); bn.bn.bias.uniform_(-0.5, 0.5)
        bn.bn.running_mean.uniform_(-0.2, 0.2); bn.bn.running_var.uniform_(0.5, 2.0)
    ref = torch.nn.BatchNorm1d(c).double()
    ref.load_state_dict({k: v.detach().cpu().double() if v.is_floating_point() else v.cpu() for k, v in bn.bn.state_dict().items()})
    bn.train(training); ref.train(training)
    base = ME.SparseTensor(features=torch.zeros(n, 1).cuda(), coordinates=torch.from_numpy(bc).cuda())
    xq = torch.from_numpy(x_np).cuda().to(dt)
    rq = torch.from_numpy(r_np).cuda().to(dt)
    x = xq.clone().requires_grad_(True)
    r = rq.clone().requires_grad_(True)
    y = bn(base._like(x), relu=relu, residual=base._like(r) if residual else None)._F
    (y.float() * torch.from_numpy(g_np).cuda()).sum().backward()
    xo = xq.double().cpu().requires_grad_(True)
    ro = rq.double().cpu().requires_grad_(True)
    yo = ref(xo) + (ro if residual else 0)
    yo = torch.relu(yo) if relu else yo
    go = torch.from_numpy(g_np).double()
    if mode == "bf16":
        go = go.to(torch.bfloat16).double()       # the incoming gradient is rounded to the storage dtype
    (yo * go).sum().backward()
    errs = {"y": rel_err(y, yo.detach()), "dx": rel_err(x.grad, xo.grad), "dgamma": rel_err(bn.bn.weight.grad, ref.weight.grad),
            "dbeta": rel_err(bn.bn.bias.grad, ref.bias.grad)}
    if residual:
        errs["dres"] = rel_err(r.grad, ro.grad)
    if training:
        errs["running_mean"] = rel_err(bn.bn.running_mean, ref.running_mean)
        errs["running_var"] = rel_err(bn.bn.running_var, ref.running_var)
        assert int(bn.state_dict()["bn.num_batches_tracked"]) == 1      # the counter is flushed when the state_dict is read
    print(mode, training, relu, residual, errs)
    for k, v in errs.items():
        assert v < (tol if k not in ("running_mean", "running_var") else 1e-5), (k, v)


def test_relu_and_cat(cuda, math_mode):
    import MinkowskiEngine as ME
    math_mode("fp32")
    bc, _ = _scene(5, 500)
    n = bc.shape[0]
    x = torch.randn(n, 8).cuda().requires_grad_(True)
    st = ME.SparseTensor(features=x, coordinates=torch.from_numpy(bc).cuda())
    y = ME.MinkowskiReLU(inplace=True)(st)
    z = ME.cat(y, st)
    assert z.F.shape == (n, 16) and torch.equal(z.F[:, :8], torch.relu(x)) and torch.equal(z.C.cpu(), torch.from_numpy(bc))
    z.F.sum().backward()
    assert torch.equal(x.grad, (x > 0).float() + 1)
    s = st + y
    s += st
    assert torch.allclose(s.F, 2 * x + torch.relu(x))


def test_devoxelize_and_point_reduce(cuda):
    from gcdlss_b200 import devoxelize, voxelize_reduce
    rng = np.random.default_rng(0)
    m, n, c = 700, 5000, 20
    inv = torch.from_numpy(rng.integers(0, m, n))
    inv[:m] = torch.arange(m)                                   # every voxel owns at least one point
    vox = torch.randn(m, c).cuda().requires_grad_(True)
    pts = devoxelize(vox, inv)
    g = torch.randn(n, c).cuda()
    (pts * g).sum().backward()
    vo = vox.detach().cpu().double().requires_grad_(True)
    po = oc.devox_gather(vo, inv.numpy())
    (po * g.cpu().double()).sum().backward()
    assert torch.equal(pts.detach().cpu(), po.detach().float())  # a gather is exact
    assert rel_err(vox.grad, vo.grad) < 1e-6
    pf = torch.randn(n, c).cuda()
    for mode in ("mean", "max", "sum"):
        out = voxelize_reduce(pf, inv, m, mode)
        ref = oc.point_to_voxel(pf.cpu().double(), inv.numpy(), m, mode)
        assert rel_err(out, ref) < 1e-6, mode
