"""gcd_block_forward / gcd_block_backward (blocks sequenced in C) against the per-launch Python path.

Both paths launch the same kernels in the same order, so everything except the fp32-atomic weight gradients is
bit-identical; the per-launch path is itself checked against the oracle in test_gpu_conv / test_gpu_minkunet."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from oracle import quantize as oq

pytestmark = pytest.mark.gpu


def _tensor(c, mode, n_points=5000, seed=0):
    import MinkowskiEngine as ME
    from gcdlss_b200 import synth
    coords = []
    for i in range(2):
        xyz, _ = synth.make_scan("kitti", i + seed, n_points=n_points)
        cq, _, _ = oq.sparse_quantize_me(xyz, 0.1)
        coords.append(cq)
    bc = oq.batched_coordinates(coords)
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(bc.shape[0], c, generator=g)
    return bc, f


def _run(module_fn, bc, f, fused, mode, up_from=None):
    import gcdlss_b200
    import MinkowskiEngine as ME
    from gcdlss_b200 import functional as fn
    gcdlss_b200.set_math_mode(mode)
    fn._FUSED_C = fused
    try:
        feats = f.cuda().requires_grad_(True)
        st = ME.SparseTensor(features=feats, coordinates=torch.from_numpy(bc).cuda())
        out, params = module_fn(st)
        g = torch.Generator().manual_seed(7)
        go = torch.randn(out.F.shape, generator=g).cuda()
        (out.F * go).sum().backward()
        res = {"out": out.F.detach().float().cpu(), "dx": feats.grad.detach().float().cpu()}
        for k, p in params.items():
            res["d" + k] = p.grad.detach().float().cpu()
            p.grad = None
        return res
    finally:
        fn._FUSED_C = True
        gcdlss_b200.set_math_mode("fp32")


def _compare(a, b, mode):
    for k in a:
        assert a[k].shape == b[k].shape, k
        if k.startswith("d") and k.endswith("kernel"):        # fp32 atomics: order of accumulation differs run to run
            err = float((a[k] - b[k]).abs().max() / max(float(b[k].abs().max()), 1e-30))
            assert err < 1e-4, (k, err)
        elif k != "out" and k != "dx" and ("norm" in k or ".bn." in k or k.startswith("dbn")):
            # batch-norm parameter gradients are sums over rows: the C-sequenced path's two-phase kernel (for conv -> BN -> ReLU
            # units the variant that re-derives the ReLU mask from x) and the stand-alone reduction are different kernels with
            # their own fma contraction, so the fp32 sums may differ in the last bits.  (dx stays bit-identical in bf16: the
            # masks are the same.)
            err = float((a[k] - b[k]).abs().max() / max(float(b[k].abs().max()), 1e-30))
            assert err < 2e-6, (k, err)
        elif mode == "fp32" and k != "out":
            # backward: the C-sequenced path runs the two-phase batch-norm kernel, the per-launch path the two stand-alone passes:
            # the same source, but ptxas decides mul/add -> fma contraction per kernel, so fp32 values may differ in the last
            # bit (bf16 storage rounds that away)
            err = float((a[k] - b[k]).abs().max() / max(float(b[k].abs().max()), 1e-30))
            assert err < 2e-6, (k, err)
        elif k == "dx" and mode == "bf16":
            # The C-sequenced path re-derives the ReLU mask of conv -> BN -> ReLU units from x, the per-launch path reads the
            # stored activation: the masks are the same, so all but a stray element agree bit for bit; where the two kernels'
            # fp32 arithmetic lands on different sides of a bf16 rounding boundary an element may differ by one bf16 ulp.
            diff = (a[k] - b[k]).abs()
            bad = diff > 0
            assert float(bad.float().mean()) < 1e-3, (k, float(bad.float().mean()))
            assert float(diff.max()) <= 2.0 ** -7 * float(b[k].abs().max()), (k, float(diff.max()))
        else:
            assert torch.equal(a[k], b[k]), (k, float((a[k] - b[k]).abs().max()))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("inplanes,planes", [(32, 32), (48, 32), (64, 96)])
def test_basic_block(cuda, mode, inplanes, planes):
    import MinkowskiEngine as ME
    from MinkowskiEngine.modules.resnet_block import BasicBlock
    import torch.nn as nn
    torch.manual_seed(3)
    ds = None
    if inplanes != planes:
        ds = nn.Sequential(ME.MinkowskiConvolution(inplanes, planes, kernel_size=1, stride=1, dimension=3), ME.MinkowskiBatchNorm(planes))
    block = BasicBlock(inplanes, planes, downsample=ds, dimension=3).cuda().train()
    pre = ME.MinkowskiConvolution(inplanes, inplanes, kernel_size=1, dimension=3).cuda()    # makes the block input need a gradient in the working dtype
    state = {k: v.clone() for k, v in block.state_dict().items()}
    bc, f = _tensor(inplanes, mode)

    def fn(st):
        block.load_state_dict(state)
        return block(pre(st)), {k: p for k, p in block.named_parameters()}
    a = _run(fn, bc, f, True, mode)
    ra = {k: v.clone() for k, v in block.state_dict().items() if "running" in k}
    b = _run(fn, bc, f, False, mode)
    rb = {k: v.clone() for k, v in block.state_dict().items() if "running" in k}
    _compare(a, b, mode)
    for k in ra:
        assert torch.equal(ra[k], rb[k]), k


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("kind", ["down", "up", "same"])
def test_conv_bn_act(cuda, mode, kind):
    import MinkowskiEngine as ME
    from gcdlss_b200.nn import conv_bn_act
    torch.manual_seed(4)
    c_in, c_out = 32, 64
    down = ME.MinkowskiConvolution(c_in, c_in, kernel_size=2, stride=2, dimension=3).cuda()
    if kind == "down":
        conv = ME.MinkowskiConvolution(c_in, c_out, kernel_size=2, stride=2, dimension=3).cuda()
    elif kind == "up":
        conv = ME.MinkowskiConvolutionTranspose(c_in, c_out, kernel_size=2, stride=2, dimension=3).cuda()
    else:
        conv = ME.MinkowskiConvolution(c_in, c_out, kernel_size=3, stride=1, dimension=3).cuda()
    bn = ME.MinkowskiBatchNorm(c_out).cuda().train()
    pre = ME.MinkowskiConvolution(c_in, c_in, kernel_size=1, dimension=3).cuda()
    state = {k: v.clone() for k, v in bn.state_dict().items()}
    bc, f = _tensor(c_in, mode, seed=2)

    def fn(st):
        bn.load_state_dict(state)
        x = pre(st)
        if kind == "up":
            x = down(x)
        y = conv_bn_act(conv, bn, x, relu=(kind != "up"))
        return y, {"conv.kernel": conv.kernel, "bn.weight": bn.bn.weight, "bn.bias": bn.bn.bias}
    a = _run(fn, bc, f, True, mode)
    b = _run(fn, bc, f, False, mode)
    _compare(a, b, mode)


def test_fused_path_is_taken(cuda):
    """The C-sequenced path really runs (launch counter moves by the block's launch count in one call)."""
    import gcdlss_b200
    import MinkowskiEngine as ME
    from MinkowskiEngine.modules.resnet_block import BasicBlock
    from gcdlss_b200 import _cabi, ops
    gcdlss_b200.set_math_mode("fp32")
    block = BasicBlock(32, 32, dimension=3).cuda().train()
    bc, f = _tensor(32, "fp32")
    st = ME.SparseTensor(features=f.cuda(), coordinates=torch.from_numpy(bc).cuda())
    seen = []
    orig = _cabi._fn_cache.get("gcd_block_forward") or getattr(_cabi.lib(), "gcd_block_forward")

    def spy(*a):
        seen.append(1)
        return orig(*a)
    _cabi._fn_cache["gcd_block_forward"] = spy
    try:
        block(st)                                   # builds the kernel maps
        n0 = ops.launch_counter["calls"]
        block(st)
        per_unit = 2 if ops.get_option(_cabi.OPT_BN_FUSED) else 3       # convolution + batch norm (one two-phase launch, or two)
        assert seen and ops.launch_counter["calls"] - n0 == 2 * per_unit
    finally:
        _cabi._fn_cache["gcd_block_forward"] = orig


@pytest.mark.parametrize("mode,arch", [("bf16", "MinkUNet34C"), ("fp32", "MinkUNet34C"), ("bf16", "MinkUNet14A")])
def test_trunk_fast_path_equals_the_per_block_path(cuda, mode, arch):
    """functional.TrunkFunction (the whole trunk as one autograd node, gcd_run_ops) launches the same kernels in the same
    order as the per-block Functions: identical stage outputs bit for bit, gradients equal up to the summation order of the
    atomics in wgrad / the BN reductions."""
    import gcdlss_b200
    import MinkowskiEngine as ME
    from gcdlss_b200 import ops
    from models import minkunet as mu
    from oracle import quantize as oq
    from gcdlss_b200 import synth
    prev = gcdlss_b200.get_math_mode()
    gcdlss_b200.set_math_mode(mode)
    try:
        coords, feats = [], []
        for i in range(2):
            xyz, f = synth.make_scan("kitti", i, n_points=20000)
            c, um, _ = oq.sparse_quantize_me(xyz, 0.05)
            coords.append(c)
            feats.append(f[um])
        bc = torch.from_numpy(oq.batched_coordinates(coords)).cuda()
        f = torch.from_numpy(np.concatenate(feats)).cuda()
        labels = torch.randint(0, 17, (bc.shape[0],), device="cuda")
        torch.manual_seed(0)
        model = getattr(mu, arch)(1, 17).cuda().train()
        res = []
        for fast in (True, False):
            model.zero_grad(set_to_none=True)
            model.__dict__.pop("_trunk_plan", None)
            if not fast:
                model.__dict__["_trunk_plan"] = None
            n0 = ops.launch_counter["calls"]
            stages = model._trunk(ME.SparseTensor(features=f, coordinates=bc))
            logits = model.final(stages[7]).F
            loss = torch.nn.functional.cross_entropy(logits.float(), labels) + 0.01 * stages[3].F.float().square().mean()     # a tap on the bottleneck too
            loss.backward()
            res.append(([s.F.detach().clone() for s in stages], torch.cat([p.grad.flatten().float() for p in model.parameters()]),
                        ops.launch_counter["calls"] - n0))
        model.__dict__.pop("_trunk_plan", None)
        (s_fast, g_fast, n_fast), (s_slow, g_slow, n_slow) = res
        for a, b in zip(s_fast, s_slow):
            assert a.dtype == b.dtype and torch.equal(a, b)
        cos = float(torch.nn.functional.cosine_similarity(g_fast, g_slow, dim=0))
        print(mode, arch, "gradient cosine fast vs per-block", cos, "max abs diff", float((g_fast - g_slow).abs().max()), "launches", n_fast, n_slow)
        assert cos > 0.9999 and torch.isfinite(g_fast).all()
    finally:
        gcdlss_b200.set_math_mode(prev)
