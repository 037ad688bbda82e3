"""The whole-trunk fast path (functional.TrunkFunction -> gcd_run_ops) without a GPU: the real Python code builds its
arenas, block structures and operation program on CPU tensors, a host emulation of the C entry point (tests/emu_ops.py)
interprets that program through the raw pointers, and the result -- eight stage outputs, the gradient of the stem
activation and of every parameter -- is compared with a plain torch-autograd statement of the same U-Net on the oracle's
kernel maps.  This pins the pointer / leading-dimension / ordering logic (ME.cat as column copies, the split of the
concatenated gradient, skip gradients, external gradients on inner stages) before any GPU time is spent."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import emu_ops
from conftest import small_cloud
from oracle import conv as oc
from oracle import coords as ocd
from test_runs_mode_plumbing import runs_manager  # noqa: F401  (fixture: the real CoordinateManager on emulated kernels)
from test_emulated_kernels import emu  # noqa: F401


@pytest.fixture()
def trunk_env(runs_manager, monkeypatch):  # noqa: F811
    import gcdlss_b200
    from gcdlss_b200 import functional, ops
    CoordinateManager, _ = runs_manager
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "pairs_from_table",          # the library's lists are int32 throughout (offsets included)
                        lambda nbr: tuple(torch.from_numpy(np.ascontiguousarray(a, np.int32)) for a in ocd.pairs_from_table(nbr.numpy().T)))

    def run_ops(prog, n_ops, backward):        # stands for gcd_run_ops_exec (same program, same results; no second stream on the host)
        import ctypes as C
        launches = C.c_int32(0)
        emu_ops.run_ops(prog, n_ops, 0, C.byref(launches))
        return launches.value

    monkeypatch.setattr(ops, "run_ops", run_ops)
    prev = (gcdlss_b200.get_math_mode(), gcdlss_b200.get_tile_sort())
    gcdlss_b200.set_math_mode("fp32")
    gcdlss_b200.set_tile_sort(False)
    yield CoordinateManager
    gcdlss_b200.set_math_mode(prev[0])
    gcdlss_b200.set_tile_sort(prev[1])


def reference_trunk(model, x, lv):
    """Stages block1..block8 with torch autograd (ref models/minkunet.py:149-217, BasicBlock of ME's resnet_block)."""
    def bn(mod, t):
        return F.batch_norm(t, None, None, mod.bn.weight.double(), mod.bn.bias.double(), True, 0.1, mod.bn.eps)

    def block(blk, t, nbr3):
        out = torch.relu(bn(blk.norm1, oc.conv_table(t, nbr3, blk.conv1.kernel.double())))
        out = bn(blk.norm2, oc.conv_table(out, nbr3, blk.conv2.kernel.double()))
        res = t if blk.downsample is None else bn(blk.downsample[1], oc.conv_1x1(t, blk.downsample[0].kernel.double()))
        return torch.relu(out + res)

    names_e = (("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"), ("conv3p4s2", "bn3", "block3"), ("conv4p8s2", "bn4", "block4"))
    names_d = (("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"), ("convtr6p4s2", "bntr6", "block7"), ("convtr7p2s2", "bntr7", "block8"))
    skips, stages, cur = [x], [], x
    for i, (c, b, k) in enumerate(names_e):
        cur = torch.relu(bn(getattr(model, b), oc.conv_table(cur, lv.down(i), getattr(model, c).kernel.double())))
        for blk in getattr(model, k):
            cur = block(blk, cur, lv.subm(i + 1, 3))
        skips.append(cur)
        stages.append(cur)
    for i, (c, b, k) in enumerate(names_d):
        lvl = 3 - i
        cur = torch.relu(bn(getattr(model, b), oc.conv_table(cur, lv.up(lvl), getattr(model, c).kernel.double())))
        cur = torch.cat((cur, skips[lvl]), 1)
        for blk in getattr(model, k):
            cur = block(blk, cur, lv.subm(lvl, 3))
        stages.append(cur)
    return stages


@pytest.mark.parametrize("arch,taps", [("MinkUNet14A", ()), ("MinkUNet18A", (3, 5))])
def test_trunk_function_against_autograd(trunk_env, arch, taps):
    import gcdlss_b200
    from gcdlss_b200.nn import run_trunk, trunk_plan
    from gcdlss_b200.sparse_tensor import CoordinateMapKey, SparseTensor
    from models import minkunet as mu
    torch.manual_seed(0)
    model = getattr(mu, arch)(1, 17).train()
    bc = np.concatenate([small_cloud(51, 1500, spread=0.5, batch=0), small_cloud(52, 900, spread=0.5, batch=1)])
    lv = ocd.CoordLevels(bc)
    mgr = trunk_env(torch.from_numpy(bc))
    enc = [(getattr(model, c), getattr(model, b), getattr(model, k)) for c, b, k in mu._ENCODER]
    dec = [(getattr(model, c), getattr(model, b), getattr(model, k)) for c, b, k in mu._DECODER]
    plan = trunk_plan(enc, dec)
    assert plan is not None and len(plan.units) == len([m for m in model.modules() if isinstance(m, gcdlss_b200.MinkowskiConvolution)
                                                          or isinstance(m, gcdlss_b200.MinkowskiConvolutionTranspose)]) - 2     # all but stem and head
    x = torch.randn(bc.shape[0], 32).requires_grad_(True)
    stages = run_trunk(plan, SparseTensor(x, coordinate_map_key=CoordinateMapKey(1), coordinate_manager=mgr))
    assert stages is not None and [s.tensor_stride_int for s in stages] == [2, 4, 8, 16, 8, 4, 2, 1]
    g = torch.Generator().manual_seed(1)
    weights = {7: torch.randn(stages[7].F.shape, generator=g)}
    weights.update({t: torch.randn(stages[t].F.shape, generator=g) for t in taps})          # external gradients on inner stages too
    sum((stages[t].F * w).sum() for t, w in weights.items()).backward()

    xr = x.detach().double().requires_grad_(True)
    ref = reference_trunk(model, xr, lv)
    params = [p for p in model.parameters()]
    got_grads = [p.grad.clone() if p.grad is not None else None for p in params]
    for p in params:
        p.grad = None
    sum((ref[t] * w.double()).sum() for t, w in weights.items()).backward()
    for i in range(8):
        torch.testing.assert_close(stages[i].F.double(), ref[i].detach(), rtol=1e-4, atol=1e-5)
    scale = float(xr.grad.abs().max())
    torch.testing.assert_close(x.grad.double(), xr.grad, rtol=1e-3, atol=1e-4 * scale)
    checked = 0
    for (name, p), got in zip(model.named_parameters(), got_grads):
        if p.grad is None:
            assert got is None or name.startswith(("conv0p1s1", "bn0", "final")), name
            continue
        torch.testing.assert_close(got.double(), p.grad.double(), rtol=2e-3, atol=2e-4 * max(float(p.grad.abs().max()), 1e-6), msg=lambda m: f"{name}: {m}")
        checked += 1
    assert checked == 3 * len(plan.units)


def test_gradient_sink_plumbing(trunk_env, monkeypatch):
    """functional.TrunkFunction.backward with a gradient sink (ddp.GradBucketReducer on more than one GPU): the kernels -- here
    their host emulation -- accumulate straight into the reducer's flat buckets, autograd gets no gradient for the trunk's
    parameters, one event per network stage is recorded in the order the stages finish, and the reducer launches every bucket
    exactly once, behind the event of the last stage that touched it; the bucket that also holds a parameter in front of the
    trunk (the stem) leaves only when that parameter's own hook has fired.  The buckets must hold the gradients of the plain path."""
    import gcdlss_b200
    from gcdlss_b200 import ops
    from gcdlss_b200.ddp import GradBucketReducer
    from gcdlss_b200.nn import run_trunk, trunk_plan
    from gcdlss_b200.sparse_tensor import CoordinateMapKey, SparseTensor
    from models import minkunet as mu

    class FakeEvent:
        n = 0

        def __init__(self):
            FakeEvent.n += 1
            self.cuda_event = 1000 + FakeEvent.n

    torch.manual_seed(0)
    model = mu.MinkUNet14A(1, 17).train()
    enc = [(getattr(model, c), getattr(model, b), getattr(model, k)) for c, b, k in mu._ENCODER]
    dec = [(getattr(model, c), getattr(model, b), getattr(model, k)) for c, b, k in mu._DECODER]
    plan = trunk_plan(enc, dec)
    stem_w = torch.nn.Parameter(torch.ones(32))          # stands for the stem: in front of the trunk, its gradient arrives last
    head_w = torch.nn.Parameter(torch.ones(96))          # stands for the head: behind the trunk, its gradient arrives first
    bc = np.concatenate([small_cloud(61, 1200, spread=0.5, batch=0), small_cloud(62, 700, spread=0.5, batch=1)])
    x0 = torch.randn(bc.shape[0], 32)
    params = [stem_w] + plan.parameters() + [head_w]     # registration order of a model: stem first, head last

    def run():
        mgr = trunk_env(torch.from_numpy(bc))
        stages = run_trunk(plan, SparseTensor(x0 * stem_w, coordinate_map_key=CoordinateMapKey(1), coordinate_manager=mgr))
        out = stages[7].F * head_w
        (out * torch.linspace(-1, 1, out.numel()).view_as(out)).sum().backward()
        return torch.cat([p.grad.flatten().clone() for p in params])

    plain = run()
    for p in params:
        p.grad = None
    red = GradBucketReducer(params, bucket_bytes=1 << 20)        # a handful of buckets
    assert len(red.buckets) >= 3
    launched = []
    red.world = 2
    red._launch = lambda b, gate=None: launched.append((b, getattr(gate, "cuda_event", None))) or len(launched)
    events = {}
    monkeypatch.setattr(red, "stage_event", lambda i: events.setdefault(i, FakeEvent()))
    monkeypatch.setattr(ops, "SINK_ON_HOST", True)
    ops.set_grad_sink(red)
    emu_ops.recorded_events.clear()
    try:
        red.reset()
        sunk = run()
    finally:
        ops.set_grad_sink(None)
    # eight stage events, recorded in the order the stages finish
    assert len(emu_ops.recorded_events) == 8 and len(set(emu_ops.recorded_events)) == 8
    # every bucket went out exactly once; all but the one that holds the stem behind a stage event, in the order of those events
    assert sorted(b for b, _ in launched) == list(range(len(red.buckets))), launched
    stem_bucket = red._owner[stem_w]
    gated = [(b, e) for b, e in launched if e is not None]
    assert len(gated) == len(red.buckets) - 1 and all(e in emu_ops.recorded_events for _, e in gated)
    order = [emu_ops.recorded_events.index(e) for _, e in gated]
    assert order == sorted(order)
    assert dict(launched)[stem_bucket] is None and launched[-1][0] == stem_bucket, "the stem's bucket leaves last, from the stem's own hook"
    assert red._pending == [0] * len(red.buckets)
    # same gradients as the plain path (the buckets ARE the .grad tensors)
    torch.testing.assert_close(sunk, plain, rtol=1e-4, atol=1e-6 * float(plain.abs().max()))
