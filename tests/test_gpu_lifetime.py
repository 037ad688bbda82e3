"""Backward after the caller's SparseTensors are gone.

A Lightning ``training_step`` builds the SparseTensor locally and returns only the loss (ref modules/exp.py:249-267);
by the time ``backward()`` runs, the tensors and their coordinate manager have been collected.  Every autograd node must
therefore own the tables its backward reads (forward table, dgrad table, pair lists, tile-sorted copies)."""
import gc

import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from gpu_util import TOL_BF16, TOL_FP32, rel_err
from oracle import quantize as oq
from oracle.minkunet import OracleMinkUNet

pytestmark = pytest.mark.gpu


def _batch(n_points=6000):
    from gcdlss_b200 import synth
    coords, feats = [], []
    for i in range(2):
        xyz, f = synth.make_scan("kitti", i, n_points=n_points)
        c, um, _ = oq.sparse_quantize_me(xyz, 0.05)
        coords.append(c)
        feats.append(f[um])
    return oq.batched_coordinates(coords), np.concatenate(feats)


def _training_step(model, bc, feats, labels):
    """What a Lightning module does: everything but the loss dies when this returns."""
    import MinkowskiEngine as ME
    st = ME.SparseTensor(features=torch.from_numpy(feats).cuda(), coordinates=torch.from_numpy(bc).cuda())
    out = model(st)
    logits = out.F if hasattr(out, "F") else out
    return torch.nn.functional.cross_entropy(logits.float(), labels)


@pytest.mark.parametrize("arch,mode,fused", [("MinkUNet34C", "bf16", True), ("MinkUNet34C", "bf16", False), ("MinkUNet34C", "fp32", True),
                                              ("MinkUNet50", "bf16", True), ("MinkUNet14A", "fp32", False)])
def test_backward_after_the_tensors_are_gone(cuda, arch, mode, fused):
    import gcdlss_b200
    from gcdlss_b200 import functional
    from models import minkunet as mu
    prev_mode, prev_fused = gcdlss_b200.get_math_mode(), functional._FUSED_C
    gcdlss_b200.set_math_mode(mode)
    functional._FUSED_C = fused
    try:
        torch.manual_seed(3)
        bc, feats = _batch()
        labels = torch.from_numpy(np.random.default_rng(0).integers(0, 17, bc.shape[0])).cuda()
        model = getattr(mu, arch)(1, 17).cuda().train()
        # reference gradients: same step with the tensors held until after backward
        import MinkowskiEngine as ME
        st = ME.SparseTensor(features=torch.from_numpy(feats).cuda(), coordinates=torch.from_numpy(bc).cuda())
        loss_held = torch.nn.functional.cross_entropy(model(st).F.float(), labels)
        loss_held.backward()
        held = torch.cat([p.grad.flatten().float() for p in model.parameters()]).clone()
        del st
        model.zero_grad(set_to_none=True)
        for m in model.modules():                  # same running statistics are irrelevant in train mode; keep counters tidy
            if hasattr(m, "reset_running_stats"):
                m.reset_running_stats()
        loss = _training_step(model, bc, feats, labels)
        gc.collect()
        torch.cuda.empty_cache()                   # a dangling pointer would now point at unmapped / reused memory
        junk = torch.full((64 << 20,), -1, dtype=torch.int32, device="cuda")      # ...and whatever is reused holds -1s
        loss.backward()
        torch.cuda.synchronize()
        del junk
        got = torch.cat([p.grad.flatten().float() for p in model.parameters()])
        assert torch.isfinite(got).all()
        # wgrad / BN reductions use atomics: equal up to summation order
        cos = float(torch.nn.functional.cosine_similarity(got, held, dim=0))
        print(arch, mode, "fused" if fused else "per-op", "loss", float(loss), float(loss_held), "cosine", cos)
        assert abs(float(loss) - float(loss_held)) < 1e-5 * max(1.0, abs(float(loss_held))) and cos > 0.9999
    finally:
        functional._FUSED_C = prev_fused
        gcdlss_b200.set_math_mode(prev_mode)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_lone_strided_convolution(cuda, mode):
    """An encoder-only use: one stride-2 convolution whose transposed map nobody ever asks for."""
    import gcdlss_b200
    import MinkowskiEngine as ME
    from oracle import conv as oc
    from oracle import coords as ocd
    prev = gcdlss_b200.get_math_mode()
    gcdlss_b200.set_math_mode(mode)
    try:
        torch.manual_seed(0)
        bc, _ = _batch(4000)
        conv = ME.MinkowskiConvolution(32, 64, kernel_size=2, stride=2, dimension=3).cuda()
        x = torch.randn(bc.shape[0], 32, device="cuda", requires_grad=True)

        def step():
            st = ME.SparseTensor(features=x, coordinates=torch.from_numpy(bc).cuda())
            return conv(st).F.square().sum()

        loss = step()
        gc.collect()
        torch.cuda.empty_cache()
        loss.backward()
        coarse, parent, code = ocd.stride2(bc, 1)
        xr = x.detach().cpu().double().requires_grad_(True)
        w = conv.kernel.detach().cpu().double().requires_grad_(True)
        oc.conv_table(xr, ocd.kmap_down2(parent, code, coarse.shape[0]), w).square().sum().backward()
        tol = TOL_FP32 if mode == "fp32" else TOL_BF16
        assert rel_err(x.grad, xr.grad) < tol and rel_err(conv.kernel.grad, w.grad) < tol
    finally:
        gcdlss_b200.set_math_mode(prev)
