"""Per-op parity at the bench's size: one kitti_b4 batch (4 full synthetic sweeps, ~270 k voxels) through MinkUNet34C,
every kernel call re-computed in fp64 on the GPU from the operands the kernel received (gpu_util.OpChecker).  The small
network tests run 2 x 6 000-point scans; this one exercises what they cannot: several row tiles per CTA, the N-split of
the deep levels, tile-sorted tables of 270 k columns, pair lists of millions of entries."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from gpu_util import TOL_BF16, TOL_FP32, OpChecker
from oracle import quantize as oq

pytestmark = pytest.mark.gpu


def _kitti_b4():
    from gcdlss_b200 import synth
    from gcdlss_b200.quantize import sparse_quantize_gpu
    coords, feats = [], []
    for b in range(4):
        xyz, f = synth.make_scan("kitti", b)
        c, um, _ = sparse_quantize_gpu(torch.from_numpy(xyz).cuda(), 0.05)
        coords.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32, device="cuda"), c], 1))
        feats.append(torch.from_numpy(f).cuda().index_select(0, um))
    return torch.cat(coords), torch.cat(feats)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_every_kernel_call_at_bench_size(cuda, mode):
    import gcdlss_b200
    import MinkowskiEngine as ME
    from models import minkunet as mu
    prev = gcdlss_b200.get_math_mode()
    gcdlss_b200.set_math_mode(mode)
    try:
        torch.manual_seed(1234)
        bc, feats = _kitti_b4()
        assert bc.shape[0] > 200_000
        model = mu.MinkUNet34C(1, 17).cuda().train()
        labels = torch.randint(0, 17, (bc.shape[0],), device="cuda")
        with OpChecker() as chk:
            out = model(ME.SparseTensor(features=feats, coordinates=bc)).F
            torch.nn.functional.cross_entropy(out.float(), labels).backward()
        worst = chk.worst()
        print(mode, "voxels", bc.shape[0], "ops checked:", len(chk.records), "worst:", worst)
        assert len(chk.records) > 200 and worst[-1] < (TOL_BF16 if mode == "bf16" else TOL_FP32), worst
        assert all(torch.isfinite(p.grad).all() for p in model.parameters())
    finally:
        gcdlss_b200.set_math_mode(prev)
