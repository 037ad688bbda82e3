"""mmdet3d-style front/back ends: hard / dynamic voxelisation, SegVFE, decode heads."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _points(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(n, 4, generator=g)
    p[:, 0] = p[:, 0] * 60 - 5          # some points outside [0, 50]
    p[:, 1] = p[:, 1] * 50 - 25
    p[:, 2] = p[:, 2] * 7 - 4.5
    return p


def test_hard_voxelize_matches_python_reference(cuda):
    from models.voxelizer import VoxelLayer
    layer = VoxelLayer(voxel_size=[0.5, 0.5, 0.5], point_cloud_range=[0, -20, -4, 50, 20, 2], max_num_points=5, max_voxels=(3000, 4000)).cuda().eval()
    pts = _points(20000)
    voxels, coors, num = layer(pts.cuda())
    # plain-python statement of mmcv's deterministic hard_voxelize
    lo, vs, grid = np.array([0, -20, -4.0], np.float32), np.float32(0.5), np.array([100, 80, 12])
    c = np.floor((pts[:, :3].numpy() - lo) / vs).astype(np.int32)
    ids, ref_vox, ref_cnt, ref_coor = {}, [], [], []
    for i in range(pts.shape[0]):
        if (c[i] < 0).any() or (c[i] >= grid).any():
            continue
        key = tuple(c[i])
        if key not in ids:
            if len(ids) >= 4000:
                continue
            ids[key] = len(ids); ref_vox.append(np.zeros((5, 4), np.float32)); ref_cnt.append(0); ref_coor.append(c[i][[2, 1, 0]])
        v = ids[key]
        if ref_cnt[v] < 5:
            ref_vox[v][ref_cnt[v]] = pts[i].numpy(); ref_cnt[v] += 1
    np.testing.assert_array_equal(coors.cpu().numpy(), np.stack(ref_coor))
    np.testing.assert_array_equal(num.cpu().numpy(), np.array(ref_cnt))
    np.testing.assert_array_equal(voxels.cpu().numpy(), np.stack(ref_vox))


def test_dynamic_and_cylindrical_voxelizer(cuda):
    from models.voxelizer import Voxelizer
    pts = [_points(3000, 1).cuda(), _points(2000, 2).cuda()]
    v = Voxelizer(voxel=True, voxel_type='dynamic', voxel_layer=dict(voxel_size=[0.5, 0.5, 0.5], point_cloud_range=[0, -20, -4, 50, 20, 2],
                                                                      max_num_points=-1, max_voxels=-1))
    d = v.voxelize(pts)
    assert d['coors'].shape == (5000, 4) and d['voxels'].shape == (5000, 4)
    c = d['coors'].cpu().numpy()
    ref = np.floor((pts[0][:, :3].cpu().numpy() - np.array([0, -20, -4], np.float32)) / np.float32(0.5)).astype(np.int32)
    ok = ((ref >= 0) & (ref < np.array([100, 80, 12]))).all(1)
    np.testing.assert_array_equal(c[:3000][ok][:, 1:], ref[ok][:, [2, 1, 0]])
    assert (c[:3000][~ok][:, 1:] == -1).all()
    vc = Voxelizer(voxel=True, voxel_type='cylindrical', voxel_layer=dict(voxel_size=None, grid_shape=[480, 360, 32],
                                                                           point_cloud_range=[0, -3.14159265359, -4, 50, 3.14159265359, 2],
                                                                           max_num_points=-1, max_voxels=-1))
    dc = vc.voxelize(pts)
    assert dc['voxels'].shape[1] == 3 + 2 + 1 and int(dc['coors'][:, 1].max()) <= 479 and int(dc['coors'][:, 1:].min()) >= 0


@pytest.mark.parametrize("mode", ["max", "avg"])
def test_segvfe_matches_torch_reference(cuda, mode):
    from models.encoder import SegVFE
    torch.manual_seed(0)
    vfe = SegVFE(in_channels=6, feat_channels=[32, 64], with_voxel_center=True, grid_shape=[480, 360, 32], voxel_size=None, mode=mode,
                 feat_compression=16).cuda().train()
    n = 6000
    feats = torch.randn(n, 6).cuda().requires_grad_(True)
    coors = torch.stack([torch.randint(0, 2, (n,)), torch.randint(0, 40, (n,)), torch.randint(0, 30, (n,)), torch.randint(0, 8, (n,))], 1).cuda().int()
    coors[::50, 2] = -1                                                  # dropped points
    out, vcoors = vfe(feats, coors)
    out.square().sum().backward()
    # torch reference of the scatter on the same point features
    x = feats.detach().clone().requires_grad_(True)
    centre = x.new_zeros(n, 3)
    centre[:, 0] = x[:, 0] - (coors[:, 1].float() * vfe.vx + vfe.x_offset)
    centre[:, 1] = x[:, 1] - (coors[:, 2].float() * vfe.vy + vfe.y_offset)
    centre[:, 2] = x[:, 2] - (coors[:, 3].float() * vfe.vz + vfe.z_offset)
    h = vfe.pre_norm(torch.cat([centre, x], -1))
    for layer in vfe.vfe_layers:
        h = layer(h)
    valid = (coors >= 0).all(1)
    uc, inv = torch.unique(coors[valid].long(), dim=0, return_inverse=True)
    m = uc.shape[0]
    if mode == "max":
        red = h.new_full((m, h.shape[1]), float("-inf")).scatter_reduce(0, inv[:, None].expand(-1, h.shape[1]), h[valid], "amax", include_self=True)
    else:
        red = h.new_zeros((m, h.shape[1])).index_add(0, inv, h[valid]) / torch.bincount(inv, minlength=m)[:, None]
    ref = vfe.compression_layers(red)
    ref.square().sum().backward()
    assert torch.equal(vcoors.long(), uc)                                # ascending coordinate order, like torch.unique
    assert float((out - ref).abs().max()) < 1e-4 * float(ref.abs().max())
    assert float((feats.grad - x.grad).abs().max()) < 2e-3 * float(x.grad.abs().max())


def test_decode_heads(cuda):
    import gcdlss_b200
    from conftest import small_cloud
    from models.decoder import Cylinder3DHead, MinkUNetHead
    from oracle import conv as oc, coords as ocd
    gcdlss_b200.set_math_mode("fp32")
    torch.manual_seed(0)
    head = MinkUNetHead(channels=96, num_classes=17, dropout_ratio=0.0, batch_first=True).cuda().eval()
    m, n = 500, 2000
    vd = {'voxel_feats': torch.randn(m, 96).cuda(), 'coors': torch.cat([torch.zeros(m, 1), torch.randint(0, 50, (m, 3))], 1).int().cuda(),
          'point2voxel_maps': [torch.randint(0, m, (n,)).cuda()]}
    pts_logits = head.predict(vd)
    ref = head.conv_seg(vd['voxel_feats'])[vd['point2voxel_maps'][0]]
    assert pts_logits[0].shape == (n, 17) and torch.allclose(pts_logits[0], ref, atol=1e-6)
    bc = small_cloud(3, 2000, spread=0.5, batch=0)
    ch = Cylinder3DHead(channels=16, num_classes=20).cuda().eval()

    class Sp:                                                            # spconv-style container
        features = torch.randn(bc.shape[0], 16).cuda()
        indices = torch.from_numpy(bc).cuda()
    logits = ch(Sp).F
    w = ch.conv_seg.kernel.detach().cpu().double()
    ref = oc.conv_table(Sp.features.cpu().double(), ocd.kmap_subm(bc, 3, 1), w, ch.conv_seg.bias.detach().cpu().double())
    assert float((logits.cpu().double() - ref).abs().max()) < 1e-4 * float(ref.abs().max())


def test_minkunet_backbone_every_kernel_call(cuda):
    """MinkUNetBackbone ('minkowski' back-end of ref models/backbone.py:47-254): every convolution / batch-norm kernel call of
    a forward + backward re-computed in fp64 from its actual operands (gpu_util.OpChecker), mmdet3d-style state-dict keys."""
    import gcdlss_b200
    from gpu_util import TOL_FP32, OpChecker
    from gcdlss_b200 import synth
    from models.backbone import MinkUNetBackbone
    from oracle import quantize as oq
    gcdlss_b200.set_math_mode("fp32")
    torch.manual_seed(0)
    xyz, f = synth.make_scan("kitti", 0, n_points=5000)
    c, um, _ = oq.sparse_quantize_me(xyz, 0.05)
    coors = torch.from_numpy(oq.batched_coordinates([c])).cuda()
    feats = torch.from_numpy(np.concatenate([xyz[um], f[um]], 1)).cuda()
    net = MinkUNetBackbone(in_channels=4, encoder_blocks=[1, 1, 1, 1], decoder_blocks=[1, 1, 1, 1]).cuda().train()
    keys = list(net.state_dict())
    assert "conv_input.0.net.0.kernel" in keys and "encoder.1.1.downsample.net.0.kernel" in keys and "decoder.0.0.net.0.kernel" in keys
    with OpChecker() as chk:
        out = net(feats, coors)
        out.square().mean().backward()
    assert out.shape == (coors.shape[0], 96) and torch.isfinite(out).all()
    print("backbone ops checked:", len(chk.records), "worst:", chk.worst())
    assert len(chk.records) > 60 and chk.worst()[-1] < TOL_FP32
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    # and the fused path (blocks sequenced in C) gives the same features
    out2 = net(feats, coors)
    assert rel_err(out2, out) < 1e-5


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_heads_as_one_product(cuda, mode):
    """final / final2 / final3 as one 96 -> (C + 3 + n) product (MinkUNetBaseRC._heads, ref models/minkunet.py:312-362)
    against the three separate convolutions: logits and every gradient."""
    import gcdlss_b200
    import MinkowskiEngine as ME
    from models import minkunet as mu
    prev = gcdlss_b200.get_math_mode()
    gcdlss_b200.set_math_mode(mode)
    try:
        torch.manual_seed(0)
        net = mu.MinkUNet34RC(1, 17).cuda()
        net.final2 = ME.MinkowskiConvolution(96, 3, kernel_size=1, bias=True, dimension=3).cuda()
        net.final3 = ME.MinkowskiConvolution(96, 2, kernel_size=1, bias=True, dimension=3).cuda()
        c = torch.unique(torch.randint(0, 40, (3000, 3), dtype=torch.int32), dim=0)
        bc = torch.cat([torch.zeros(c.shape[0], 1, dtype=torch.int32), c], 1).cuda()
        x = torch.randn(bc.shape[0], 96, device="cuda")
        if mode == "bf16":
            x = x.to(torch.bfloat16).float()
        res = []
        for fused in (True, False):
            xin = x.clone().requires_grad_(True)
            st = ME.SparseTensor(features=xin.to(torch.bfloat16) if mode == "bf16" else xin, coordinates=bc)
            net.zero_grad(set_to_none=True)
            if fused:
                logits = net.forward_novel(st)
            else:
                rc = torch.max(net.final2(st).F, dim=1, keepdim=True)[0]
                logits = torch.cat([net.final(st).F, net.final3(st).F, rc], dim=1)
            (logits * torch.linspace(-1, 1, logits.shape[1], device="cuda")).sum().backward()
            res.append((logits.detach(), xin.grad.clone(), [getattr(net, n).kernel.grad.clone() for n in ("final", "final2", "final3")],
                        [getattr(net, n).bias.grad.clone() for n in ("final", "final2", "final3")]))
        tol = 1e-5 if mode == "fp32" else 2e-2
        assert res[0][0].shape == (bc.shape[0], 17 + 2 + 1)
        assert rel_err(res[0][0], res[1][0]) < tol and rel_err(res[0][1], res[1][1]) < tol
        for a, b in zip(res[0][2] + res[0][3], res[1][2] + res[1][3]):
            assert rel_err(a, b) < tol
    finally:
        gcdlss_b200.set_math_mode(prev)
