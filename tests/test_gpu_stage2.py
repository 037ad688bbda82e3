"""Stage-2 mean-teacher step harness on the GPU (BASELINE config 4, single rank)."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401

pytestmark = pytest.mark.gpu


def _half(kind, idx0, n_scans, n_points, dev, with_labels):
    from gcdlss_b200.steps import make_stage2_half
    return make_stage2_half(kind, idx0, n_scans, n_points, dev, with_labels)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_stage2_step_runs_and_updates(cuda, mode):
    import copy
    import gcdlss_b200
    import MinkowskiEngine as ME
    from gcdlss_b200.steps import Stage2Harness
    from models.multiheadminkunet import MinkUNetRC
    gcdlss_b200.set_math_mode(mode)
    try:
        torch.manual_seed(0)
        student = MinkUNetRC(17).cuda().train()
        for name, n_out in (("final2", 3), ("final3", 2)):       # heads bolted on by the caller (ref exp_merge_mean_teacher.py:128-153)
            setattr(student.encoder, name, ME.MinkowskiConvolution(96, n_out, kernel_size=1, bias=True, dimension=3).cuda())
        teacher = copy.deepcopy(student)
        opt = torch.optim.SGD(student.parameters(), lr=0.01, momentum=0.9)
        h = Stage2Harness(student, teacher, opt, voxel_size=0.05)
        sup = _half("kitti", 0, 2, 5000, cuda, True)
        unsup = _half("kitti", 10, 2, 5000, cuda, False)
        t0 = [p.detach().clone() for p in teacher.parameters()]
        s0 = [p.detach().clone() for p in student.parameters()]
        losses = [float(h.step(sup, unsup)) for _ in range(3)]
        assert all(np.isfinite(l) for l in losses), losses
        moved_s = sum(float((a - b.detach()).abs().sum()) for a, b in zip(s0, student.parameters()))
        moved_t = sum(float((a - b.detach()).abs().sum()) for a, b in zip(t0, teacher.parameters()))
        assert moved_s > 0 and moved_t > 0 and moved_t < moved_s          # EMA follows the student slowly
        assert all(not p.requires_grad for p in teacher.parameters())
        print(mode, "stage-2 losses", losses)
    finally:
        gcdlss_b200.set_math_mode("fp32")


def test_laser_mix_partitions_points(cuda):
    from gcdlss_b200.steps import laser_mix
    g = torch.Generator(device="cpu").manual_seed(0)
    ps, pu = torch.randn(1000, 3, generator=g).cuda() * 10, torch.randn(800, 3, generator=g).cuda() * 10
    ps[:, 2], pu[:, 2] = ps[:, 2] * 0.15 - 1.0, pu[:, 2] * 0.15 - 1.0          # pitch angles spread over the [-25, 3] degree bands
    fs, fu = torch.ones(1000, 1).cuda(), torch.zeros(800, 1).cuda()
    ls, lu = torch.arange(1000).cuda(), -torch.arange(1, 801).cuda()
    (m1p, m1f, m1l), (m2p, m2f, m2l) = laser_mix(ps, pu, fs, fu, ls, lu, 4)
    assert m1p.shape[0] + m2p.shape[0] == 1800                           # every point lands in exactly one mixed scan
    assert sorted(torch.cat([m1l, m2l]).tolist()) == sorted(torch.cat([ls, lu]).tolist())
    assert 0 < m1f.sum() < 1000                                          # both sources contribute to each mix


def test_stage2_loss_against_the_oracle(cuda):
    """The whole Stage-2 data path of one step against the CPU oracle in fp64: shared SparseTensor -> teacher + student
    forward, CE + 200 * MSE(softmax), teacher pseudo labels devoxelised to points, LaserMix (the reference's row order,
    tests/test_lasermix_pinned.py), inline quantisation of the mixed batch (batch column divided too), second student
    forward, 0.1 * CE(ignore -1)  (ref modules/exp_merge_mean_teacher.py:2772-2875)."""
    import copy
    import gcdlss_b200
    import MinkowskiEngine as ME
    import torch.nn.functional as F
    from gcdlss_b200.steps import Stage2Harness, mix_transform
    from gpu_util import TOL_FP32
    from models.multiheadminkunet import MinkUNetRC
    from oracle import quantize as oq
    from oracle.minkunet import OracleMinkUNet
    gcdlss_b200.set_math_mode("fp32")
    torch.manual_seed(0)
    student = MinkUNetRC(17).cuda().train()
    student.encoder.final2 = ME.MinkowskiConvolution(96, 3, kernel_size=1, bias=True, dimension=3).cuda()
    student.encoder.final3 = ME.MinkowskiConvolution(96, 2, kernel_size=1, bias=True, dimension=3).cuda()
    teacher = copy.deepcopy(student)
    with torch.no_grad():                         # a teacher that differs from the student, so the consistency term is not zero
        for p in teacher.parameters():
            p.mul_(1.0 + 0.05 * torch.randn_like(p))
    sup, unsup = _half("kitti", 0, 2, 3000, cuda, True), _half("kitti", 10, 2, 3000, cuda, False)

    def oracle_params(model):
        return {k[len("encoder."):]: (v.detach().cpu().double() if v.is_floating_point() else v.cpu()) for k, v in model.state_dict().items()}

    om_s, om_t = OracleMinkUNet(oracle_params(student), "MinkUNet34RC", True), OracleMinkUNet(oracle_params(teacher), "MinkUNet34RC", True)
    h = Stage2Harness(student, teacher, torch.optim.SGD(student.parameters(), lr=0.0), voxel_size=0.05)
    loss = float(h.step(sup, unsup))

    cpu = lambda t: t.detach().cpu()
    uc = cpu(unsup["coords"]).clone()
    uc[:, 0] += 2
    bc = torch.cat([cpu(sup["coords"]), uc]).numpy()
    feats = torch.cat([cpu(sup["feats"]), cpu(unsup["feats"])]).double()
    n_sup = sup["coords"].shape[0]
    f_s, _, lv = om_s.features(bc, feats)
    f_t, _, _ = om_t.features(bc, feats, lv)
    lo_s, lo_t = om_s.forward_dummy(f_s), om_t.forward_dummy(f_t)
    ref = F.cross_entropy(lo_s[:n_sup], cpu(sup["labels"]).long())
    prob_s, prob_t = F.softmax(lo_s[n_sup:], 1), F.softmax(lo_t[n_sup:], 1)
    ref = ref + F.mse_loss(prob_s, prob_t) * 200.0
    mp, tl = torch.max(prob_t, 1)
    inv = torch.cat([cpu(i) for i in unsup["inverse_maps"]])
    pl = tl[inv].clone()
    pl[mp[inv] < 0.9] = -1
    pts = lambda d: {k: cpu(v) for k, v in d["points"].items()}
    mb, mf, ml = mix_transform(pts(sup), pts(unsup), pl, [3, 4])          # the harness's draws of step 0
    qc, um, _ = oq.sparse_quantize_me(mb.numpy(), 0.05)
    lo_m = om_s.forward_dummy(om_s.features(qc, mf[um].double())[0])
    labels_m = ml[um].long()
    if bool((labels_m >= 0).any()):
        ref = ref + 0.1 * F.cross_entropy(lo_m, labels_m, ignore_index=-1)
    print("stage-2 loss", loss, "oracle", float(ref), "confident pseudo labels:", int((pl >= 0).sum()))
    assert abs(loss - float(ref)) < TOL_FP32 * abs(float(ref))
