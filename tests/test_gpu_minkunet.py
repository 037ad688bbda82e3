"""Whole-network parity: MinkUNet34C / 34RC forward + backward on the GPU vs the oracle."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from gpu_util import TOL_FP32, rel_err
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


def _oracle_params(model, dtype=torch.float64):
    return {k: (v.detach().cpu().to(dtype) if v.is_floating_point() else v.detach().cpu().clone()) for k, v in model.state_dict().items()}


def _l2_rel(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.mark.parametrize("arch", ["MinkUNet34C", "MinkUNet14A"])
def test_forward_backward_fp32(cuda, arch):
    """fp32 path.  Forward: logits vs the fp64 oracle within 1e-4.  Backward: (1) every kernel call of the
    step re-computed in fp64 on its actual inputs within 1e-4 (see gpu_util.OpChecker for why the per-op
    form is the sharp one); (2) end-to-end parameter gradients vs the oracle's autograd: relative L2 error
    per tensor (ReLU-mask flips at |pre-activation| < 1 ulp make a max-norm bound meaningless; the fp32 and
    fp64 CPU oracles differ from each other by the same amounts)."""
    import gcdlss_b200
    import MinkowskiEngine as ME
    from gpu_util import OpChecker
    from models import minkunet as mu
    gcdlss_b200.set_math_mode("fp32")
    torch.manual_seed(1234)
    bc, feats = _batch()
    model = getattr(mu, arch)(1, 17).cuda().train()
    params = _oracle_params(model)
    for k in params:
        if params[k].is_floating_point() and "running" not in k:
            params[k].requires_grad_(True)
    labels = torch.from_numpy(np.random.default_rng(0).integers(0, 17, bc.shape[0]))

    with OpChecker() as chk:
        st = ME.SparseTensor(features=torch.from_numpy(feats).cuda(), coordinates=torch.from_numpy(bc).cuda())
        logits = model(st).F
        loss = torch.nn.functional.cross_entropy(logits, labels.cuda())
        loss.backward()
    print(arch, "ops checked:", len(chk.records), "worst per-op rel err:", chk.worst())
    assert len(chk.records) > 90 and chk.worst()[-1] < TOL_FP32

    om = OracleMinkUNet(params, arch, training=True)
    lo, f96, _ = om.forward(bc, torch.from_numpy(feats).double())
    loss_o = torch.nn.functional.cross_entropy(lo, labels)
    loss_o.backward()

    e_logits = rel_err(logits.detach(), lo.detach())
    print(arch, "logits rel err", e_logits, "loss", float(loss), float(loss_o))
    assert e_logits < TOL_FP32
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        worst = max(worst, (name, _l2_rel(p.grad, params[name].grad)), key=lambda t: t[1])
    print(arch, "worst end-to-end grad relative L2 err", worst)
    assert worst[1] < 5e-2, worst
    sd = model.state_dict()
    for k in sd:
        if "running" in k:
            assert rel_err(sd[k], params[k]) < 1e-4, k
        if "num_batches_tracked" in k:
            assert int(sd[k]) == 1


def test_rc_heads_and_row_order(cuda):
    """Stage-2 model: forward_dummy / forward_novel logits layout, teacher+student sharing one SparseTensor."""
    import gcdlss_b200
    import MinkowskiEngine as ME
    from models.multiheadminkunet import MinkUNetRC
    gcdlss_b200.set_math_mode("fp32")
    torch.manual_seed(0)
    bc, feats = _batch(3000)
    student, teacher = MinkUNetRC(17).cuda(), MinkUNetRC(17).cuda()
    for m in (student, teacher):
        m.encoder.final2 = ME.MinkowskiConvolution(96, 3, kernel_size=1, bias=True, dimension=3).cuda()
        m.encoder.final3 = ME.MinkowskiConvolution(96, 2, kernel_size=1, bias=True, dimension=3).cuda()
    teacher.load_state_dict(student.state_dict())
    st = ME.SparseTensor(features=torch.from_numpy(feats).cuda(), coordinates=torch.from_numpy(bc).cuda())
    out_t, out_s = teacher(st), student(st)
    assert len(st.coordinate_manager._kmaps) >= 10                       # maps built once, shared by both models
    assert out_s["logits"].shape == (bc.shape[0], 18) and out_s["feats"].shape == (bc.shape[0], 96)
    assert rel_err(out_t["logits"], out_s["logits"]) < 1e-6
    disc = student.forward_discover(st)["logits"]
    assert disc.shape == (bc.shape[0], 17 + 2 + 1)
    params = {k[len("encoder."):]: v.detach().cpu().double() if v.is_floating_point() else v.cpu() for k, v in student.state_dict().items()}
    om = OracleMinkUNet(params, "MinkUNet34RC", training=True)
    f96, _, _ = om.features(bc, torch.from_numpy(feats).double())
    assert rel_err(out_s["feats"], f96) < TOL_FP32
    assert rel_err(out_s["logits"], om.forward_dummy(f96)) < TOL_FP32
    # eval mode (validation path, ref exp_merge_mean_teacher.py:2263-2330): running statistics
    student.eval()
    with torch.no_grad():
        ev = student.forward_discover(st)["logits"]
    params = {k[len("encoder."):]: v.detach().cpu().double() if v.is_floating_point() else v.cpu() for k, v in student.state_dict().items()}
    ome = OracleMinkUNet(params, "MinkUNet34RC", training=False)
    f96e, _, _ = ome.features(bc, torch.from_numpy(feats).double())
    assert rel_err(ev, ome.forward_novel(f96e)) < TOL_FP32


def test_bf16_tensor_core_path_end_to_end(cuda):
    """bf16 tcgen05 path on the whole network.  Stated bounds: every kernel call within 3e-2 (max-norm,
    relative) of its fp64 re-computation from the same bf16 inputs; logits within 5e-2 and loss within 1e-3
    relative of the fp32 path; concatenated parameter gradient cosine > 0.95 against the fp32 path (measured
    0.969: activations, BN inputs and activation gradients are all stored in bf16 across 63 conv layers)."""
    import gcdlss_b200
    import MinkowskiEngine as ME
    from gpu_util import TOL_BF16, OpChecker
    from models import minkunet as mu
    torch.manual_seed(1234)
    bc, feats = _batch()
    model = mu.MinkUNet34C(1, 17).cuda().train()
    labels = torch.from_numpy(np.random.default_rng(0).integers(0, 17, bc.shape[0])).cuda()
    res = {}
    for mode in ("fp32", "bf16"):
        gcdlss_b200.set_math_mode(mode)
        model.zero_grad(set_to_none=True)
        with OpChecker() as chk:
            st = ME.SparseTensor(features=torch.from_numpy(feats).cuda(), coordinates=torch.from_numpy(bc).cuda())
            logits = model(st).F
            loss = torch.nn.functional.cross_entropy(logits, labels)
            loss.backward()
        if mode == "bf16":
            print("bf16 ops checked:", len(chk.records), "worst per-op rel err:", chk.worst())
            assert chk.worst()[-1] < TOL_BF16, chk.worst()
        res[mode] = (logits.detach().clone(), float(loss), torch.cat([p.grad.flatten() for p in model.parameters()]))
    gcdlss_b200.set_math_mode("fp32")
    e = rel_err(res["bf16"][0], res["fp32"][0])
    cos = torch.nn.functional.cosine_similarity(res["bf16"][2], res["fp32"][2], dim=0).item()
    print("bf16 vs fp32: logits rel err", e, "loss", res["bf16"][1], res["fp32"][1], "global grad cosine", cos)
    assert e < 5e-2 and abs(res["bf16"][1] - res["fp32"][1]) < 1e-3 * abs(res["fp32"][1])
    assert cos > 0.95
