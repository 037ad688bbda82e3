"""Run-table kernel-map search (gcd_runtable_build + gcd_kmap_subm_runs, opt-in GCDLSS_KMAP=runs) on the GPU: bit exact
against the oracle and against the point-wise search.  The per-thread logic is the source tests/test_emulated_kernels.py
already holds to the oracle on the CPU; what runs for the first time here is the 256-bit slot load and real concurrency
of the inserts.  (File name sorts last on purpose: the entry points are new and opt-in.)"""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from conftest import small_cloud
from oracle import coords as ocd
from oracle import quantize as oq

pytestmark = pytest.mark.gpu


@pytest.fixture()
def runs_mode():
    import gcdlss_b200
    gcdlss_b200.set_kmap_search("runs")
    yield
    gcdlss_b200.set_kmap_search("points")


def both(bc, k, ts):
    from gcdlss_b200 import ops
    c = torch.from_numpy(np.ascontiguousarray(bc, np.int32)).cuda()
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    runs = ops.kmap_subm_runs(c, ops.runtable_build(c, ts, status), k, ts)
    points = ops.kmap_subm(c, ops.hash_build(c, status), k, ts)
    assert int(status.item()) == 0
    return runs.cpu().numpy(), points.cpu().numpy()


@pytest.mark.parametrize("k", [3, 5])
@pytest.mark.parametrize("ts", [1, 2, 8])
def test_runs_equal_points_and_oracle(cuda, k, ts):
    parts = [small_cloud(20 + i, 4000, spread=0.5, batch=20 * i) for i in range(3)]
    bc = np.concatenate(parts)
    bc[:, 1:] -= 11
    bc[:, 1:] *= ts
    runs, points = both(bc, k, ts)
    np.testing.assert_array_equal(runs, points)
    np.testing.assert_array_equal(runs.T, ocd.kmap_subm(bc, k, ts))


def test_kitti_batch_all_levels_through_the_manager(cuda, runs_mode):
    from gcdlss_b200 import synth
    from gcdlss_b200.coords import CoordinateManager
    scans = [oq.sparse_quantize_me(synth.make_scan("kitti", i, n_points=None)[0], 0.05)[0] for i in range(2)]
    bc = oq.batched_coordinates(scans)
    mgr = CoordinateManager(torch.from_numpy(bc).cuda())
    assert mgr.runs and mgr.maps[1].table is None
    lv = ocd.CoordLevels(bc)
    for l in range(5):
        np.testing.assert_array_equal(mgr.kernel_map(1 << l, 3, 1, False).nbr.cpu().numpy().T, lv.subm(l, 3))
    np.testing.assert_array_equal(mgr.kernel_map(1, 5, 1, False).nbr.cpu().numpy().T, lv.subm(0, 5))
    mgr.check()


def test_dense_scan_properties(cuda):
    # BASELINE config 5 size (ten merged sweeps, ~1 M points): the oracle is too slow here, so size-independent
    # properties: identical to the point-wise search, centre tap = identity, stride-1 maps are symmetric
    from gcdlss_b200 import synth
    from gcdlss_b200.quantize import sparse_quantize_gpu
    pts = np.concatenate([synth.make_scan("kitti", i)[0] + np.array([0.5 * i, 0, 0], np.float32) for i in range(10)])
    c, _, _ = sparse_quantize_gpu(torch.from_numpy(pts).cuda(), 0.05)
    bc = torch.cat([torch.zeros((c.shape[0], 1), dtype=torch.int32, device="cuda"), c], 1).cpu().numpy()
    runs, points = both(bc, 3, 1)
    np.testing.assert_array_equal(runs, points)
    n = bc.shape[0]
    np.testing.assert_array_equal(runs[13], np.arange(n))
    for k in (0, 5, 12):
        o = np.nonzero(runs[k] >= 0)[0]
        np.testing.assert_array_equal(runs[26 - k][runs[k][o]], o)


def test_status_bits(cuda, runs_mode):
    import MinkowskiEngine as ME
    c = torch.tensor([[0, 1, 2, 3], [0, 1, 2, 3]], dtype=torch.int32).cuda()
    st = ME.SparseTensor(features=torch.ones(2, 1).cuda(), coordinates=c)
    with pytest.raises(RuntimeError, match="duplicate"):
        st.C
    c = torch.tensor([[0, 1 << 20, 2, 3]], dtype=torch.int32).cuda()
    with pytest.raises(RuntimeError, match="64-bit key"):
        ME.SparseTensor(features=torch.ones(1, 1).cuda(), coordinates=c).C


def test_model_forward_is_identical_in_both_modes(cuda):
    import gcdlss_b200
    import MinkowskiEngine as ME
    from gcdlss_b200 import synth
    from models import minkunet as mu
    xyz, feat = synth.make_scan("kitti", 1, n_points=6000)
    c, um, _ = oq.sparse_quantize_me(xyz, 0.05)
    bc = torch.from_numpy(oq.batched_coordinates([c])).cuda()
    f = torch.from_numpy(feat[um]).cuda()
    torch.manual_seed(0)
    model = mu.MinkUNet14A(1, 17).cuda().eval()
    outs = []
    for kind in ("points", "runs"):
        gcdlss_b200.set_kmap_search(kind)
        try:
            with torch.no_grad():
                outs.append(model(ME.SparseTensor(features=f, coordinates=bc)).F.clone())
        finally:
            gcdlss_b200.set_kmap_search("points")
    assert torch.equal(outs[0], outs[1])


# ---- fused consistency terms (gcd_consistency_rows): same file because it is as new as the run table
@pytest.mark.parametrize("n,c", [(1, 2), (4099, 17), (200000, 20)])
def test_consistency_terms_against_torch(cuda, n, c):
    import torch.nn.functional as F
    from gcdlss_b200.steps import consistency_terms
    g = torch.Generator(device="cuda").manual_seed(n)
    ls = (torch.randn(n, c, device="cuda", generator=g) * 3).requires_grad_(True)
    lt = torch.randn(n, c, device="cuda", generator=g) * 3
    mse, prob, label = consistency_terms(ls, lt, threshold=0.9)
    (mse * 200.0).backward()
    got_grad = ls.grad.clone()
    ls.grad = None
    # reference in double precision: the gradient 2 ps ((ps - pt) - sum_j ps_j (ps_j - pt_j)) is a difference of
    # probabilities, so an fp32 evaluation (torch's or the kernel's) carries an absolute error of a few ulp(1) times the
    # loss scale 200 / (n c) -- at n = 1, c = 2 two fp32 evaluations differ by 1.4e-4 relative from each other
    ld = ls.detach().double().requires_grad_(True)
    ps, pt = F.softmax(ld, 1), F.softmax(lt.double(), 1)
    ref = F.mse_loss(ps, pt)
    (ref * 200.0).backward()
    ref_prob, ref_label = torch.max(pt, 1)
    assert abs(float(mse) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    torch.testing.assert_close(prob.double(), ref_prob, rtol=1e-5, atol=1e-7)
    eps = 2.0 ** -23
    torch.testing.assert_close(got_grad.double(), ld.grad, rtol=1e-4, atol=1e-9 + 16 * eps * 200.0 / (n * c))
    sure = (ref_prob - 0.9).abs() > 1e-6
    expect = torch.where(ref_prob < 0.9, torch.full_like(ref_label, -1), ref_label)
    assert torch.equal(label[sure], expect[sure])
