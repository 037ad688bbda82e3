"""GPU-side Dataset / collate work (gcdlss_b200/dataprep.py) against the reference's numpy statements:
float64 rigid transform of float32 points (ref utils/dataset_remission.py:821-833), sparse_quantize of the float64
coordinates, and the collate functions' keys / dtypes / batch column (ref utils/collation.py:29-42, 430-467)."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from oracle import quantize as oq

pytestmark = pytest.mark.gpu


def _matrices(seed):
    """affine_mtx @ voxel_mtx as utils/voxelizer.get_transformation_matrix composes them (scale, rotation, translation)."""
    rng = np.random.default_rng(seed)
    voxel = np.eye(4)
    np.fill_diagonal(voxel[:3, :3], rng.uniform(0.95, 1.05))
    ang = rng.uniform(-np.pi, np.pi)
    rot = np.eye(4)
    rot[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    rot[:3, 3] = rng.uniform(-3, 3, 3)
    return rot @ voxel


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_transform_and_quantise_match_numpy(cuda, seed):
    from gcdlss_b200 import ops, synth
    from gcdlss_b200.dataprep import prepare_scan
    xyz, f = synth.make_scan("kitti", seed)
    rigid = _matrices(seed)
    sel = np.sort(np.random.default_rng(seed).choice(xyz.shape[0], 80000, replace=False))
    ref = oq.dataset_transform(xyz[sel], rigid)
    got = ops.affine_f64(torch.from_numpy(xyz[sel]).cuda(), rigid).cpu().numpy()
    ulp = np.abs(got - ref) / np.spacing(np.abs(ref))
    print("transform: max |ulp| vs numpy", ulp.max(), "exactly equal:", float((got == ref).mean()))
    assert ulp.max() <= 1.0                                     # BLAS may order the three products differently: one ulp at most
    labels = np.random.default_rng(seed).integers(0, 17, xyz.shape[0])
    c, feats, lab, sidx, mlab, inv = prepare_scan(torch.from_numpy(xyz).cuda(), torch.from_numpy(f[:, 0]).cuda(), torch.from_numpy(labels).cuda(),
                                                  None, rigid, 0.05, selected_idx=sel)
    c0, um0, inv0 = oq.sparse_quantize_me(got, 0.05)            # float64 coordinates, floor in float64
    np.testing.assert_array_equal(c.cpu().numpy(), c0)
    np.testing.assert_array_equal(inv.cpu().numpy(), inv0)
    np.testing.assert_array_equal(sidx.cpu().numpy(), sel[um0])
    np.testing.assert_array_equal(lab.cpu().numpy(), labels[sel][um0])
    np.testing.assert_array_equal(feats.cpu().numpy(), f[sel][um0])
    # voxels whose coordinate sits within an ulp of a voxel boundary may differ between numpy's product and ours
    flips = (np.floor(ref / 0.05) != np.floor(got / 0.05)).any(1).sum()
    assert flips <= 2, flips


def test_collate_has_the_references_layout(cuda):
    from gcdlss_b200 import synth
    from gcdlss_b200.dataprep import collate, collate_lasermix, prepare_scan
    samples, points = [], []
    for b in range(3):
        xyz, f = synth.make_scan("kitti", b, n_points=3000)
        lab = torch.arange(xyz.shape[0], device="cuda") % 17
        p = torch.from_numpy(xyz).cuda()
        samples.append(prepare_scan(p, torch.from_numpy(f[:, 0]).cuda(), lab, lab, None, 0.05))
        points.append((p, torch.from_numpy(f).cuda(), lab, torch.arange(xyz.shape[0], device="cuda"), lab))
    bc, feats, labels, sel, mapped, invs, idx = collate(samples, pcd_indexes=[7, 8, 9])
    assert bc.dtype == torch.int32 and bc.shape[1] == 4 and feats.dtype == torch.float32 and labels.dtype == torch.int32
    assert sel.dtype == torch.int64 and idx.dtype == torch.int16 and idx.tolist() == [7, 8, 9] and len(invs) == 3
    ref = oq.batched_coordinates([s[0].cpu().numpy() for s in samples])
    np.testing.assert_array_equal(bc.cpu().numpy(), ref)
    d = collate_lasermix(points, samples)
    assert set(d) == {"points", "voxel"} and set(d["voxel"]) == {"coords", "feats", "labels", "selected_idx", "mapped_labels", "pcd_indexes", "inverse_maps"}
    assert d["points"]["coords"].dtype == torch.float32 and d["points"]["coords"].shape[1] == 4
    assert torch.equal(d["points"]["coords"][:, 0].unique().cpu(), torch.tensor([0.0, 1.0, 2.0]))
