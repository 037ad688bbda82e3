"""CUDA quantiser + dedup vs the oracle and the reference-pinned fixtures: bit exact."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from oracle import quantize as oq

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["f32", "f64", "b4"])
def test_sparse_quantize_matches_frozen(cuda, oracle_frozen, tag):
    import MinkowskiEngine as ME
    pts = oracle_frozen[f"me_{tag}_in"]
    c, um, inv = ME.utils.sparse_quantize(coordinates=pts, return_index=True, return_inverse=True, quantization_size=0.05)
    assert isinstance(c, np.ndarray) and c.dtype == np.int32
    assert um.dtype == torch.int64 and inv.dtype == torch.int64 and not um.is_cuda and not inv.is_cuda
    np.testing.assert_array_equal(c, oracle_frozen[f"me_{tag}_coords"])
    np.testing.assert_array_equal(um.numpy(), oracle_frozen[f"me_{tag}_umap"])
    np.testing.assert_array_equal(inv.numpy(), oracle_frozen[f"me_{tag}_inv"])
    # torch input (the inline Stage-2 call, ref exp_merge_mean_teacher.py:2856): coords stay torch
    ct, um2, inv2 = ME.utils.sparse_quantize(coordinates=torch.from_numpy(pts), return_index=True, return_inverse=True,
                                             quantization_size=0.05)
    assert isinstance(ct, torch.Tensor) and ct.dtype == torch.int32
    np.testing.assert_array_equal(ct.numpy(), c)
    feats = torch.arange(pts.shape[0]).float()[:, None]
    assert torch.equal(feats[um2], feats[torch.from_numpy(oracle_frozen[f"me_{tag}_umap"])])   # fancy-indexing contract


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_voxelize_minkunet_matches_reference(cuda, ref_pinned, tag):
    from gcdlss_b200.quantize import voxelize_minkunet
    pts = torch.from_numpy(ref_pinned[f"voxA_{tag}_points"]).cuda()
    d = voxelize_minkunet([pts, pts], [0.05, 0.05, 0.05])
    m = ref_pinned[f"voxA_{tag}_inds"].shape[0]
    for b in range(2):
        np.testing.assert_array_equal(d["voxel_inds"][b].cpu().numpy(), ref_pinned[f"voxA_{tag}_inds"])
        np.testing.assert_array_equal(d["point2voxel_maps"][b].cpu().numpy(), ref_pinned[f"voxA_{tag}_inverse"])
    coors = d["coors"].cpu().numpy()
    np.testing.assert_array_equal(coors[:m, 1:], ref_pinned[f"voxA_{tag}_coors"][ref_pinned[f"voxA_{tag}_inds"]])
    assert (coors[:m, 0] == 0).all() and (coors[m:, 0] == 1).all()
    np.testing.assert_array_equal(d["voxels"][:m].cpu().numpy(), ref_pinned[f"voxA_{tag}_points"][ref_pinned[f"voxA_{tag}_inds"]])


@pytest.mark.parametrize("kind,q", [("kitti", 0.05), ("nuscenes", 0.1)])
def test_full_size_scan_vs_oracle(cuda, kind, q):
    import MinkowskiEngine as ME
    from gcdlss_b200 import synth
    xyz, _ = synth.make_scan(kind, 3)
    for arr in (xyz, xyz.astype(np.float64) * 1.0000001):
        c, um, inv = ME.utils.sparse_quantize(coordinates=arr, return_index=True, return_inverse=True, quantization_size=q)
        c0, um0, inv0 = oq.sparse_quantize_me(arr, q)
        np.testing.assert_array_equal(c, c0)
        np.testing.assert_array_equal(um.numpy(), um0)
        np.testing.assert_array_equal(inv.numpy(), inv0)


def test_dense_1m_points_properties(cuda):
    """BASELINE config 5 size: size-independent properties instead of the (slow) oracle."""
    import MinkowskiEngine as ME
    from gcdlss_b200 import synth
    xyz, _ = synth.make_dense_scan(0, sweeps=10)
    assert xyz.shape[0] > 1_000_000
    c, um, inv = ME.utils.sparse_quantize(coordinates=xyz, return_index=True, return_inverse=True, quantization_size=0.05)
    d = oq.floor_div(xyz, 0.05)
    np.testing.assert_array_equal(c[inv.numpy()], d)                    # decode(encode(x)) == quantised x
    assert np.all(np.diff(um.numpy()) > 0)                               # first-occurrence order
    assert np.unique(c, axis=0).shape[0] == c.shape[0]                   # no duplicate voxels
    np.testing.assert_array_equal(d[um.numpy()], c)
    c2, um2, _ = ME.utils.sparse_quantize(coordinates=c.astype(np.float32), return_index=True, return_inverse=True, quantization_size=1.0)
    np.testing.assert_array_equal(c2, c)                                 # idempotent
    np.testing.assert_array_equal(um2.numpy(), np.arange(c.shape[0]))


def test_edge_cases(cuda):
    import MinkowskiEngine as ME
    c, um, inv = ME.utils.sparse_quantize(coordinates=np.zeros((0, 3), np.float32), return_index=True, return_inverse=True, quantization_size=0.05)
    assert c.shape == (0, 3) and um.numel() == 0 and inv.numel() == 0
    c, um, inv = ME.utils.sparse_quantize(coordinates=np.array([[-0.01, 0.0, 0.049]] * 5, np.float32), return_index=True, return_inverse=True,
                                          quantization_size=0.05)
    np.testing.assert_array_equal(c, [[-1, 0, 0]])
    np.testing.assert_array_equal(inv.numpy(), [0] * 5)
    with pytest.raises(RuntimeError, match="64-bit key"):
        ME.utils.sparse_quantize(coordinates=np.array([[1e7, 0, 0]], np.float32), return_index=True, return_inverse=True, quantization_size=0.05)
    bc = ME.utils.batched_coordinates([np.zeros((2, 3)), np.ones((1, 3))])
    assert bc.dtype == torch.int32 and bc.tolist() == [[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 1, 1]]
