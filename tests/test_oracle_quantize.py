"""Oracle vs the reference's own outputs (tests/golden/reference_pinned.npz) + property tests."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import quantize as oq


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_flavour_a_matches_reference(ref_pinned, tag):
    pts = ref_pinned[f"voxA_{tag}_points"]
    c = oq.round_half_even_div_f32(pts[:, :3], [0.05, 0.05, 0.05])
    c = c - c.min(0)
    np.testing.assert_array_equal(c, ref_pinned[f"voxA_{tag}_coors"])
    np.testing.assert_array_equal(oq.ravel_hash(c), ref_pinned[f"voxA_{tag}_hash"])
    inds, inv = oq.sparse_quantize_np_unique(c)
    np.testing.assert_array_equal(inds, ref_pinned[f"voxA_{tag}_inds"])
    np.testing.assert_array_equal(inv, ref_pinned[f"voxA_{tag}_inverse"])


def test_voxelize_minkunet_dict(ref_pinned):
    pts = [ref_pinned["voxA_a_points"], ref_pinned["voxA_b_points"]]
    d = oq.voxelize_minkunet(pts, [0.05, 0.05, 0.05])
    m0 = ref_pinned["voxA_a_inds"].shape[0]
    np.testing.assert_array_equal(d["voxel_inds"][0], ref_pinned["voxA_a_inds"])
    np.testing.assert_array_equal(d["point2voxel_maps"][1], ref_pinned["voxA_b_inverse"])
    assert d["coors"].dtype == np.int32 and d["coors"].shape[1] == 4
    assert (d["coors"][:m0, 0] == 0).all() and (d["coors"][m0:, 0] == 1).all()
    np.testing.assert_array_equal(d["voxels"][:m0], pts[0][ref_pinned["voxA_a_inds"]])


def test_aug_matrices_match_reference(ref_pinned):
    import _paths  # noqa: F401
    from utils.voxelizer import Voxelizer
    np.random.seed(1234)
    v = Voxelizer(voxel_size=0.05, use_augmentation=True, scale_augmentation_bound=(0.95, 1.05),
                  rotation_augmentation_bound=((-np.pi / 20, np.pi / 20), (-np.pi / 20, np.pi / 20), (-np.pi, np.pi)),
                  translation_augmentation_ratio_bound=((-3, 3), (-3, 3), (-0.5, 0.5)))
    for i in range(3):
        s, r = v.get_transformation_matrix()
        np.testing.assert_array_equal(s, ref_pinned["aug_scale"][i])
        np.testing.assert_allclose(r, ref_pinned["aug_rigid"][i], rtol=0, atol=1e-15)
    s, r = Voxelizer().get_transformation_matrix()
    np.testing.assert_array_equal(s, np.eye(4))
    np.testing.assert_array_equal(r, np.eye(4))


@pytest.mark.parametrize("tag", ["f32", "f64", "b4"])
def test_flavour_b_frozen(oracle_frozen, tag):
    c, um, inv = oq.sparse_quantize_me(oracle_frozen[f"me_{tag}_in"], 0.05)
    np.testing.assert_array_equal(c, oracle_frozen[f"me_{tag}_coords"])
    np.testing.assert_array_equal(um, oracle_frozen[f"me_{tag}_umap"])
    np.testing.assert_array_equal(inv, oracle_frozen[f"me_{tag}_inv"])


def test_lasermix_batch_column_quirk():
    # ref modules/exp_merge_mean_teacher.py:2856: the batch column is divided by the voxel size too
    p = np.array([[0, 0.01, 0.01, 0.01], [1, 0.01, 0.01, 0.01], [2, 0.2, 0, 0], [3, 0, 0, -0.01]], np.float32)
    c, um, inv = oq.sparse_quantize_me(p, 0.05)
    np.testing.assert_array_equal(c[:, 0], [0, 20, 40, 60])
    assert c[3, 3] == -1


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(0, 300), st.sampled_from([0.05, 0.1, 0.37]))
def test_me_quantize_properties(seed, n, q):
    rng = np.random.default_rng(seed)
    p = (rng.normal(0, 0.5, (n, 3)) * rng.choice([1, 10])).astype(rng.choice([np.float32, np.float64]))
    c, um, inv = oq.sparse_quantize_me(p, q)
    d = oq.floor_div(p, q)
    assert c.dtype == np.int32 and um.dtype == np.int64 and inv.dtype == np.int64
    np.testing.assert_array_equal(c[inv], d)                       # coords[unique][inverse] == coords
    assert np.all(np.diff(um) > 0)                                 # first-occurrence => ascending
    assert np.unique(c, axis=0).shape[0] == c.shape[0]             # bijection onto distinct voxels
    for j in range(min(5, c.shape[0])):                            # unique_map really is the FIRST occurrence
        first = np.nonzero((d == c[j]).all(1))[0][0]
        assert um[j] == first
    c2, um2, inv2 = oq.sparse_quantize_me(c.astype(np.float64), 1.0)  # idempotent on voxel coords
    np.testing.assert_array_equal(c2, c)
    np.testing.assert_array_equal(um2, np.arange(c.shape[0]))


def test_empty_and_single():
    c, um, inv = oq.sparse_quantize_me(np.zeros((0, 3), np.float32), 0.05)
    assert c.shape == (0, 3) and um.shape == (0,) and inv.shape == (0,)
    c, um, inv = oq.sparse_quantize_me(np.array([[-0.01, 0.0, 0.049]], np.float32), 0.05)
    np.testing.assert_array_equal(c, [[-1, 0, 0]])
    bc = oq.batched_coordinates([np.zeros((2, 3)), np.ones((1, 3))])
    np.testing.assert_array_equal(bc, [[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 1, 1]])
