"""CUDA coordinate maps / kernel maps vs the oracle: bit exact."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from conftest import small_cloud
from oracle import coords as ocd
from oracle import quantize as oq

pytestmark = pytest.mark.gpu


def build_manager(bc):
    from gcdlss_b200.coords import CoordinateManager
    return CoordinateManager(torch.from_numpy(bc).cuda())


def check_levels(mgr, lv):
    for l in range(1, 5):
        m = mgr.get_map(1 << l)
        np.testing.assert_array_equal(m.coords.cpu().numpy(), lv.coords[l])
        fine = mgr.get_map(1 << (l - 1))
        np.testing.assert_array_equal(fine.parent.cpu().numpy(), lv.parent[l - 1])
        np.testing.assert_array_equal(fine.code.cpu().numpy(), lv.code[l - 1])
    for l in range(5):
        km = mgr.kernel_map(1 << l, 3, 1, False)
        np.testing.assert_array_equal(km.nbr.cpu().numpy().T, lv.subm(l, 3))
    np.testing.assert_array_equal(mgr.kernel_map(1, 5, 1, False).nbr.cpu().numpy().T, lv.subm(0, 5))
    for l in range(4):
        np.testing.assert_array_equal(mgr.kernel_map(1 << l, 2, 2, False).nbr.cpu().numpy().T, lv.down(l))
        np.testing.assert_array_equal(mgr.kernel_map(2 << l, 2, 2, True).nbr.cpu().numpy().T, lv.up(l))


def test_maps_match_frozen_fixture(cuda, oracle_frozen):
    bc = oracle_frozen["map_coords0"]
    mgr = build_manager(bc)
    lv = ocd.CoordLevels(bc)
    for l in range(1, 5):                                   # oracle itself still equals the frozen file
        np.testing.assert_array_equal(lv.coords[l], oracle_frozen[f"map_coords{l}"])
    check_levels(mgr, lv)


def test_maps_kitti_batch(cuda):
    from gcdlss_b200 import synth
    scans = [oq.sparse_quantize_me(synth.make_scan("kitti", i, n_points=20000)[0], 0.05)[0] for i in range(2)]
    bc = oq.batched_coordinates(scans)
    check_levels(build_manager(bc), ocd.CoordLevels(bc))


def test_pair_lists(cuda):
    bc = small_cloud(4, 3000, spread=0.6, batch=0)
    mgr = build_manager(bc)
    lv = ocd.CoordLevels(bc)
    for km, table in ((mgr.kernel_map(1, 3, 1, False), lv.subm(0, 3)), (mgr.kernel_map(1, 2, 2, False), lv.down(0)),
                      (mgr.kernel_map(2, 2, 2, True), lv.up(0))):
        pi, po, off = ocd.pairs_from_table(table)
        gi, go, goff = km.pairs
        np.testing.assert_array_equal(goff.cpu().numpy(), off)
        np.testing.assert_array_equal(gi.cpu().numpy()[: off[-1]], pi)
        np.testing.assert_array_equal(go.cpu().numpy()[: off[-1]], po)
        assert km.num_pairs() == off[-1]


@pytest.mark.parametrize("form", [0, 1, 2])
def test_pair_list_forms_agree(cuda, form):
    """gcd_pairs_from_table as flag / scan / emit (0), two passes over the table (1) and one pass with decoupled look-back (2):
    the same lists, on a table that spans many look-back tiles and on tables smaller than one."""
    from gcdlss_b200 import _cabi, ops
    from gcdlss_b200 import synth
    prev = ops.get_option(_cabi.OPT_PAIRS_FUSED)
    ops.set_option(_cabi.OPT_PAIRS_FUSED, form)
    try:
        clouds = [small_cloud(1, 1, spread=0.3, batch=0), small_cloud(2, 150, spread=0.3, batch=0),
                  oq.batched_coordinates([oq.sparse_quantize_me(synth.make_scan("kitti", i, n_points=40000)[0], 0.05)[0] for i in range(2)])]
        for bc in clouds:
            for ks in (3, 5):
                table = ocd.kmap_subm(bc, ks, 1)
                pi, po, off = ocd.pairs_from_table(table)
                gi, go, goff = ops.pairs_from_table(torch.from_numpy(np.ascontiguousarray(table.T)).cuda())
                np.testing.assert_array_equal(goff.cpu().numpy(), off)
                np.testing.assert_array_equal(gi.cpu().numpy()[: off[-1]], pi)
                np.testing.assert_array_equal(go.cpu().numpy()[: off[-1]], po)
    finally:
        ops.set_option(_cabi.OPT_PAIRS_FUSED, prev)


def test_lasermix_batch_ids_and_negative_coords(cuda):
    # batch ids 0,20,40,60 (ref exp_merge_mean_teacher.py:2856 quirk) and negative coordinates
    parts = [small_cloud(10 + i, 500, spread=0.5, batch=20 * i) for i in range(4)]
    bc = np.concatenate(parts)
    check_levels(build_manager(bc), ocd.CoordLevels(bc))


def test_duplicate_and_range_errors(cuda):
    import MinkowskiEngine as ME
    c = torch.tensor([[0, 1, 2, 3], [0, 1, 2, 3]], dtype=torch.int32).cuda()
    st = ME.SparseTensor(features=torch.ones(2, 1).cuda(), coordinates=c)
    with pytest.raises(RuntimeError, match="duplicate"):
        st.C
    c = torch.tensor([[0, 1 << 20, 2, 3]], dtype=torch.int32).cuda()
    with pytest.raises(RuntimeError, match="64-bit key"):
        ME.SparseTensor(features=torch.ones(1, 1).cuda(), coordinates=c).C


def test_tiny_inputs(cuda):
    bc = np.array([[0, -3, 5, 7]], np.int32)
    mgr = build_manager(bc)
    check_levels(mgr, ocd.CoordLevels(bc))


@pytest.mark.parametrize("kernel_size", [3, 5])
def test_warp_cooperative_probing_gives_the_same_table(cuda, kernel_size):
    """gcd_kmap_subm with GCD_OPT_KMAP_COOP (four lanes per voxel, one 32-byte sector of the table per probe) against the
    one-thread-per-voxel search and the oracle: bit-identical tables, crowded table included (long chains, wrap-around)."""
    from gcdlss_b200 import _cabi, ops, synth
    from oracle import coords as ocd
    from oracle import quantize as oq
    xyz, _ = synth.make_scan("kitti", 1, n_points=40000)
    c = oq.sparse_quantize_me(xyz, 0.05)[0]
    bc = torch.from_numpy(oq.batched_coordinates([c, c[: c.shape[0] // 3] + np.array([3, -2, 1], np.int32)])).cuda()
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    table = ops.hash_build(bc, status)
    assert int(status.item()) == 0
    ref = ops.kmap_subm(bc, table, kernel_size, 1)
    ops.set_option(_cabi.OPT_KMAP_COOP, 1)
    try:
        got = ops.kmap_subm(bc, table, kernel_size, 1)
    finally:
        ops.set_option(_cabi.OPT_KMAP_COOP, 0)
    assert torch.equal(got, ref)
    small = bc[:3000].cpu().numpy()
    table_s = ops.hash_build(bc[:3000].contiguous(), status)
    ops.set_option(_cabi.OPT_KMAP_COOP, 1)
    try:
        got_s = ops.kmap_subm(bc[:3000].contiguous(), table_s, kernel_size, 1)
    finally:
        ops.set_option(_cabi.OPT_KMAP_COOP, 0)
    np.testing.assert_array_equal(got_s.cpu().numpy().T, ocd.kmap_subm(small, kernel_size, 1))
