"""Tile sort of the 3x3x3 neighbour tables (csrc/tilesort.cuh, opt-in GCDLSS_TILE_SORT=1).

* the per-thread device code (key, permute), compiled unchanged by g++ (tests/emu/tilesort_emu.cpp), against numpy;
* the property the convolution relies on: gathering through the sorted table and scattering through ``out_rows`` gives
  the same result as the table in scan order (checked with the CPU oracle's sparse convolution);
* what the sort buys on synthetic sweeps: kernel offsets with a hit per 128-column tile, share of real rows.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, small_cloud
from oracle import conv as oc
from oracle import coords as ocd
from oracle import quantize as oq

CSRC = os.path.join(ROOT, "generalized-class-discovery-for-lidar-semantic-segmentation_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="session")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "libtilesort_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, "-o", out, os.path.join(EMU, "tilesort_emu.cpp")],
                   check=True, capture_output=True)
    return C.CDLL(out)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def bit_order():
    """bit[k] of offset k: offsets ranked by (number of non-zero components, k) -- centre lowest, corners on top."""
    offs = ocd.kernel_offsets(3)
    cls = (offs != 0).sum(1)
    rank = np.lexsort((np.arange(27), cls))
    bit = np.empty(27, np.int64)
    bit[rank] = np.arange(27)
    return bit


def sort_reference(nbr_cols):
    """nbr_cols [27 | 8, n] -> (sorted table, rows, keys) in numpy."""
    valid = (nbr_cols >= 0).astype(np.uint64)
    bits = bit_order() if nbr_cols.shape[0] == 27 else np.arange(8)
    keys = (valid << bits.astype(np.uint64)[:, None]).sum(0)
    rows = np.argsort(keys, kind="stable").astype(np.int32)
    return nbr_cols[:, rows], rows, keys[rows]


def emu_sort(emu, nbr_cols):
    nbr_cols = np.ascontiguousarray(nbr_cols, np.int32)
    kv, n = nbr_cols.shape
    out = np.full_like(nbr_cols, -9)
    rows = np.full(max(n, 1), -9, np.int32)
    keys = np.zeros(max(n, 1), np.uint64)
    emu.emu_kmap_tile_sort(_ptr(nbr_cols), C.c_int64(n), C.c_int32(kv), _ptr(out), _ptr(rows), _ptr(keys))
    return out, rows[:n], keys[:n]


def tile_masks_reference(sorted_cols):
    """Per 128-column tile: bit k set iff some column of the tile has a neighbour at offset k (numpy)."""
    kv, n = sorted_cols.shape
    out = np.zeros((n + 127) // 128, np.uint32)
    for t in range(out.shape[0]):
        present = (sorted_cols[:, 128 * t:128 * t + 128] >= 0).any(1)
        out[t] = sum(1 << k for k in range(kv) if present[k])
    return out.view(np.int32)


def emu_tile_masks(emu, sorted_keys, kv):
    """The device code's mask of each tile from the OR of its sorted keys (csrc/tilesort.cuh:tile_mask_from_keys)."""
    emu.emu_tile_mask_from_keys.restype = C.c_uint32
    n = sorted_keys.shape[0]
    out = np.zeros((n + 127) // 128, np.uint32)
    for t in range(out.shape[0]):
        out[t] = emu.emu_tile_mask_from_keys(C.c_ulonglong(int(np.bitwise_or.reduce(sorted_keys[128 * t:128 * t + 128]))), C.c_int32(kv))
    return out.view(np.int32)


def test_tile_masks_from_sorted_keys(emu):
    c = small_cloud(17, 3000, spread=0.4, batch=0)
    coarse, parent, code = ocd.stride2(c, 1)
    for cols in (ocd.kmap_subm(c, 3, 1).T, ocd.kmap_down2(parent, code, coarse.shape[0]).T, ocd.kmap_up2(parent, code).T):
        got, rows, keys = emu_sort(emu, np.ascontiguousarray(cols))
        np.testing.assert_array_equal(emu_tile_masks(emu, keys, cols.shape[0]), tile_masks_reference(got))


def test_bit_order(emu):
    bits = [emu.emu_tile_sort_bit(k) for k in range(27)]
    assert bits == list(bit_order())
    assert bits[13] == 0 and sorted(bits) == list(range(27))                 # centre lowest; a permutation of 0..26
    offs = ocd.kernel_offsets(3)
    assert all(bits[k] >= 19 for k in range(27) if (offs[k] != 0).all())       # the eight corners take the top bits


@pytest.mark.parametrize("seed,n,ts", [(0, 1, 1), (1, 77, 1), (2, 3000, 1), (3, 2500, 2), (4, 129, 8)])
def test_emulated_sort_equals_numpy(emu, seed, n, ts):
    c = small_cloud(seed, n, spread=0.4, batch=seed % 2)
    c[:, 1:] *= ts
    nbr = np.ascontiguousarray(ocd.kmap_subm(c, 3, ts).T)
    got, rows, keys = emu_sort(emu, nbr)
    ref, ref_rows, ref_keys = sort_reference(nbr)
    np.testing.assert_array_equal(rows, ref_rows)
    np.testing.assert_array_equal(keys, ref_keys)
    np.testing.assert_array_equal(got, ref)
    assert sorted(rows.tolist()) == list(range(nbr.shape[1]))


def test_stride2_tables(emu):
    # 2x2x2 maps: a transposed (up) map has one entry per column -> exactly one offset per tile after the sort
    c = small_cloud(21, 4000, spread=0.5, batch=0)
    coarse, parent, code = ocd.stride2(c, 1)
    for nbr in (ocd.kmap_down2(parent, code, coarse.shape[0]), ocd.kmap_up2(parent, code)):
        cols = np.ascontiguousarray(nbr.T)
        got, rows, keys = emu_sort(emu, cols)
        ref, ref_rows, ref_keys = sort_reference(cols)
        np.testing.assert_array_equal(rows, ref_rows)
        np.testing.assert_array_equal(got, ref)
        a0, r0 = tile_stats(cols)
        a1, r1 = tile_stats(got)
        assert a1 < a0 and r1 > r0
    assert a1 < 1.3 and r1 > 0.75                                             # the up map


def test_sorted_table_gives_the_same_convolution(emu):
    # what conv_fwd_tc_kernel<.., kPerm> does: gather through the sorted table, write column i to row out_rows[i]
    c = small_cloud(11, 1500, spread=0.35, batch=0)
    nbr = ocd.kmap_subm(c, 3, 1)                                             # [n, 27]
    n = nbr.shape[0]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 8, dtype=torch.float64, generator=g)
    w = torch.randn(27, 8, 5, dtype=torch.float64, generator=g)
    ref = oc.conv_table(x, nbr, w)
    sorted_cols, rows, _ = emu_sort(emu, np.ascontiguousarray(nbr.T))
    out_sorted = oc.conv_table(x, np.ascontiguousarray(sorted_cols.T), w)  # row i of this is output row rows[i]
    out = torch.empty_like(ref)
    out[torch.from_numpy(rows.astype(np.int64))] = out_sorted
    torch.testing.assert_close(out, ref, rtol=0, atol=1e-12)


def tile_stats(nbr_cols, tile=128):
    kv, n = nbr_cols.shape
    pad = (-n) % tile
    v = np.concatenate([nbr_cols >= 0, np.zeros((kv, pad), bool)], 1).reshape(kv, -1, tile)
    active = v.any(2)                                                        # [kv, tiles]
    return active.sum(0).mean(), v.sum() / (active.sum() * tile)


def test_what_the_sort_buys_on_synthetic_sweeps(emu):
    from gcdlss_b200 import synth
    for kind, q, n_scans in (("kitti", 0.05, 1), ("nuscenes", 0.1, 2)):
        scans = [oq.sparse_quantize_me(synth.make_scan(kind, i)[0], q)[0] for i in range(n_scans)]
        lv = ocd.CoordLevels(oq.batched_coordinates(scans))
        for level in (0, 1):
            nbr = np.ascontiguousarray(lv.subm(level, 3).T)
            a0, r0 = tile_stats(nbr)
            a1, r1 = tile_stats(emu_sort(emu, nbr)[0])
            print(f"{kind} level {level}: {nbr.shape[1]} rows; offsets per tile {a0:.1f} -> {a1:.1f}, real rows {100 * r0:.0f} % -> {100 * r1:.0f} %")
            assert a1 < 0.6 * a0 and r1 > 1.8 * r0


def test_kernel_map_hands_out_the_sorted_table(emu, monkeypatch):
    # host logic of coords.KernelMap.tc_table / tc_back_table (no GPU: the sort itself is the emulated one)
    import gcdlss_b200
    from gcdlss_b200 import config, coords, ops

    prev = (gcdlss_b200.get_tile_sort(), config.tile_sort_min_rows())
    calls = []

    def fake_sort(nbr):
        calls.append(nbr.shape)
        s, rows, keys = emu_sort(emu, nbr.numpy())
        return torch.from_numpy(s), torch.from_numpy(rows), torch.from_numpy(emu_tile_masks(emu, keys, nbr.shape[0]))

    monkeypatch.setattr(ops, "kmap_tile_sort", fake_sort)

    c = small_cloud(5, 800, spread=0.3, batch=0)
    nbr = torch.from_numpy(np.ascontiguousarray(ocd.kmap_subm(c, 3, 1).T))
    n = nbr.shape[1]
    t3 = coords.NeighbourTable(nbr, 27, n)
    t_dn = coords.NeighbourTable(torch.zeros((8, 10), dtype=torch.int32), 8, 10)
    t_up = coords.NeighbourTable(torch.zeros((8, n), dtype=torch.int32), 8, n)
    km3 = coords.KernelMap(t3, n, n, t3, True)
    km_down = coords.KernelMap(t_dn, n, 10, t_up, False)        # a stride-2 map and its transpose share their tables
    km_up = coords.KernelMap(t_up, 10, n, t_dn, False)
    km_1x1 = coords.KernelMap(coords.NeighbourTable(None, 1, n), n, n, None, False)
    try:
        gcdlss_b200.set_tile_sort(False)
        assert km3.tc_table() == (km3.nbr, None, None) and not calls
        gcdlss_b200.set_tile_sort(True, min_rows=n + 1)
        assert km3.tc_table()[1] is None and not calls                      # too few rows to pay for the sort
        gcdlss_b200.set_tile_sort(True, min_rows=1)
        table, rows, masks = km3.tc_table()
        assert rows is not None and torch.equal(table, nbr[:, rows.long()]) and len(calls) == 1
        assert masks.tolist() == tile_masks_reference(table.numpy()).tolist()
        assert km3.tc_table()[0] is table and len(calls) == 1               # cached
        assert km3.tc_back_table()[0] is table                              # stride-1 maps are self-transposed
        t_down, r_down, _ = km_down.tc_table()                              # 2x2x2 tables are sorted as well
        assert r_down is not None and torch.equal(t_down, km_down.nbr[:, r_down.long()])
        assert km_up.tc_back_table()[0] is t_down and km_down.tc_back_table()[0] is km_up.tc_table()[0]
        t5 = coords.NeighbourTable(torch.zeros((125, n), dtype=torch.int32), 125, n)
        km5 = coords.KernelMap(t5, n, n, t5, True)
        assert km5.tc_table() == (km5.nbr, None, None)                      # the 5x5x5 stem goes through im2col: not sorted
        assert km_1x1.tc_table() == (None, None, None) and km_1x1.tc_back_table() == (None, None, None)
    finally:
        gcdlss_b200.set_tile_sort(prev[0], min_rows=prev[1])


def frozen_tables(oracle_frozen):
    """name -> [kv, n] table of the frozen fixture (tests/golden/make_golden.py part 3)."""
    n1 = oracle_frozen["map_coords1"].shape[0]
    return {"subm3_0": oracle_frozen["map_subm3_0"].T, "subm3_1": oracle_frozen["map_subm3_1"].T,
            "down_0": ocd.kmap_down2(oracle_frozen["map_parent0"], oracle_frozen["map_code0"], n1).T,
            "up_0": ocd.kmap_up2(oracle_frozen["map_parent0"], oracle_frozen["map_code0"]).T}


def test_emulated_sort_against_the_frozen_permutations(emu, oracle_frozen):
    frozen = np.load(os.path.join(ROOT, "tests", "golden", "tile_sort_frozen.npz"))
    for name, table in frozen_tables(oracle_frozen).items():
        table = np.ascontiguousarray(table, np.int32)
        got, rows, keys = emu_sort(emu, table)
        np.testing.assert_array_equal(rows, frozen[f"rows_{name}"])
        np.testing.assert_array_equal(keys, frozen[f"keys_{name}"][frozen[f"rows_{name}"]])
        np.testing.assert_array_equal(got, table[:, frozen[f"rows_{name}"]])
