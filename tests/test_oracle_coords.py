"""Oracle coordinate / kernel maps: brute force, dense conv3d equivalence and frozen fixtures."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from conftest import small_cloud
from oracle import conv as oc
from oracle import coords as ocd


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10**6), st.integers(1, 120), st.sampled_from([3, 5]), st.sampled_from([1, 2, 4]))
def test_subm_map_equals_brute_force(seed, n, k, ts):
    c = small_cloud(seed, n, spread=0.15, batch=seed % 3)
    c[:, 1:] *= ts
    nbr = ocd.kmap_subm(c, k, ts)
    np.testing.assert_array_equal(nbr, ocd.brute_force_subm(c, k, ts))
    kv = k ** 3
    np.testing.assert_array_equal(nbr[:, kv // 2], np.arange(c.shape[0]))          # centre tap = identity
    for kk in range(kv):                                                           # symmetry of a stride-1 map
        o = np.nonzero(nbr[:, kk] >= 0)[0]
        np.testing.assert_array_equal(nbr[nbr[o, kk], kv - 1 - kk], o)


def test_offsets_order_x_fastest():
    offs = ocd.kernel_offsets(3)
    np.testing.assert_array_equal(offs[0], [-1, -1, -1])
    np.testing.assert_array_equal(offs[1], [0, -1, -1])
    np.testing.assert_array_equal(offs[3], [-1, 0, -1])
    np.testing.assert_array_equal(offs[9], [-1, -1, 0])
    np.testing.assert_array_equal(ocd.kernel_offsets(2)[[0, 1, 2, 4]], [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10**6), st.integers(1, 200), st.sampled_from([1, 2, 8]))
def test_stride2_consistency(seed, n, ts):
    c = small_cloud(seed, n, spread=0.3, batch=0)
    c[:, 1:] *= ts
    coarse, parent, code = ocd.stride2(c, ts)
    s = 2 * ts
    np.testing.assert_array_equal(coarse[parent][:, 1:], np.floor_divide(c[:, 1:], s) * s)   # floor for negatives
    np.testing.assert_array_equal(coarse[parent][:, 0], c[:, 0])
    assert np.unique(coarse, axis=0).shape[0] == coarse.shape[0]
    offs = ocd.kernel_offsets(2)
    np.testing.assert_array_equal(coarse[parent][:, 1:] + offs[code] * ts, c[:, 1:])         # code <-> offset (x fastest)
    first = np.full(coarse.shape[0], -1)
    for f in range(c.shape[0] - 1, -1, -1):
        first[parent[f]] = f
    assert np.all(np.diff(first) > 0)                                                         # first-occurrence numbering
    down, up = ocd.kmap_down2(parent, code, coarse.shape[0]), ocd.kmap_up2(parent, code)
    assert (down >= 0).sum() == c.shape[0] == (up >= 0).sum()
    f = np.arange(c.shape[0])
    np.testing.assert_array_equal(down[parent, code], f)
    np.testing.assert_array_equal(up[f, code], parent)


def test_frozen_maps(oracle_frozen):
    lv = ocd.CoordLevels(oracle_frozen["map_coords0"])
    for l in range(1, 5):
        np.testing.assert_array_equal(lv.coords[l], oracle_frozen[f"map_coords{l}"])
        np.testing.assert_array_equal(lv.parent[l - 1], oracle_frozen[f"map_parent{l - 1}"])
        np.testing.assert_array_equal(lv.code[l - 1], oracle_frozen[f"map_code{l - 1}"])
    for l in range(5):
        np.testing.assert_array_equal(lv.subm(l, 3), oracle_frozen[f"map_subm3_{l}"])
    np.testing.assert_array_equal(lv.subm(0, 5), oracle_frozen["map_subm5_0"])


def test_pairs_from_table():
    c = small_cloud(3, 200, spread=0.2, batch=0)
    nbr = ocd.kmap_subm(c, 3, 1)
    pi, po, off = ocd.pairs_from_table(nbr)
    assert off[-1] == (nbr >= 0).sum()
    for k in (0, 13, 26):
        seg = slice(off[k], off[k + 1])
        assert np.all(np.diff(po[seg]) > 0)
        np.testing.assert_array_equal(nbr[po[seg], k], pi[seg])


# ---- independent pin: the same convolutions on a dense grid with torch.nn.functional ------------
@pytest.mark.parametrize("k", [3, 5])
def test_conv_equals_dense_conv3d(k):
    torch.manual_seed(0)
    c = small_cloud(7, 300, spread=0.2, batch=0)
    x = torch.randn(c.shape[0], 4, dtype=torch.float64)
    w = torch.randn(k ** 3, 4, 6, dtype=torch.float64)
    y = oc.conv_table(x, ocd.kmap_subm(c, k, 1), w)
    y_dense = oc.dense_conv_reference(c, x, w, k)
    torch.testing.assert_close(y, y_dense, rtol=1e-12, atol=1e-12)


def test_strided_and_transposed_equal_dense():
    torch.manual_seed(1)
    c = small_cloud(9, 400, spread=0.25, batch=0)
    coarse, parent, code = ocd.stride2(c, 1)
    x = torch.randn(c.shape[0], 3, dtype=torch.float64)
    w = torch.randn(8, 3, 5, dtype=torch.float64)
    y = oc.conv_table(x, ocd.kmap_down2(parent, code, coarse.shape[0]), w)
    # dense: grid aligned to even coordinates, conv3d(kernel 2, stride 2)
    lo = (np.floor_divide(c[:, 1:].min(0), 2) * 2)
    p = c[:, 1:] - lo
    shape = ((p.max(0) + 2) // 2 * 2).tolist()
    grid = torch.zeros((1, 3, *shape), dtype=torch.float64)
    grid[0, :, p[:, 0], p[:, 1], p[:, 2]] = x.t()
    wd = w.reshape(2, 2, 2, 3, 5).permute(4, 3, 2, 1, 0).contiguous()        # [Cout,Cin,kx,ky,kz]
    yd = torch.nn.functional.conv3d(grid, wd, stride=2)
    pc = (coarse[:, 1:] - lo) // 2
    torch.testing.assert_close(y, yd[0, :, pc[:, 0], pc[:, 1], pc[:, 2]].t(), rtol=1e-12, atol=1e-12)
    # transposed: conv_transpose3d of the coarse grid, read at the fine active sites
    xc = torch.randn(coarse.shape[0], 5, dtype=torch.float64)
    wt = torch.randn(8, 5, 3, dtype=torch.float64)
    yt = oc.conv_table(xc, ocd.kmap_up2(parent, code), wt)
    gridc = torch.zeros((1, 5, *[s // 2 for s in shape]), dtype=torch.float64)
    gridc[0, :, pc[:, 0], pc[:, 1], pc[:, 2]] = xc.t()
    wtd = wt.reshape(2, 2, 2, 5, 3).permute(3, 4, 2, 1, 0).contiguous()       # [Cin,Cout,kx,ky,kz]
    ytd = torch.nn.functional.conv_transpose3d(gridc, wtd, stride=2)
    torch.testing.assert_close(yt, ytd[0, :, p[:, 0], p[:, 1], p[:, 2]].t(), rtol=1e-12, atol=1e-12)


def test_conv_gradcheck_and_frozen(oracle_frozen):
    c = small_cloud(5, 40, spread=0.1, batch=0)
    nbr = ocd.kmap_subm(c, 3, 1)
    x = torch.randn(c.shape[0], 2, dtype=torch.float64, requires_grad=True)
    w = torch.randn(27, 2, 3, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: oc.conv_table(a, nbr, b), (x, w))
    lv_nbr = oracle_frozen["map_subm3_0"]
    xx = torch.tensor(oracle_frozen["conv_x"], requires_grad=True)
    ww = torch.tensor(oracle_frozen["conv_w"], requires_grad=True)
    y = oc.conv_table(xx, lv_nbr, ww)
    (y * torch.tensor(oracle_frozen["conv_g"])).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), oracle_frozen["conv_y"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(xx.grad.numpy(), oracle_frozen["conv_dx"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ww.grad.numpy(), oracle_frozen["conv_dw"], rtol=1e-12, atol=1e-10)
