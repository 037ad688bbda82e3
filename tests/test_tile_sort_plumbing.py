"""Host plumbing of the tile-sorted convolution path without a GPU: the real ``functional.conv_fwd_impl`` /
``conv_dgrad_impl`` / ``_fill_unit`` run on CPU tensors with the library calls replaced by torch stand-ins that follow the
C ABI's contract (gather through ``nbr``, write column i to row ``out_rows[i]``).  Checks that every call gets a matching
(table, out_rows) pair -- forward, self-transposed dgrad with mirrored offsets, and the swapped tables of stride-2 /
transposed convolutions -- by comparing the result with the oracle's convolution."""
import numpy as np
import pytest
import torch

from conftest import small_cloud
from oracle import conv as oc
from oracle import coords as ocd
from test_tile_sort_model import emu, emu_sort, emu_tile_masks  # noqa: F401  (emu is a fixture)


def fake_conv_forward(inp, nbr, w3, n_out, *, transpose_w=False, mirror=False, bias=None, out_dtype=None, math_mode=0, w_packed=None,
                      stats=None, out_rows=None, tile_masks=None):
    """gcd_conv_forward's contract in torch (fp64 accumulation)."""
    kv = w3.shape[0]
    x = inp.double()
    cols = torch.zeros((n_out, w3.shape[1] if transpose_w else w3.shape[2]), dtype=torch.float64)
    for k in range(kv):
        wk = w3[kv - 1 - k if mirror else k].double()
        b = wk.t() if transpose_w else wk
        idx = nbr[k].long()
        o = torch.nonzero(idx >= 0).reshape(-1)
        cols.index_add_(0, o, x[idx[o]] @ b)
    return cols if out_rows is None else torch.zeros_like(cols).index_copy_(0, out_rows.long(), cols)     # kept in fp64 for the comparison


@pytest.fixture()
def patched(emu, monkeypatch):  # noqa: F811
    import gcdlss_b200
    from gcdlss_b200 import config, functional, ops
    prev = (gcdlss_b200.get_math_mode(), gcdlss_b200.get_tile_sort(), config.tile_sort_min_rows())
    calls = []

    def conv_forward(*a, **kw):
        calls.append(kw.get("out_rows") is not None)
        return fake_conv_forward(*a, **kw)

    def tile_sort(nbr):
        s, rows, keys = emu_sort(emu, nbr.numpy())
        return torch.from_numpy(s), torch.from_numpy(rows), torch.from_numpy(emu_tile_masks(emu, keys, nbr.shape[0]))

    monkeypatch.setattr(ops, "conv_forward", conv_forward)
    monkeypatch.setattr(ops, "kmap_tile_sort", tile_sort)
    monkeypatch.setattr(ops, "pack_weights", lambda w, t, m: None)
    monkeypatch.setattr(ops, "pairs_from_table", lambda nbr: tuple(torch.from_numpy(a) for a in ocd.pairs_from_table(nbr.numpy().T)))
    gcdlss_b200.set_math_mode("bf16")
    gcdlss_b200.set_tile_sort(True, min_rows=1)
    yield functional, calls
    gcdlss_b200.set_tile_sort(prev[1], min_rows=prev[2])
    gcdlss_b200.set_math_mode(prev[0])


def scene():
    from gcdlss_b200 import coords
    c = small_cloud(31, 1200, spread=0.35, batch=0)
    coarse, parent, code = ocd.stride2(c, 1)
    n, m = c.shape[0], coarse.shape[0]
    t3, tdown, tup = ocd.kmap_subm(c, 3, 1), ocd.kmap_down2(parent, code, m), ocd.kmap_up2(parent, code)
    as_cols = lambda t: torch.from_numpy(np.ascontiguousarray(t.T))          # [kv, n_out] as the product stores it
    n3, ndn, nup = coords.NeighbourTable(as_cols(t3), 27, n), coords.NeighbourTable(as_cols(tdown), 8, m), coords.NeighbourTable(as_cols(tup), 8, n)
    km3 = coords.KernelMap(n3, n, n, n3, True)
    km_down = coords.KernelMap(ndn, n, m, nup, False)
    km_up = coords.KernelMap(nup, m, n, ndn, False)
    return None, (km3, t3), (km_down, tdown), (km_up, tup)


@pytest.mark.parametrize("which", ["subm", "down", "up"])
def test_forward_and_dgrad_through_sorted_tables(patched, which):
    functional, calls = patched
    mgr, s3, sd, su = scene()
    km, table = {"subm": s3, "down": sd, "up": su}[which]
    kv = km.kv
    g = torch.Generator().manual_seed(3)
    x = torch.randn(km.n_in, 16, generator=g).to(torch.bfloat16)
    w = (torch.randn(kv, 16, 32, generator=g) * 0.1)
    y = functional.conv_fwd_impl(x, w, None, km, torch.bfloat16, None)
    assert calls[-1], "the forward call must carry out_rows"
    ref = oc.conv_table(x.double(), table, w.double())
    torch.testing.assert_close(y, ref, rtol=0, atol=1e-9)
    # dgrad: d/dx of sum(y * gy) == conv through the transposed map with W^T (mirrored offsets for stride-1 maps)
    gy = torch.randn(km.n_out, 32, generator=g).to(torch.bfloat16)
    xr = x.double().requires_grad_(True)
    (oc.conv_table(xr, table, w.double()) * gy.double()).sum().backward()
    dx = functional.conv_dgrad_impl(gy, w, km, torch.bfloat16, None)
    assert calls[-1], "the dgrad call must carry out_rows"
    torch.testing.assert_close(dx, xr.grad, rtol=0, atol=1e-9)


def test_fill_unit_points_at_the_sorted_tables(patched, monkeypatch):
    functional, _ = patched
    from gcdlss_b200._cabi import ConvBnUnit
    mgr, (km3, _), (km_down, _), (km_up, _) = scene()

    class Holder:                       # what a convolution module carries for the tcgen05 path
        _pk_mirror = True
        _pk_fwd = torch.zeros(8, dtype=torch.uint8)
        _pk_bwd = torch.zeros(8, dtype=torch.uint8)

    monkeypatch.setattr(functional.packed_weights, "ensure", lambda holder: None)
    bn = torch.nn.BatchNorm1d(32)
    w = torch.zeros(27, 16, 32)
    u = ConvBnUnit()
    functional._fill_unit(u, w, bn.weight, bn.bias, bn, km3, Holder, True, True)
    table, rows, masks = km3.tc_table()
    assert (u.nbr, u.out_rows, u.tile_masks) == (table.data_ptr(), rows.data_ptr(), masks.data_ptr())
    assert (u.back_nbr, u.back_out_rows, u.back_tile_masks, u.back_mirror) == (table.data_ptr(), rows.data_ptr(), masks.data_ptr(), 1)
    assert u.pair_in == km3.pairs[0].data_ptr() and u.n_pairs == km3.pairs[0].shape[0]      # pair lists of the ORIGINAL table
    Holder._pk_mirror = False
    u = ConvBnUnit()
    functional._fill_unit(u, torch.zeros(8, 16, 32), bn.weight, bn.bias, bn, km_down, Holder, True, True)
    t_down, r_down, m_down = km_down.tc_table()
    t_up, r_up, m_up = km_up.tc_table()
    assert (u.nbr, u.out_rows, u.back_nbr, u.back_out_rows) == (t_down.data_ptr(), r_down.data_ptr(), t_up.data_ptr(), r_up.data_ptr())
    assert (u.tile_masks, u.back_tile_masks) == (m_down.data_ptr(), m_up.data_ptr())
    # SIMT units (tc False) keep the scan-order table and no row map
    u = ConvBnUnit()
    functional._fill_unit(u, w, bn.weight, bn.bias, bn, km3, Holder, False, True)
    assert (u.nbr, u.out_rows, u.back_out_rows, u.tile_masks) == (km3.nbr.data_ptr(), None, None, None)
