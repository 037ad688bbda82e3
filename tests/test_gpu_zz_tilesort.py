"""Tile-sorted neighbour tables for the tcgen05 convolution (gcd_kmap_tile_sort + gcd_conv_args.out_rows, opt-in
GCDLSS_TILE_SORT=1) on the GPU.  The sort's per-thread code and the equivalence of the sorted table are held on the CPU
by tests/test_tile_sort_model.py; new here are the radix sort on real keys, the kernel's permuted epilogue and the
plumbing through the per-op and the fused-block paths.  (Sorts last on purpose: the feature is new and opt-in.)"""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from conftest import small_cloud
from gpu_util import TOL_BF16, OpChecker, rel_err
from oracle import coords as ocd
from oracle import quantize as oq
from test_tile_sort_model import sort_reference, tile_masks_reference

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tile_sort():
    import gcdlss_b200
    from gcdlss_b200 import config
    prev = (gcdlss_b200.get_math_mode(), gcdlss_b200.get_tile_sort(), config.tile_sort_min_rows())
    gcdlss_b200.set_math_mode("bf16")
    gcdlss_b200.set_tile_sort(True, min_rows=1)
    yield
    gcdlss_b200.set_tile_sort(prev[1], min_rows=prev[2])
    gcdlss_b200.set_math_mode(prev[0])


def kitti_coords(n_scans=1, n_points=30000):
    from gcdlss_b200 import synth
    scans = [oq.sparse_quantize_me(synth.make_scan("kitti", i, n_points=n_points)[0], 0.05)[0] for i in range(n_scans)]
    return oq.batched_coordinates(scans)


@pytest.mark.parametrize("n", [1, 127, 5000])
def test_sort_equals_numpy(cuda, n):
    from gcdlss_b200 import ops
    c = small_cloud(n, n, spread=0.4, batch=0)
    nbr = np.ascontiguousarray(ocd.kmap_subm(c, 3, 1).T)
    got, rows, masks = ops.kmap_tile_sort(torch.from_numpy(nbr).cuda())
    ref, ref_rows, _ = sort_reference(nbr)
    np.testing.assert_array_equal(rows.cpu().numpy(), ref_rows)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    np.testing.assert_array_equal(masks.cpu().numpy(), tile_masks_reference(ref))


def test_sort_of_stride2_tables_equals_numpy(cuda):
    from gcdlss_b200 import ops
    c = small_cloud(3, 6000, spread=0.5, batch=0)
    coarse, parent, code = ocd.stride2(c, 1)
    for nbr in (ocd.kmap_down2(parent, code, coarse.shape[0]), ocd.kmap_up2(parent, code)):
        cols = np.ascontiguousarray(nbr.T)
        got, rows, masks = ops.kmap_tile_sort(torch.from_numpy(cols).cuda())
        ref, ref_rows, _ = sort_reference(cols)
        np.testing.assert_array_equal(rows.cpu().numpy(), ref_rows)
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
        np.testing.assert_array_equal(masks.cpu().numpy(), tile_masks_reference(ref))


def test_sort_against_the_frozen_permutations(cuda, oracle_frozen):
    import os
    from conftest import GOLDEN
    from gcdlss_b200 import ops
    from test_tile_sort_model import frozen_tables
    frozen = np.load(os.path.join(GOLDEN, "tile_sort_frozen.npz"))
    for name, table in frozen_tables(oracle_frozen).items():
        table = np.ascontiguousarray(table, np.int32)
        got, rows, _ = ops.kmap_tile_sort(torch.from_numpy(table).cuda())
        np.testing.assert_array_equal(rows.cpu().numpy(), frozen[f"rows_{name}"])
        np.testing.assert_array_equal(got.cpu().numpy(), table[:, frozen[f"rows_{name}"]])


@pytest.mark.parametrize("cin,cout", [(32, 32), (96, 96), (256, 128)])
def test_conv_forward_and_dgrad_with_sorted_table(cuda, tile_sort, cin, cout):
    from gcdlss_b200 import ops
    from gcdlss_b200.coords import CoordinateManager
    bc = kitti_coords()
    km = CoordinateManager(torch.from_numpy(bc).cuda()).kernel_map(1, 3, 1, False)
    table, rows, masks = km.tc_table()
    assert rows is not None and table.data_ptr() != km.nbr.data_ptr()
    n = km.n_out
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(27, cin, cout, device="cuda", generator=g) * 0.05
    plain = ops.conv_forward(x, km.nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=ops.pack_weights(w, False, False))
    tiled = ops.conv_forward(x, table, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=ops.pack_weights(w, False, False), out_rows=rows,
                             tile_masks=masks)
    no_masks = ops.conv_forward(x, table, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=ops.pack_weights(w, False, False), out_rows=rows)
    assert torch.equal(tiled, no_masks), "the per-tile masks only tell the kernel which slices to stage: same offsets, same order, same bits"
    ref = torch.zeros(n, cout, dtype=torch.float64, device="cuda")
    for k in range(27):
        idx = km.nbr[k].long()
        o = torch.nonzero(idx >= 0).reshape(-1)
        ref.index_add_(0, o, x.double()[idx[o]] @ w[k].double())
    print(f"{cin}->{cout}: sorted vs fp64 {rel_err(tiled, ref):.2e}, scan order vs fp64 {rel_err(plain, ref):.2e}")
    assert rel_err(tiled, ref) < TOL_BF16 and rel_err(tiled, plain) < TOL_BF16
    # dgrad orientation through the same table (mirrored offsets)
    gy = torch.randn(n, cout, device="cuda", generator=g).to(torch.bfloat16)
    pk = ops.pack_weights(w, True, True)
    d_plain = ops.conv_forward(gy, km.nbr, w, n, transpose_w=True, mirror=True, out_dtype=torch.bfloat16, math_mode=1, w_packed=pk)
    d_tiled = ops.conv_forward(gy, table, w, n, transpose_w=True, mirror=True, out_dtype=torch.bfloat16, math_mode=1, w_packed=pk, out_rows=rows,
                               tile_masks=masks)
    assert rel_err(d_tiled, d_plain) < TOL_BF16


def test_every_kernel_call_of_a_training_step(cuda, tile_sort):
    # per-launch path under OpChecker: every convolution (now through sorted tables) against an fp64 re-computation
    import MinkowskiEngine as ME
    from models import minkunet as mu
    bc = kitti_coords(2, 20000)
    torch.manual_seed(0)
    model = mu.MinkUNet14A(1, 17).cuda().train()
    f = torch.rand(bc.shape[0], 1).cuda()
    labels = torch.randint(0, 17, (bc.shape[0],)).cuda()
    with OpChecker() as chk:
        out = model(ME.SparseTensor(features=f, coordinates=torch.from_numpy(bc).cuda())).F
        torch.nn.functional.cross_entropy(out.float(), labels).backward()
    worst = chk.worst()
    print("worst op:", worst, "of", len(chk.records))
    assert worst[-1] < TOL_BF16
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_fused_blocks_match_the_scan_order_path(cuda):
    import gcdlss_b200
    import MinkowskiEngine as ME
    from models import minkunet as mu
    bc = torch.from_numpy(kitti_coords(2, 20000)).cuda()
    f = torch.rand(bc.shape[0], 1).cuda()
    labels = torch.randint(0, 17, (bc.shape[0],)).cuda()
    from gcdlss_b200 import config
    prev = (gcdlss_b200.get_math_mode(), gcdlss_b200.get_tile_sort(), config.tile_sort_min_rows())
    gcdlss_b200.set_math_mode("bf16")
    results = []
    try:
        for on in (False, True):
            gcdlss_b200.set_tile_sort(on, min_rows=1)
            torch.manual_seed(0)
            model = mu.MinkUNet14A(1, 17).cuda().train()
            out = model(ME.SparseTensor(features=f, coordinates=bc)).F
            loss = torch.nn.functional.cross_entropy(out.float(), labels)
            loss.backward()
            grads = torch.cat([p.grad.reshape(-1).float() for p in model.parameters()])
            results.append((out.detach().float(), float(loss), grads))
    finally:
        gcdlss_b200.set_tile_sort(prev[1], min_rows=prev[2])
        gcdlss_b200.set_math_mode(prev[0])
    (o0, l0, g0), (o1, l1, g1) = results
    cos = float(torch.nn.functional.cosine_similarity(g0, g1, dim=0))
    print(f"logits rel diff {rel_err(o1, o0):.2e}, loss {l0:.5f} vs {l1:.5f}, gradient cosine {cos:.4f}")
    # the two runs differ only in the order fp32 partial sums are added inside a tile (bf16 storage between layers)
    assert rel_err(o1, o0) < 5e-2 and abs(l0 - l1) < 1e-2 and cos > 0.95
