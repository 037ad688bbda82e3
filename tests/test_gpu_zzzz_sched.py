"""Scheduling features of the tcgen05 path on the GPU: the dynamic tile schedule of the two persistent kernels
(gcd_conv_args.sched / gcd_wgrad_args.sched), the execution context of gcd_run_ops_exec (weight gradients on a second stream)
and programmatic dependent launch.  None of them may change a result: a tile's arithmetic does not depend on the CTA that
claims it, so forward / dgrad outputs are compared bit for bit; weight gradients are atomic sums in every mode and are
compared within the summation-order noise."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from gpu_util import rel_err
from oracle import quantize as oq

pytestmark = pytest.mark.gpu


def kitti_coords(n_scans=1, n_points=30000):
    from gcdlss_b200 import synth
    scans = [oq.sparse_quantize_me(synth.make_scan("kitti", i, n_points=n_points)[0], 0.05)[0] for i in range(n_scans)]
    return oq.batched_coordinates(scans)


@pytest.mark.parametrize("cin,cout,ts", [(32, 32, 1), (96, 96, 1), (256, 256, 4), (128, 64, 2)])
def test_dynamic_tile_schedule_is_bit_identical(cuda, cin, cout, ts):
    from gcdlss_b200 import ops
    from gcdlss_b200.coords import CoordinateManager
    km = CoordinateManager(torch.from_numpy(kitti_coords(2)).cuda()).kernel_map(ts, 3, 1, False)
    n = km.n_out
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(n, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(27, cin, cout, device="cuda", generator=g) * 0.05
    pk = ops.pack_weights(w, False, False)
    sorted_table, rows, masks = ops.kmap_tile_sort(km.nbr)
    sched = torch.zeros(2, dtype=torch.int32, device="cuda")
    for table, kw in ((km.nbr, {}), (sorted_table, dict(out_rows=rows, tile_masks=masks)), (sorted_table, dict(out_rows=rows))):
        static = ops.conv_forward(x, table, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=pk, **kw)
        for _ in range(3):          # the counters are shared by consecutive launches of a stream: zero again after each
            dynamic = ops.conv_forward(x, table, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=pk, sched=sched, **kw)
            assert torch.equal(static, dynamic)
            assert sched.tolist() == [0, 0]
    # fp32 output (the head), identity map (1x1 convolution)
    w1 = torch.randn(1, cin, cout, device="cuda", generator=g) * 0.05
    pk1 = ops.pack_weights(w1, False, False)
    a = ops.conv_forward(x, None, w1, n, out_dtype=torch.float32, math_mode=1, w_packed=pk1)
    b = ops.conv_forward(x, None, w1, n, out_dtype=torch.float32, math_mode=1, w_packed=pk1, sched=sched)
    assert torch.equal(a, b) and sched.tolist() == [0, 0]


def test_dynamic_schedule_with_fewer_tiles_than_ctas_and_empty_input(cuda):
    from gcdlss_b200 import ops
    from oracle import coords as ocd
    from conftest import small_cloud
    sched = torch.zeros(2, dtype=torch.int32, device="cuda")
    for n_pts in (1, 100, 300):
        c = small_cloud(n_pts, n_pts, spread=0.3, batch=0)
        nbr = torch.from_numpy(np.ascontiguousarray(ocd.kmap_subm(c, 3, 1).T)).cuda()
        n = nbr.shape[1]
        x = torch.randn(n, 64, device="cuda").to(torch.bfloat16)
        w = torch.randn(27, 64, 64, device="cuda") * 0.05
        pk = ops.pack_weights(w, False, False)
        a = ops.conv_forward(x, nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=pk)
        b = ops.conv_forward(x, nbr, w, n, out_dtype=torch.bfloat16, math_mode=1, w_packed=pk, sched=sched)
        assert torch.equal(a, b) and sched.tolist() == [0, 0]


@pytest.mark.parametrize("cin,cout", [(32, 32), (96, 96), (256, 128)])
def test_dynamic_wgrad_schedule(cuda, cin, cout):
    from gcdlss_b200 import ops
    from gcdlss_b200.coords import CoordinateManager
    km = CoordinateManager(torch.from_numpy(kitti_coords(2)).cuda()).kernel_map(1, 3, 1, False)
    n = km.n_out
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, cin, device="cuda", generator=g).to(torch.bfloat16)
    gy = torch.randn(n, cout, device="cuda", generator=g).to(torch.bfloat16)
    sched = torch.zeros(2, dtype=torch.int32, device="cuda")
    static = torch.zeros(27, cin, cout, device="cuda")
    ops.conv_wgrad(x, gy, km.pairs, 27, static, math_mode=1)
    for _ in range(2):
        dynamic = torch.zeros(27, cin, cout, device="cuda")
        ops.conv_wgrad(x, gy, km.pairs, 27, dynamic, math_mode=1, sched=sched)
        assert sched.tolist() == [0, 0]
        assert rel_err(dynamic, static) < 1e-5
    # identity map (1x1 convolution): one offset, rows as pairs
    s1, d1 = torch.zeros(1, cin, cout, device="cuda"), torch.zeros(1, cin, cout, device="cuda")
    ops.conv_wgrad(x, gy, None, 1, s1, math_mode=1)
    ops.conv_wgrad(x, gy, None, 1, d1, math_mode=1, sched=sched)
    assert rel_err(d1, s1) < 1e-5 and sched.tolist() == [0, 0]


def test_execution_context_modes_give_the_same_step(cuda):
    """The trunk (gcd_run_ops_exec) with the weight gradients on the second stream (both event placements) or on the caller's
    stream, with dynamic or static tiles, with and without programmatic dependent launch: identical stage outputs, the same
    gradients."""
    import gcdlss_b200
    import MinkowskiEngine as ME
    from gcdlss_b200 import _cabi, ops
    from models import minkunet as mu
    bc = torch.from_numpy(kitti_coords(2, 20000)).cuda()
    f = torch.rand(bc.shape[0], 1).cuda()
    labels = torch.randint(0, 17, (bc.shape[0],), device="cuda")
    prev_mode = gcdlss_b200.get_math_mode()
    opts = (_cabi.OPT_WGRAD_SIDE, _cabi.OPT_DYN_TILES, _cabi.OPT_PDL)
    prev = [ops.get_option(o) for o in opts]
    gcdlss_b200.set_math_mode("bf16")
    torch.manual_seed(0)
    model = mu.MinkUNet34C(1, 17).cuda().train()
    results = []
    try:
        for side, dyn, pdl in ((0, 0, 0), (1, 1, 1), (2, 1, 1), (1, 0, 1), (0, 1, 0), (1, 1, 0)):
            for o, v in zip(opts, (side, dyn, pdl)):
                ops.set_option(o, v)
            model.zero_grad(set_to_none=True)
            stages = model._trunk(ME.SparseTensor(features=f, coordinates=bc))
            logits = model.final(stages[7]).F
            (torch.nn.functional.cross_entropy(logits.float(), labels) + 0.01 * stages[3].F.float().square().mean()).backward()
            torch.cuda.synchronize()
            results.append(([s.F.detach().clone() for s in stages], torch.cat([p.grad.flatten().float() for p in model.parameters()])))
    finally:
        for o, v in zip(opts, prev):
            ops.set_option(o, v)
        gcdlss_b200.set_math_mode(prev_mode)
    s0, g0 = results[0]
    assert torch.isfinite(g0).all()
    for s, g in results[1:]:
        for a, b in zip(s, s0):
            assert torch.equal(a, b)
        cos = float(torch.nn.functional.cosine_similarity(g, g0, dim=0))
        assert cos > 0.99999, cos
