"""Host emulation of the integer kernels: the per-thread device functions of csrc/runtable.cuh, compiled unchanged by
g++ (tests/emu/), against the CPU oracle.  Runs without a GPU; the same source is what nvcc builds for sm_100a, so the
logic of the CUDA path is checked here and only its hardware-specific parts (the 256-bit load, real concurrency) are
left to tests/test_gpu_zzz_runtable.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import ROOT, small_cloud
from oracle import coords as ocd

CSRC = os.path.join(ROOT, "generalized-class-discovery-for-lidar-semantic-segmentation_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="session")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "librt_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-psabi", "-shared", "-fPIC", "-I", CSRC, "-o", out, os.path.join(EMU, "runtable_emu.cpp")],
                   check=True, capture_output=True)
    lib = C.CDLL(out)
    lib.emu_runtable_slot_bytes.restype = C.c_int64
    lib.emu_count_slot_loads.restype = C.c_int64
    return lib


@pytest.fixture(scope="session")
def emu_gather(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "libgather_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, "-o", out, os.path.join(EMU, "gather_emu.cpp")],
                   check=True, capture_output=True)
    return C.CDLL(out)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def capacity(n):
    cap = 1024
    while cap < 2 * n:
        cap <<= 1
    return cap


def build(emu, coords, ts, order=None, cap=None):
    coords = np.ascontiguousarray(coords, np.int32)
    n = coords.shape[0]
    cap = cap or capacity(n)
    raw = np.zeros(cap * 32 + 32, np.uint8)
    off = (-raw.ctypes.data) % 32                      # the slot array must be 32-byte aligned
    slots = raw[off:off + cap * 32]
    status = np.zeros(1, np.int32)
    order_p = _ptr(np.ascontiguousarray(order, np.int64)) if order is not None else None
    emu.emu_runtable_build(_ptr(coords), C.c_int64(n), C.c_int32(ts), _ptr(slots), C.c_int64(cap), _ptr(status), order_p)
    return slots, cap, int(status[0])


def kmap(emu, coords, slots, cap, k, ts):
    coords = np.ascontiguousarray(coords, np.int32)
    n = coords.shape[0]
    nbr = np.full((k ** 3, n), -7, np.int32)
    emu.emu_kmap_subm_runs(_ptr(coords), C.c_int64(n), _ptr(slots), C.c_int64(cap), C.c_int32(k), C.c_int32(ts), _ptr(nbr))
    return nbr


def test_slot_is_one_sector(emu):
    assert emu.emu_runtable_slot_bytes() == 32


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10**6), st.integers(1, 400), st.sampled_from([3, 5]), st.sampled_from([1, 2, 4, 16]))
def test_kernel_map_equals_oracle(emu, seed, n, k, ts):
    c = small_cloud(seed, n, spread=0.2, batch=(seed % 4) * 20)        # batch ids 0/20/40/60: the LaserMix quirk
    c[:, 1:] *= ts
    slots, cap, status = build(emu, c, ts)
    assert status == 0
    np.testing.assert_array_equal(kmap(emu, c, slots, cap, k, ts), ocd.kmap_subm(c, k, ts).T)


def test_insert_order_does_not_matter(emu):
    c = small_cloud(7, 3000, spread=0.4, batch=1)
    ref = ocd.kmap_subm(c, 3, 1).T
    for seed in range(3):
        order = np.random.default_rng(seed).permutation(c.shape[0])
        slots, cap, status = build(emu, c, 1, order=order)
        assert status == 0
        np.testing.assert_array_equal(kmap(emu, c, slots, cap, 3, 1), ref)


def test_crowded_table_probes_past_collisions(emu):
    # capacity == 2n exactly and a tiny table: long probe sequences, wrap-around at the end of the array
    rng = np.random.default_rng(3)
    c = np.unique(rng.integers(-40, 40, (600, 3)).astype(np.int32), axis=0)
    c = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], 1)
    cap = 1
    while cap < 2 * c.shape[0]:
        cap <<= 1
    slots, cap, status = build(emu, c, 1, cap=cap)
    assert status == 0
    for k in (3, 5):
        np.testing.assert_array_equal(kmap(emu, c, slots, cap, k, 1), ocd.kmap_subm(c, k, 1).T)


def test_multi_batch_and_negative_coordinates(emu):
    parts = [small_cloud(s, 500, spread=0.3, batch=b) for s, b in ((1, 0), (2, 1), (3, 5))]
    c = np.concatenate(parts)
    c[:, 1:] -= 37                                      # runs straddle zero: floor semantics of cell >> 2
    slots, cap, status = build(emu, c, 1)
    assert status == 0
    for k in (3, 5):
        np.testing.assert_array_equal(kmap(emu, c, slots, cap, k, 1), ocd.kmap_subm(c, k, 1).T)


def test_coarse_levels_of_a_scan(emu):
    from gcdlss_b200 import synth
    from oracle import quantize as oq
    xyz, _ = synth.make_scan("kitti", 3, n_points=30000)
    c = oq.batched_coordinates([oq.sparse_quantize_me(xyz, 0.05)[0]])
    ts = 1
    for level in range(4):
        slots, cap, status = build(emu, c, ts)
        assert status == 0
        ref = ocd.kmap_subm(c, 3, ts).T
        np.testing.assert_array_equal(kmap(emu, c, slots, cap, 3, ts), ref)
        if level == 0:
            np.testing.assert_array_equal(kmap(emu, c, slots, cap, 5, ts), ocd.kmap_subm(c, 5, ts).T)
            # the point of the layout: scattered 32-byte loads per voxel (the point-wise table needs 26 key loads plus a
            # value load per hit for K = 3, 124 + hits for K = 5)
            n = c.shape[0]
            l3 = emu.emu_count_slot_loads(_ptr(c), C.c_int64(n), _ptr(slots), C.c_int64(cap), C.c_int32(3), C.c_int32(ts)) / n
            l5 = emu.emu_count_slot_loads(_ptr(c), C.c_int64(n), _ptr(slots), C.c_int64(cap), C.c_int32(5), C.c_int32(ts)) / n
            hits3 = (ref >= 0).sum() / n - 1
            print(f"slot loads per voxel: K=3 {l3:.1f} (point-wise table: {26 + hits3:.1f}+), K=5 {l5:.1f}")
            assert l3 < 20 and l5 < 72     # 9 x 1.5 and 25 x 2 first probes, x ~1.3 for linear-probing collisions at this load
        c = ocd.stride2(c, ts)[0]
        ts *= 2


def test_edge_of_the_key_range(emu):
    lim = 1 << 17
    c = np.array([[0, lim - 1, 5, 5], [0, lim - 2, 5, 5], [0, -lim, 5, 5], [0, -lim + 1, 5, 5],
                  [0, 3, lim - 1, -lim], [0, 3, lim - 2, -lim], [1022, 0, 0, 0]], np.int32)
    slots, cap, status = build(emu, c, 1)
    assert status == 0
    for k in (3, 5):
        np.testing.assert_array_equal(kmap(emu, c, slots, cap, k, 1), ocd.kmap_subm(c, k, 1).T)


def test_status_bits(emu):
    dup = np.array([[0, 1, 2, 3], [0, 4, 4, 4], [0, 1, 2, 3]], np.int32)
    assert build(emu, dup, 1)[2] == 2                                   # GCD_DEV_DUPLICATE
    far = np.array([[0, 1 << 17, 0, 0]], np.int32)
    assert build(emu, far, 1)[2] == 1                                   # GCD_DEV_KEY_RANGE
    assert build(emu, np.array([[1023, 0, 0, 0]], np.int32), 1)[2] == 1  # batch 1023 is reserved
    odd = np.array([[0, 3, 0, 0]], np.int32)
    assert build(emu, odd, 2)[2] == 1                                   # x is not a multiple of the tensor stride


def test_empty_and_single(emu):
    e = np.zeros((0, 4), np.int32)
    slots, cap, status = build(emu, e, 1)
    assert status == 0 and kmap(emu, e, slots, cap, 3, 1).shape == (27, 0)
    one = np.array([[2, -5, 7, 9]], np.int32)
    slots, cap, status = build(emu, one, 1)
    nbr = kmap(emu, one, slots, cap, 3, 1)
    assert status == 0 and nbr[13, 0] == 0 and (np.delete(nbr[:, 0], 13) == -1).all()


@pytest.mark.parametrize("n_in,n_out,c,pad_in,pad_out", [(50, 1, 4, 0, 0), (300, 1000, 96, 0, 0), (300, 1037, 96, 4, 8), (17, 4099, 100, 0, 12),
                                                           (9, 513, 256, 0, 0), (64, 2048, 16, 0, 0), (5, 0, 96, 0, 0)])
def test_flat_gather_is_a_row_copy(emu_gather, n_in, n_out, c, pad_in, pad_out):
    # devoxelise gather (csrc/gather_rows.cuh): every element of out[i, :c] == in[idx[i], :c], padding columns untouched
    rng = np.random.default_rng(n_out + c)
    src = rng.normal(size=(n_in, c + pad_in)).astype(np.float32)
    idx = rng.integers(0, n_in, n_out).astype(np.int64)
    out = np.full((max(n_out, 1), c + pad_out), -3.0, np.float32)
    emu_gather.emu_rows_gather_flat(_ptr(src), C.c_int64(c + pad_in), _ptr(idx), C.c_int64(n_out), C.c_int32(c), _ptr(out), C.c_int64(c + pad_out))
    np.testing.assert_array_equal(out[:n_out, :c], src[idx][:, :c])
    assert (out[:n_out, c:] == -3.0).all() and (out[n_out:] == -3.0).all()


@pytest.fixture(scope="session")
def emu_loss(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "libloss_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, "-o", out, os.path.join(EMU, "loss_emu.cpp")],
                   check=True, capture_output=True)
    return C.CDLL(out)


@pytest.mark.parametrize("n,c,scale", [(1, 2, 1.0), (257, 17, 3.0), (1000, 20, 8.0), (64, 3, 0.01)])
def test_consistency_rows_against_torch(emu_loss, n, c, scale):
    # gcd_consistency_rows (csrc/loss_rows.cuh) vs softmax / mse_loss / max of torch, forward and gradient
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(n + c)
    ls = (torch.randn(n, c, generator=g) * scale).requires_grad_(True)
    lt = torch.randn(n, c, generator=g) * scale
    lt[0, :] = lt[0, 0]                                            # a tie: the first maximum wins
    ps, pt = F.softmax(ls, 1), F.softmax(lt, 1)
    ref_sq = ((ps - pt) ** 2).sum(1)
    ref_prob, ref_label = torch.max(pt, 1)
    ref_sq.sum().backward()
    sq = np.zeros(n, np.float32); mp = np.zeros(n, np.float32); lab = np.zeros(n, np.int64); grad = np.zeros((n, c), np.float32)
    a, b = ls.detach().numpy().copy(), lt.numpy().copy()
    emu_loss.emu_consistency_rows(_ptr(a), C.c_int64(c), _ptr(b), C.c_int64(c), C.c_int64(n), C.c_int32(c), C.c_float(0.9), _ptr(sq), _ptr(mp),
                                  _ptr(lab), _ptr(grad), C.c_int64(c))
    np.testing.assert_allclose(sq, ref_sq.detach().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(mp, ref_prob.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(grad, ls.grad.numpy(), rtol=1e-4, atol=2e-6)      # fp32 cancellation noise of both sides
    expect = np.where(ref_prob.numpy() < 0.9, -1, ref_label.numpy())
    sure = np.abs(ref_prob.numpy() - 0.9) > 1e-6                   # rows within rounding of the threshold may fall either way
    np.testing.assert_array_equal(lab[sure], expect[sure])
    assert lab[0] in (-1, 0)
    # mse_loss itself
    assert abs(sq.sum() / (n * c) - float(F.mse_loss(ps, pt).detach())) < 1e-6
