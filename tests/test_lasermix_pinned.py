"""LaserMix of the Stage-2 step against the REFERENCE's own code: tests/golden/make_golden.py (part 4) executes
``mix_transform`` / ``laser_mix_transform`` of /root/reference/modules/exp_merge_mean_teacher.py:1577-1787 unchanged and
freezes what they return; ``gcdlss_b200.steps.mix_transform`` must reproduce the arrays row for row (order of the rows
decides which point represents a voxel after the first-occurrence quantiser, ref :2856-2864)."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from oracle import quantize as oq

CASES = ["b22", "b21", "b22_small"]


@pytest.fixture(scope="module")
def pinned():
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "lasermix_pinned.npz"))


def _inputs(pinned, tag, dev):
    t = lambda k: torch.from_numpy(pinned[f"{tag}_{k}"]).to(dev)
    sup = {"coords": t("sup_coords"), "feats": t("sup_feats"), "mapped_labels": t("sup_labels")}
    unsup = {"coords": t("unsup_coords"), "feats": t("unsup_feats")}
    return sup, unsup, t("pseudo"), [int(a) for a in pinned[f"{tag}_areas"]]


@pytest.mark.parametrize("tag", CASES)
def test_mix_transform_equals_the_reference_on_cpu(pinned, tag):
    from gcdlss_b200.steps import mix_transform
    sup, unsup, pseudo, areas = _inputs(pinned, tag, "cpu")
    bc, f, l = mix_transform(sup, unsup, pseudo, areas)
    assert bc.dtype == torch.float32 and f.dtype == torch.float32 and l.dtype == torch.int32
    np.testing.assert_array_equal(bc.numpy(), pinned[f"{tag}_mix_bcoords"])
    np.testing.assert_array_equal(f.numpy(), pinned[f"{tag}_mix_feats"])
    np.testing.assert_array_equal(l.numpy(), pinned[f"{tag}_mix_labels"])
    # and the inline quantisation of the mixed batch (batch column divided by the voxel size too)
    qc, um, inv = oq.sparse_quantize_me(bc.numpy(), 0.05)
    np.testing.assert_array_equal(qc, pinned[f"{tag}_q_coords"])
    np.testing.assert_array_equal(um, pinned[f"{tag}_q_umap"])


def test_split_form_equals_the_reference_lists(pinned):
    # laser_mix(return_split=True) hands back the two scans the way laser_mix_transform does
    from gcdlss_b200.steps import laser_mix
    tag = "b21"
    sup, unsup, pseudo, areas = _inputs(pinned, tag, "cpu")
    s0, u0 = sup["coords"][:, 0] == 0, unsup["coords"][:, 0] == 0
    n_s, n_u = int(s0.sum()), int(u0.sum())
    (p1, f1, l1), (p2, f2, l2) = laser_mix(sup["coords"][s0][:, 1:], unsup["coords"][u0][:, 1:], sup["feats"][s0], unsup["feats"][u0],
                                           sup["mapped_labels"][:n_s], pseudo[:n_u], areas[0])
    ref = pinned[f"{tag}_mix_bcoords"]
    np.testing.assert_array_equal(p1.numpy(), ref[ref[:, 0] == 0][:, 1:])
    np.testing.assert_array_equal(p2.numpy(), ref[ref[:, 0] == 1][:, 1:])
    assert p1.shape[0] + p2.shape[0] == n_s + n_u


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_mix_transform_and_quantiser_on_the_gpu(cuda, pinned, tag):
    """The same on the device (the reference runs LaserMix on CUDA tensors too), followed by the GPU quantiser.  CUDA's
    atan2f may differ from the host's in the last bit, so a point whose pitch sits within 2 ulp of a band edge may change
    band; the fixture has none (asserted), hence equality."""
    from gcdlss_b200.quantize import sparse_quantize_gpu
    from gcdlss_b200.steps import mix_transform
    sup, unsup, pseudo, areas = _inputs(pinned, tag, cuda)
    bc, f, l = mix_transform(sup, unsup, pseudo, areas)
    np.testing.assert_array_equal(bc.cpu().numpy(), pinned[f"{tag}_mix_bcoords"])
    np.testing.assert_array_equal(f.cpu().numpy(), pinned[f"{tag}_mix_feats"])
    np.testing.assert_array_equal(l.cpu().numpy(), pinned[f"{tag}_mix_labels"])
    qc, um, inv = sparse_quantize_gpu(bc, 0.05)
    np.testing.assert_array_equal(qc.cpu().numpy(), pinned[f"{tag}_q_coords"])
    np.testing.assert_array_equal(um.cpu().numpy(), pinned[f"{tag}_q_umap"])
    np.testing.assert_array_equal(inv.cpu().numpy(), pinned[f"{tag}_q_inv"])
