"""Drop-in surface: class names, constructors, state_dict keys/shapes (CPU only, no kernels run)."""
import importlib.util
import os
import sys

import pytest
import torch

import _paths  # noqa: F401
import MinkowskiEngine as ME
from models import minkunet as mu
from models import multiheadminkunet as mh

REF = "/root/reference"


def test_minkunet34c_parameter_inventory():
    torch.manual_seed(0)
    m = mu.MinkUNet34C(1, 17)
    sd = m.state_dict()
    assert sd["conv0p1s1.kernel"].shape == (125, 1, 32)
    assert sd["block1.0.conv1.kernel"].shape == (27, 32, 32)
    assert sd["block2.0.downsample.0.kernel"].shape == (32, 64)          # 2-D kernel when K^3 == 1
    assert sd["block5.0.conv1.kernel"].shape == (27, 384, 256)
    assert sd["convtr4p16s2.kernel"].shape == (8, 256, 256)
    assert sd["final.kernel"].shape == (96, 17) and sd["final.bias"].shape == (1, 17)
    assert sd["bn0.bn.running_mean"].shape == (32,) and "bn0.bn.num_batches_tracked" in sd
    convs = [mod for mod in m.modules() if isinstance(mod, (ME.MinkowskiConvolution, ME.MinkowskiConvolutionTranspose))]
    bns = [mod for mod in m.modules() if isinstance(mod, ME.MinkowskiBatchNorm)]
    assert len(convs) == 63 and len(bns) == 62                             # SURVEY 8(a) layer list
    n_conv_weights = sum(c.kernel.numel() for c in convs)
    assert n_conv_weights == 37_830_144
    assert not hasattr(m, "dropout") and hasattr(mu.MinkUNet34RC(1, 17), "dropout")


def test_weight_init_follows_reference():
    torch.manual_seed(0)
    m = mu.MinkUNet34C(1, 17)
    k = m.block3[0].conv1.kernel                                           # kaiming fan_out: std = sqrt(2 / (Cout * 27))
    assert abs(k.std().item() - (2.0 / (128 * 27)) ** 0.5) < 2e-4
    t = m.convtr5p8s2.kernel                                               # transposed convs keep ME's uniform init
    bound = 1.0 / (128 * 8) ** 0.5
    assert t.abs().max().item() <= bound + 1e-7 and t.abs().max().item() > 0.9 * bound
    assert torch.all(m.bn3.bn.weight == 1) and torch.all(m.bn3.bn.bias == 0)


def test_variants_and_wrappers_construct():
    for name in ["MinkUNet14A", "MinkUNet18A", "MinkUNet34A", "MinkUNet34B", "MinkUNet34C", "MinkUNet34RC", "MinkUNet50"]:
        getattr(mu, name)(1, 5)
    assert mu.MinkUNet50(1, 5).final.in_channels == 96 * 4
    s1 = mh.MinkUNetBase(num_classes=17)
    assert "encoder.final.kernel" in s1.state_dict()
    s2 = mh.MinkUNetRC(num_labeled=17)
    s2.encoder.final2 = ME.MinkowskiConvolution(96, 3, kernel_size=1, bias=True, dimension=3)   # as exp_merge_mean_teacher.py:128-153
    assert s2.state_dict()["encoder.final2.bias"].shape == (1, 3)
    nops = mh.MultiHeadMinkUnet(num_labeled=17, num_unlabeled=2, overcluster_factor=3, num_heads=2)
    assert isinstance(nops.encoder.final, torch.nn.Identity)
    assert nops.state_dict()["head_unlab_over.prototypes.1.prototypes.kernel"].shape == (96, 6)


def test_cpu_tensors_are_refused_loudly():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ME.SparseTensor(features=torch.zeros(3, 1), coordinates=torch.zeros(3, 4, dtype=torch.int32))


def _load_reference_module(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_state_dict_identical_to_unmodified_reference_models():
    """The reference's own models/*.py import against the MinkowskiEngine shim and yield the same keys/shapes."""
    saved = {k: sys.modules.get(k) for k in ("models", "models.resnet", "models.minkunet", "models.multiheadminkunet")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        pkg = type(sys)("models")
        pkg.__path__ = [os.path.join(REF, "models")]
        sys.modules["models"] = pkg
        _load_reference_module("models.resnet", "models/resnet.py")
        ref_mu = _load_reference_module("models.minkunet", "models/minkunet.py")
        ref_mh = _load_reference_module("models.multiheadminkunet", "models/multiheadminkunet.py")
        pairs = [(ref_mu.MinkUNet34C(1, 17), mu.MinkUNet34C(1, 17)), (ref_mu.MinkUNet34RC(1, 17), mu.MinkUNet34RC(1, 17)),
                 (ref_mu.MinkUNet50(1, 17), mu.MinkUNet50(1, 17)), (ref_mh.MinkUNetRC(17), mh.MinkUNetRC(17)),
                 (ref_mh.MultiHeadMinkUnet(17, 2, 3, 2), mh.MultiHeadMinkUnet(17, 2, 3, 2))]
        for ref_model, ours in pairs:
            a, b = ref_model.state_dict(), ours.state_dict()
            assert list(a.keys()) == list(b.keys())
            assert all(a[k].shape == b[k].shape for k in a)
            ours.load_state_dict(a, strict=True)
    finally:
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)


def test_point_cross_entropy_matches_torch():
    """steps.point_cross_entropy == F.cross_entropy (mean over counted points), with and without ignore_index, incl. grads."""
    import torch
    import torch.nn.functional as F
    from gcdlss_b200.steps import point_cross_entropy
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(500, 17, generator=g, dtype=torch.float64, requires_grad=True)
    labels = torch.randint(0, 17, (500,), generator=g)
    a = point_cross_entropy(logits, labels)
    b = F.cross_entropy(logits, labels)
    assert torch.allclose(a, b, rtol=1e-12, atol=0)
    ga, = torch.autograd.grad(a, logits)
    gb, = torch.autograd.grad(b, logits)
    assert torch.allclose(ga, gb, rtol=1e-10, atol=1e-15)
    labels_ign = labels.clone()
    labels_ign[::3] = -1
    a = point_cross_entropy(logits, labels_ign, ignore_index=-1)
    b = F.cross_entropy(logits, labels_ign, ignore_index=-1)
    assert torch.allclose(a, b, rtol=1e-12, atol=0)
