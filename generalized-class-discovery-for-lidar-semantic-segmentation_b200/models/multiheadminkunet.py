"""Drop-in for the reference ``models/multiheadminkunet.py``: thin head wrappers around the
MinkUNet backbone (ref models/multiheadminkunet.py:9-629).  ``MinkUNetBase`` serves Stage 1
(``ExpPretrain``, ref modules/exp.py:78), ``MinkUNetRC`` serves the Stage-2 mean-teacher modules
(ref modules/exp_merge_mean_teacher.py:67-72), the ``MultiHeadMinkUnet*`` classes the NOPS-style
baselines.  Same constructors, attribute names and returned dict keys.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import MinkowskiEngine as ME

from models.minkunet import MinkUNet34C, MinkUNet34RC


def _stack_views(per_view_feats, per_view_out):
    merged = {"feats": torch.stack(per_view_feats)}
    for key in per_view_out[0].keys():
        merged[key] = torch.stack([o[key] for o in per_view_out])
    return merged


class Prototypes(nn.Module):
    """1x1 sparse conv classifier without bias; returns the dense logits."""

    def __init__(self, output_dim, num_prototypes, D=3):
        super().__init__()
        self.prototypes = ME.MinkowskiConvolution(output_dim, num_prototypes, kernel_size=1, bias=False, dimension=D)

    def forward(self, x):
        return self.prototypes(x).F


class ProjectionHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(96, 128), nn.ReLU(), nn.Linear(128, 256), nn.ReLU(), nn.Linear(256, 256))
        self.apply(self.init_weights)

    def init_weights(self, m):
        for sub in self.modules():
            if isinstance(sub, nn.Linear):
                nn.init.xavier_uniform_(sub.weight.data)
                if sub.bias is not None:
                    sub.bias.data.zero_()

    def forward(self, x):
        return self.mlp(x)


class CosinePrototypes(nn.Module):
    def __init__(self, output_dim, num_prototypes, D=3):
        super().__init__()
        self.prototypes = nn.Linear(output_dim, num_prototypes, bias=False)
        self.prototypes.weight.data.uniform_(-1, 1).renorm_(2, 1, 1e-5).mul_(1e5)

    def forward(self, x):
        return 10 * torch.mm(F.normalize(x.F, dim=-1), F.normalize(self.prototypes.weight, dim=-1).T)


class _MultiHeadBase(nn.Module):
    HEAD = None

    def __init__(self, input_dim, num_prototypes, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.prototypes = torch.nn.ModuleList([self.HEAD(input_dim, num_prototypes) for _ in range(num_heads)])

    def forward_head(self, head_idx, feats):
        return self.prototypes[head_idx](feats), feats.F

    def forward(self, feats):
        per_head = [self.forward_head(h, feats) for h in range(self.num_heads)]
        return [torch.stack(o) for o in map(list, zip(*per_head))]


class MultiHead(_MultiHeadBase):
    HEAD = Prototypes


class MultiHeadCosine(_MultiHeadBase):
    HEAD = CosinePrototypes


class _HeadedEncoder(nn.Module):
    """Backbone with ``final`` removed + heads; forward handles a tensor or a list of views."""

    def forward(self, views):
        if isinstance(views, list):
            feats = [self.encoder(view) for view in views]
            return _stack_views(feats, [self.forward_heads(f) for f in feats])
        feats = self.encoder(views)
        out = self.forward_heads(feats)
        out["feats"] = feats.F
        return out

    def _unlab_heads(self, feats, out):
        for name, key in (("head_unlab", ""), ("head_unlab_over", "_over")):
            if hasattr(self, name):
                logits, proj = getattr(self, name)(feats)
                out["logits_unlab" + key] = logits
                out["proj_feats_unlab" + key] = proj
        return out


class MultiHeadMinkUnetFineTune(_HeadedEncoder):
    def __init__(self, num_labeled, num_classes):
        super().__init__()
        self.encoder = MinkUNet34C(1, num_labeled)
        self.feat_dim = self.encoder.final.in_channels
        self.encoder.final = nn.Identity()
        self.head_lab = Prototypes(output_dim=self.feat_dim, num_prototypes=num_labeled)
        self.head_lab2 = nn.Linear(in_features=self.feat_dim, out_features=num_classes)

    def forward_heads(self, feats):
        return {"logits_lab": self.head_lab2(feats.F)}


class MultiHeadMinkUnet(_HeadedEncoder):
    def __init__(self, num_labeled, num_unlabeled, overcluster_factor=None, num_heads=1, in_channels=1):
        super().__init__()
        self.encoder = MinkUNet34C(1, num_labeled)
        self.feat_dim = self.encoder.final.in_channels
        self.encoder.final = nn.Identity()
        self.head_lab = Prototypes(output_dim=self.feat_dim, num_prototypes=num_labeled)
        if num_heads is not None:
            self.head_unlab = MultiHead(input_dim=self.feat_dim, num_prototypes=num_unlabeled, num_heads=num_heads)
        if overcluster_factor is not None:
            self.head_unlab_over = MultiHead(input_dim=self.feat_dim, num_prototypes=num_unlabeled * overcluster_factor,
                                             num_heads=num_heads)

    def forward_heads(self, feats):
        return self._unlab_heads(feats, {"logits_lab": self.head_lab(feats)})


class MultiHeadMinkUnetCosine(_HeadedEncoder):
    def __init__(self, num_labeled, num_unlabeled, overcluster_factor=None, num_heads=1, in_channels=1):
        super().__init__()
        self.encoder = MinkUNet34C(in_channels, num_labeled)
        self.feat_dim = self.encoder.final.in_channels
        self.encoder.final = nn.Identity()
        self.head_lab = CosinePrototypes(output_dim=self.feat_dim, num_prototypes=num_labeled)
        if num_heads is not None:
            self.head_unlab = MultiHeadCosine(input_dim=self.feat_dim, num_prototypes=num_unlabeled, num_heads=num_heads)
            if overcluster_factor is not None:
                self.head_unlab_over = MultiHeadCosine(input_dim=self.feat_dim, num_prototypes=num_unlabeled * overcluster_factor,
                                                       num_heads=num_heads)

    def forward_heads(self, feats):
        return self._unlab_heads(feats, {"logits_lab": self.head_lab(feats)})


class MinkUNetBase(nn.Module):
    """Stage-1 model (ref models/multiheadminkunet.py:309-340)."""

    def __init__(self, num_classes, in_channels=1):
        super().__init__()
        self.encoder = MinkUNet34RC(in_channels, num_classes)

    def forward(self, views):
        if isinstance(views, list):
            feats = [self.encoder.forward_no_logits(view) for view in views]
            return _stack_views(feats, [self.encoder.forward(view) for view in views])
        out = dict()
        if hasattr(self.encoder, 'final3'):
            feats = self.encoder.forward_no_logits(views)
            out['logits'] = self.encoder.forward_novel(views)
            out["feats"] = feats.F
            return out
        logits, feats = self.encoder.forward(views, use_last=True)
        out['logits'] = logits.F
        out["feats"] = feats.F
        return out


class MinkUNetRC(nn.Module):
    """Stage-2 teacher / student (ref models/multiheadminkunet.py:342-392); callers bolt ``final2`` /
    ``final3`` 1x1 heads onto ``self.encoder`` (ref modules/exp_merge_mean_teacher.py:128-153)."""

    def __init__(self, num_labeled, in_channels=1):
        super().__init__()
        self.encoder = MinkUNet34RC(in_channels, num_labeled)

    def forward(self, views):
        if isinstance(views, list):
            raise NotImplementedError
        feats = self.encoder.forward_no_logits(views)
        return {'logits': self.encoder.forward_dummy(feats), "feats": feats.F}

    def forward_discover(self, views):
        if isinstance(views, list):
            raise NotImplementedError
        feats = self.encoder.forward_no_logits(views)
        return {'logits': self.encoder.forward_novel(feats)}


class MinkUNetRCAblation(nn.Module):
    def __init__(self, num_labeled, in_channels=1, ncc_head_mean=False, ncc_head_sum=False):
        super().__init__()
        self.ncc_head_mean = ncc_head_mean
        self.ncc_head_sum = ncc_head_sum
        self.encoder = MinkUNet34RC(in_channels, num_labeled)

    def forward(self, views):
        if isinstance(views, list):
            raise NotImplementedError
        feats = self.encoder.forward_no_logits(views)
        if self.ncc_head_mean:
            return {'logits': self.encoder.forward_dummy_mean(feats), "feats": feats.F}
        if self.ncc_head_sum:
            return {'logits': self.encoder.forward_dummy_sum(feats), "feats": feats.F}
        return None

    def forward_discover(self, views):
        if isinstance(views, list):
            raise NotImplementedError
        return {'logits': self.encoder.forward_novel(self.encoder.forward_no_logits(views))}


class MinkUNetBaseCosine(nn.Module):
    def __init__(self, num_classes, in_channels=1):
        super().__init__()
        self.encoder = MinkUNet34RC(in_channels, num_classes)
        self.feat_dim = self.encoder.final.in_channels
        self.encoder.final = nn.Identity()
        self.head_lab = CosinePrototypes(output_dim=self.feat_dim, num_prototypes=num_classes)

    def forward_heads(self, feats):
        return {"logits": self.head_lab(feats)}

    def forward(self, views):
        if isinstance(views, list):
            feats = [self.encoder(view) for view in views]
            return _stack_views(feats, [self.forward_heads(f) for f in feats])
        feats = self.encoder.forward_no_logits(views)
        out = self.forward_heads(feats)
        out["feats"] = feats.F
        return out


class MinkUNetRCCosine(nn.Module):
    def __init__(self, num_labeled, in_channels=1):
        super().__init__()
        self.encoder = MinkUNet34RC(in_channels, num_labeled)
        self.feat_dim = self.encoder.final.in_channels
        self.encoder.final = nn.Identity()
        self.head_lab = CosinePrototypes(output_dim=self.feat_dim, num_prototypes=num_labeled)

    def forward(self, views):
        if isinstance(views, list):
            feats = [self.encoder.forward_no_logits(view) for view in views]
            return _stack_views(feats, [self.encoder.forward_dummy(view) for view in views])
        feats = self.encoder.forward_no_logits(views)
        known = self.head_lab(feats)
        rc = torch.max(self.head_ncc(feats), dim=1, keepdim=True)[0]   # ``head_ncc`` is attached by the caller
        return {'logits_ncc': torch.cat([known, rc], dim=1), "feats": feats.F}


class _SelfSupBase(nn.Module):
    def __init__(self, dataset='SemanticKITTI'):
        super().__init__()
        if dataset == 'nuScenes':
            in_channels = 1
        elif dataset == 'SemanticKITTI':
            in_channels = 4
        else:
            raise NotImplementedError
        self.backbone = MinkUNet34RC(in_channels, 128, D=3)
        self.metric_learner = ProjectionHead()


class MultiHeadSelfSupMinkUnet(_SelfSupBase):
    def forward(self, views):
        feats = self.backbone.forward_no_logits(views)
        return {"feats": feats.F, 'logits': self.backbone.final(feats).F}


class MultiHeadSelfSupMinkUnet2(_SelfSupBase):
    def __init__(self, dataset='SemanticKITTI', SimGCD=False):
        super().__init__(dataset)
        self.SimGCD = SimGCD

    def forward(self, views):
        feats = self.backbone.forward_no_logits(views)
        out = {'proj_feats': self.metric_learner(feats.F)}
        if self.SimGCD:
            normed = ME.SparseTensor(features=F.normalize(feats.F, dim=1), coordinates=feats.C)
            out['logits'] = self.backbone.final(normed).F
        else:
            out['logits'] = self.backbone.final(feats)
        return out


class MultiHeadSelfSupMinkUnetTest(_SelfSupBase):
    def forward(self, views):
        return {'feats': self.backbone.forward_no_logits(views).F}
