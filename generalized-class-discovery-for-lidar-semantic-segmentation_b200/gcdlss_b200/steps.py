"""Step harnesses: the tensor flow of the reference's two Lightning ``training_step``s around the hot path, without
Lightning (not installable here).  They exist so the path can be exercised and timed the way its callers use it:

* ``stage1_step``  — ``ExpPretrain.training_step`` (ref modules/exp.py:249-267): SparseTensor -> model -> CE.
* ``Stage2Harness.step`` — ``ExpMergeDiscover_LaserMix_MeanTeacher_NCCAdaptive.training_step``
  (ref modules/exp_merge_mean_teacher.py:2772-2875): teacher and student share one SparseTensor (and therefore all
  kernel maps), CE on labelled voxels + 200 * MSE(student, teacher) on unlabelled ones, teacher pseudo-labels gathered
  back to points, LaserMix by pitch-angle bands (ref :1731-1787), inline ``sparse_quantize`` of the mixed points
  (batch column divided by the voxel size too, as in the reference), second student forward, backward, SGD, EMA.
  The loss terms that are dense torch on logits (calibration, hinge, k-means/Hungarian) are outside the hot path and
  are not reproduced.  Everything stays on the GPU (the reference round-trips LaserMix through numpy).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .functional import devoxelize
from .quantize import sparse_quantize_gpu
from .sparse_tensor import SparseTensor


def point_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int | None = None) -> torch.Tensor:
    """``F.cross_entropy(logits, labels, ignore_index=...)`` (mean over the counted points) through the per-point
    kernels: torch's fused mean reduction is a single-block kernel (0.21 ms forward + 0.12 ms backward on 200 k points,
    more than two of the sparse convolutions), the unreduced form plus a mean is a few microseconds."""
    if ignore_index is None:
        return F.cross_entropy(logits, labels, reduction="none").mean()
    per_point = F.cross_entropy(logits, labels, reduction="none", ignore_index=ignore_index)
    return per_point.sum() / (labels != ignore_index).sum().clamp(min=1)


class _ConsistencyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_s, logits_t, threshold):
        from . import ops
        sq_err, max_prob, label, grad = ops.consistency_rows(logits_s.detach(), logits_t.detach(), threshold, want_grad=ctx.needs_input_grad[0])
        n, c = logits_s.shape
        ctx.save_for_backward(grad)
        ctx.numel, ctx.in_dtype = n * c, logits_s.dtype
        ctx.mark_non_differentiable(max_prob, label)
        return sq_err.sum() / (n * c), max_prob, label

    @staticmethod
    def backward(ctx, g_mse, _g_prob, _g_label):
        (grad,) = ctx.saved_tensors
        return (grad * (g_mse / ctx.numel)).to(ctx.in_dtype), None, None


def consistency_terms(logits_s: torch.Tensor, logits_t: torch.Tensor, threshold: float = 0.0):
    """(F.mse_loss(softmax(logits_s), softmax(logits_t)), max teacher probability [n], teacher argmax [n]) from one pass
    over the two logit matrices (gcd_consistency_rows; ref modules/exp_merge_mean_teacher.py:2832-2850).  Differentiable
    with respect to ``logits_s`` only (the reference detaches the teacher); labels are -1 where the teacher's maximum
    probability is below ``threshold`` (0 = keep all)."""
    return _ConsistencyFunction.apply(logits_s, logits_t, float(threshold))


def stage1_step(model, feats, bcoords, labels):
    st = SparseTensor(features=feats.float(), coordinates=bcoords.int())
    out = model(st)
    return point_cross_entropy(out["logits"], labels.long())


def _pitch_bands(points, lo: float, hi: float, edges):
    """Band index of every point: band i covers (angle_list[i + 1], angle_list[i]] of the clamped pitch angle, exactly the
    reference's comparisons (ref exp_merge_mean_teacher.py:1735-1765: float32 pitch against the float64 ``np.linspace`` edges,
    which torch compares in float32), counted instead of masked: band = number of interior edges at or above the pitch."""
    rho = torch.sqrt(points[:, 0] ** 2 + points[:, 1] ** 2)
    pitch = torch.clamp(torch.atan2(points[:, 2], rho), lo + 1e-5, hi - 1e-5)
    band = torch.zeros(points.shape[0], dtype=torch.int64, device=points.device)
    for e in edges[1:-1]:
        band += pitch <= float(e)
    return band


def laser_mix(points_sup, points_unsup, feats_sup, feats_unsup, labels_sup, labels_unsup, num_areas: int,
              pitch_angles=(-25, 3), return_split: bool = True):
    """Pitch-angle band swap between one labelled and one unlabelled scan, row for row what
    ``laser_mix_transform`` returns (ref exp_merge_mean_teacher.py:1731-1787): mix 1 is, band by band from the top,
    the labelled scan's points of the even bands and the unlabelled scan's points of the odd bands (each band's points in
    their original order); mix 2 is the complement.  Points are [P, 3] sensor-frame xyz.

    The reference builds 4 * num_areas boolean-mask selections (a device->host sync each); here every point gets the key
    ``mix * 8 + band`` and ONE stable sort of the keys yields both mixes, bands in order, original order inside a band (a
    band of a mix comes from one source only, so the sort never has to order the two sources against each other).
    ``return_split=False`` returns the sorted rows and their mix index without splitting (no host sync at all)."""
    import numpy as np
    lo, hi = pitch_angles[0] / 180 * np.pi, pitch_angles[1] / 180 * np.pi
    edges = np.linspace(hi, lo, num_areas + 1)
    bs, bu = _pitch_bands(points_sup, lo, hi, edges), _pitch_bands(points_unsup, lo, hi, edges)
    key = torch.cat([(bs % 2) * 8 + bs, (1 - bu % 2) * 8 + bu])             # labelled: even band -> mix 1 (index 0); unlabelled: odd band -> mix 1
    key, order = torch.sort(key, stable=True)
    pts = torch.cat([points_sup, points_unsup]).index_select(0, order)
    feats = torch.cat([feats_sup, feats_unsup]).index_select(0, order)
    labels = torch.cat([labels_sup, labels_unsup.to(labels_sup.dtype)]).index_select(0, order)
    mix = key >> 3
    if not return_split:
        return pts, feats, labels, mix
    n1 = int((mix == 0).sum())
    return (pts[:n1], feats[:n1], labels[:n1]), (pts[n1:], feats[n1:], labels[n1:])


def mix_transform(sup_data, unsup_data, mix_unsup_pseudo_labels, num_areas, pitch_angles=(-25, 3)):
    """Point assembly of the reference's ``mix_transform`` (ref exp_merge_mean_teacher.py:1577-1729) with the same
    arguments: ``sup_data`` / ``unsup_data`` are the collated point-level dicts ('coords' [P, 4] float (batch, x, y, z),
    'feats' [P, C], sup also 'mapped_labels' [P]); scan 0 of the labelled half is mixed with scan 0 of the unlabelled half
    and, when both halves hold the same number of scans, everything else of one half with everything else of the other
    (``coords[:, 0] != 0``: the reference is written for two scans per half).  ``num_areas``: the band counts the reference
    draws with ``np.random.choice(semi_train_cfg['num_areas'])``, one per LaserMix call, passed in so that the caller owns
    the randomness.  Returns (mix_bcoords float32 [Q, 4], mix_feats float32 [Q, C], mix_labels int32 [Q]) with batch
    indices 0, 1 (, 2, 3), row for row the reference's arrays, and everything stays on the device."""
    sc, uc = sup_data["coords"], unsup_data["coords"]
    first_s, first_u = sc[:, 0] == 0, uc[:, 0] == 0
    n_first_s, n_first_u = int(first_s.sum()), int(first_u.sum())
    labels, pseudo = sup_data["mapped_labels"], mix_unsup_pseudo_labels
    pairs = [(sc[first_s][:, 1:], uc[first_u][:, 1:], sup_data["feats"][first_s], unsup_data["feats"][first_u],
              labels[:n_first_s], pseudo[:n_first_u])]
    if len(torch.unique(sc[:, 0])) == len(torch.unique(uc[:, 0])):
        pairs.append((sc[~first_s][:, 1:], uc[~first_u][:, 1:], sup_data["feats"][~first_s], unsup_data["feats"][~first_u],
                      labels[n_first_s:], pseudo[n_first_u:]))
    areas = list(num_areas) if isinstance(num_areas, (list, tuple)) else [num_areas] * len(pairs)
    rows, feats, labs = [], [], []
    for i, (ps, pu, fs, fu, ls, lu) in enumerate(pairs):
        p, f, l, mix = laser_mix(ps, pu, fs, fu, ls, lu, int(areas[i]), pitch_angles, return_split=False)
        rows.append(torch.cat([(2 * i + mix).to(p.dtype)[:, None], p], 1))
        feats.append(f)
        labs.append(l)
    return torch.cat(rows).float(), torch.cat(feats).float(), torch.cat(labs).int()


def make_stage2_half(kind: str, idx0: int, n_scans: int, n_points, dev, with_labels: bool, n_classes: int = 17, host_scans=None):
    """One half (labelled or unlabelled) of a Stage-2 batch from synthetic scans ``idx0 .. idx0 + n_scans - 1``, laid out
    like the reference's collate output (ref utils/collation.py:430-467) and quantised on the GPU: voxel-level
    'coords' / 'feats' (/ 'labels'), point-level 'points' dict, 'inverse_maps' for the unlabelled half.
    ``host_scans``: optional list of (points [P, 4] pinned host tensor (x, y, z, remission), labels [P]) to start from host
    memory (bench.py's end-to-end loop); otherwise the scans are generated here."""
    import numpy as np

    from . import synth
    q = synth.voxel_size(kind)
    coords, feats, labels, pcoords, pfeats, plabels, invs = [], [], [], [], [], [], []
    for b in range(n_scans):
        if host_scans is not None:
            pts, lab = host_scans[b]
            pts, lab = pts.to(dev, non_blocking=True), lab.to(dev, non_blocking=True)
            p, ff = pts[:, :3].contiguous(), pts[:, 3:4].contiguous()
        else:
            xyz, f = synth.make_scan(kind, idx0 + b, n_points=n_points)
            p, ff = torch.from_numpy(xyz).to(dev), torch.from_numpy(f).to(dev)
            lab = torch.from_numpy(np.random.default_rng(idx0 + b).integers(0, n_classes, xyz.shape[0])).to(dev)
        c, um, inv = sparse_quantize_gpu(p, q)
        coords.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32, device=dev), c], 1))
        feats.append(ff[um])
        labels.append(lab[um])
        pcoords.append(torch.cat([torch.full((p.shape[0], 1), float(b), device=dev), p], 1))
        pfeats.append(ff)
        plabels.append(lab)
        invs.append(inv)
    d = {"coords": torch.cat(coords), "feats": torch.cat(feats), "n_scans": n_scans,
         "points": {"coords": torch.cat(pcoords), "feats": torch.cat(pfeats)}}
    if with_labels:
        d["labels"] = torch.cat(labels)
        d["points"]["mapped_labels"] = torch.cat(plabels)
    else:
        d["inverse_maps"] = invs
    return d


class Stage2Harness:
    def __init__(self, student, teacher, optimizer, voxel_size: float, mse_coeff: float = 200.0, ema_momentum: float = 0.01,
                 num_areas=(3, 4, 5, 6), reducer=None, fused_loss: bool = False):
        self.student, self.teacher, self.opt = student, teacher, optimizer
        self.voxel_size, self.mse_coeff, self.ema_momentum, self.num_areas = voxel_size, mse_coeff, ema_momentum, num_areas
        self.reducer = reducer
        if reducer is not None:
            reducer.direct = False       # two student passes per step accumulate into the same parameters (see GradBucketReducer.grad_ptr)
        self.fused_loss = fused_loss      # consistency terms through gcd_consistency_rows (one pass) instead of torch ops
        self._step = 0
        for p in self.teacher.parameters():
            p.requires_grad_(False)                                        # ref :155, :251-254

    def step(self, sup, unsup):
        """sup / unsup: what the Stage-2 collate hands the reference's training_step (ref :2774-2793), on the device:
        voxel level 'coords' [M, 4] int32, 'feats' [M, C]; point level 'points' = {'coords' [P, 4] float (batch, x, y, z),
        'feats' [P, C]}; sup also 'labels' [M] (voxels) and points['mapped_labels'] [P]; unsup also 'inverse_maps' (list of
        [P_i] int64, voxel of each point *within its scan*)."""
        n_sup_scans = int(sup["coords"][:, 0].max().item()) + 1 if "n_scans" not in sup else sup["n_scans"]
        unsup_coords = unsup["coords"].clone()
        unsup_coords[:, 0] += n_sup_scans                                  # ref :2797
        coords_cat = torch.cat((sup["coords"], unsup_coords), 0)
        feats_cat = torch.cat((sup["feats"], unsup["feats"]), 0)
        st = SparseTensor(features=feats_cat.float(), coordinates=coords_cat.int())
        with torch.no_grad():
            out_t = self.teacher(st)                                       # shares st's kernel maps with the student
        out_s = self.student(st)
        n_sup = sup["coords"].shape[0]
        loss = point_cross_entropy(out_s["logits"][:n_sup], sup["labels"].long())
        if self.fused_loss:
            mse, max_prob_t, target_t = consistency_terms(out_s["logits"][n_sup:], out_t["logits"][n_sup:])
            loss = loss + mse * self.mse_coeff
        else:
            prob_s = F.softmax(out_s["logits"][n_sup:], dim=1)
            prob_t = F.softmax(out_t["logits"][n_sup:], dim=1)
            loss = loss + F.mse_loss(prob_s, prob_t.detach()) * self.mse_coeff
            max_prob_t, target_t = torch.max(prob_t, dim=1)
        # voxel -> point devoxelisation of the pseudo labels.  The reference concatenates the scans' inverse maps without
        # offsetting the second by the first scan's voxel count (ref :2787-2790); preserved, not fixed.
        inv_cat = torch.cat(unsup["inverse_maps"])
        pts_prob = devoxelize(max_prob_t[:, None], inv_cat)[:, 0]
        pts_label = devoxelize(target_t[:, None].float(), inv_cat)[:, 0].long()
        pts_label[pts_prob < 0.9] = -1
        # LaserMix (ref :2854 -> mix_transform): scan 0 with scan 0, the rest with the rest; the band counts are this
        # step's draws (the reference draws them with np.random.choice inside laser_mix_transform)
        areas = [self.num_areas[(2 * self._step + j) % len(self.num_areas)] for j in range(2)]
        lm_points, lm_feats, lm_labels = mix_transform(sup["points"], unsup["points"], pts_label, areas)
        # inline quantisation of the mixed batch: the batch column is divided by the voxel size too (ref :2856-2861)
        lm_coords, lm_umap, _ = sparse_quantize_gpu(lm_points, self.voxel_size)
        mix_st = SparseTensor(features=lm_feats[lm_umap].float(), coordinates=lm_coords)      # the first point of a voxel represents it (ref :2864)
        out_mix = self.student(mix_st)
        mix_labels = lm_labels[lm_umap]
        if bool((mix_labels >= 0).any()):
            loss = loss + 0.1 * point_cross_entropy(out_mix["logits"], mix_labels.long(), ignore_index=-1)
        else:
            loss = loss + 0.0 * out_mix["logits"].sum()
        if self.reducer is not None:
            self.reducer.reset()
        else:
            self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        with torch.no_grad():                                              # EMA teacher <- student, parameters only (ref :246-248)
            ps, pt = list(self.student.parameters()), list(self.teacher.parameters())
            torch._foreach_mul_(pt, 1.0 - self.ema_momentum)
            torch._foreach_add_(pt, ps, alpha=self.ema_momentum)
        self._step += 1
        return loss.detach()
