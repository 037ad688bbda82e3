"""GPU-side counterparts of what the reference's Dataset / collate functions do per scan on CPU cores:

* ``prepare_scan``: down-sampled points -> float64 rigid / voxelisation transform -> ``sparse_quantize`` -> unique features /
  labels (ref utils/dataset_remission.py:813-880), everything on the device, maps left there;
* ``collate``: batch index column + concatenation, with the keys and dtypes of ``collation_fn_restricted_dataset``
  (ref utils/collation.py:29-42); ``collate_lasermix`` those of ``collation_fn_lasermix_dataset`` (ref :430-467).

The random choices (down-sampling indices, augmentation matrices) stay with the caller, as in the reference's Dataset: they
are host-side numpy draws and arrive here as arguments."""
from __future__ import annotations

import torch

from . import ops
from ._cabi import ROUND_FLOOR


def prepare_scan(points: torch.Tensor, remission: torch.Tensor, labels, mapped_labels, rigid_transformation, voxel_size: float,
                 selected_idx=None):
    """points [P, 3] float32 CUDA, remission [P] / [P, 1], labels / mapped_labels [P] (or None), rigid_transformation 4x4 float64
    (``affine_mtx @ voxel_mtx``; None = no augmentation), selected_idx optional [S] int64 (sorted down-sampling indices).
    Returns the tuple ``__getitem__`` returns (ref :882-890): (coords int32 [M, 3], feats [M, 1], labels [M], selected_idx [M],
    mapped_labels [M], inverse_map [S]) -- on the device."""
    dev = points.device
    if selected_idx is None:
        selected_idx = torch.arange(points.shape[0], device=dev)
    else:
        selected_idx = torch.as_tensor(selected_idx, device=dev, dtype=torch.int64)
        points = points.index_select(0, selected_idx)
        remission = remission.index_select(0, selected_idx)
    feats = remission.reshape(-1, 1)
    if rigid_transformation is not None:
        coordinates = ops.affine_f64(points, rigid_transformation)              # float64, as numpy's float32 @ float64
    else:
        coordinates = points
    icoords = ops.quantize(coordinates, float(voxel_size), 3, ROUND_FLOOR)
    unique_map, inverse_map, _ = ops.unique_rows(icoords, order=0)
    sel = selected_idx.index_select(0, unique_map)
    pick = (lambda t: None) if labels is None else (lambda t: torch.as_tensor(t, device=dev).index_select(0, sel))
    return (icoords.index_select(0, unique_map), feats.index_select(0, unique_map), pick(labels), sel,
            pick(mapped_labels if mapped_labels is not None else labels), inverse_map)


def _batched(coords_list, dtype=torch.int32):
    rows = [torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device), c.to(dtype)], 1) for b, c in enumerate(coords_list)]
    return torch.cat(rows, 0)


def collate(samples, pcd_indexes=None):
    """samples: list of ``prepare_scan`` tuples.  Returns (bcoords int32 [M, 4], feats float32, labels int32, selected_idx int64,
    mapped_labels int32, inverse_maps (tuple), pcd_indexes int16) like ``collation_fn_restricted_dataset``."""
    coords, feats, labels, selected_idx, mapped_labels, inverse_maps = list(zip(*samples))
    idx = torch.tensor(list(range(len(samples))) if pcd_indexes is None else list(pcd_indexes), dtype=torch.int16)
    cat_i32 = lambda ts: None if ts[0] is None else torch.cat(ts, 0).int()
    return (_batched(coords), torch.cat(feats, 0).float(), cat_i32(labels), torch.cat(selected_idx, 0).long(), cat_i32(mapped_labels),
            inverse_maps, idx)


def collate_lasermix(point_samples, voxel_samples, pcd_indexes=None):
    """``collation_fn_lasermix_dataset``: point-level (coords float32 [P, 3], feats, labels, selected_idx, mapped_labels) and
    voxel-level ``prepare_scan`` tuples per scan -> {'points': {...}, 'voxel': {...}} with the reference's keys."""
    pc, pf, pl, ps, pm = list(zip(*point_samples))
    bcoords, feats, labels, sel, mapped, inverse_maps, idx = collate(voxel_samples, pcd_indexes)
    return {"points": {"coords": _batched(pc, torch.float32), "feats": torch.cat(pf, 0).float(), "labels": torch.cat(pl, 0).int(),
                       "selected_idx": torch.cat(ps, 0).long(), "mapped_labels": torch.cat(pm, 0).int()},
            "voxel": {"coords": bcoords, "feats": feats, "labels": labels, "selected_idx": sel, "mapped_labels": mapped,
                      "pcd_indexes": idx, "inverse_maps": inverse_maps}}
