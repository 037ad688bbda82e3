"""Oracle: point -> voxel quantisation (CPU, numpy).  Test infrastructure only.

Two different quantisers exist in the reference and they give different outputs:

A. ``Voxelizer.voxelize`` 'minkunet' branch, reference ``models/voxelizer.py:271-302`` with
   ``ravel_hash`` (``:312-332``) and ``sparse_quantize`` (``:334-360``):
   round-half-even(xyz / voxel) in fp32, shift by the per-scan minimum, np.unique on a
   x-major Horner hash => voxels ordered by ascending hash, ``indices`` = first occurrence,
   ``inverse`` = position in the sorted order.

B. ``ME.utils.sparse_quantize(coordinates, return_index=True, return_inverse=True,
   quantization_size=q)`` as called from ``utils/dataset_remission.py:868-873`` and
   ``modules/exp_merge_mean_teacher.py:2856-2861`` [ME-upstream semantics, SURVEY 8(a) a6]:
   floor(coords / q) in the *input dtype*, -> int32, unique rows in first-occurrence order:
   ``unique_map`` ascending, ``inverse_map[i]`` = voxel of point i.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- flavour A
def ravel_hash(x: np.ndarray) -> np.ndarray:
    """x-major Horner hash of non-negative-shifted integer coords (ref models/voxelizer.py:312-332)."""
    assert x.ndim == 2, x.shape
    x = x - np.min(x, axis=0)
    x = x.astype(np.uint64, copy=False)
    extent = np.max(x, axis=0).astype(np.uint64) + np.uint64(1)
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for d in range(x.shape[1] - 1):
        h += x[:, d]
        h *= extent[d + 1]
    h += x[:, -1]
    return h


def sparse_quantize_np_unique(coords: np.ndarray):
    """np.unique flavour (ref models/voxelizer.py:334-360) -> (indices [M], inverse [N]), int64."""
    _, indices, inverse = np.unique(ravel_hash(coords), return_index=True, return_inverse=True)
    return indices.astype(np.int64), inverse.reshape(-1).astype(np.int64)


def round_half_even_div_f32(xyz: np.ndarray, voxel_size) -> np.ndarray:
    """torch.round(res[:, :3] / voxel_size).int() in fp32 (ref models/voxelizer.py:275).

    numpy's rint is round-half-even like torch.round; the division is an IEEE fp32 division.
    """
    v = np.asarray(voxel_size, dtype=np.float32)
    return np.rint(xyz.astype(np.float32, copy=False) / v).astype(np.int32)


def voxelize_minkunet(points_list, voxel_size, batch_first: bool = True):
    """Whole 'minkunet' branch (ref models/voxelizer.py:271-302), max_voxels cap not applied.

    points_list: list of [N_i, 3+C] float32.  Returns dict with 'voxels' [sum M, 3+C] f32,
    'coors' [sum M, 4] int32, 'point2voxel_maps' list of [N_i] int64, 'voxel_inds' list of [M_i] int64.
    """
    voxels, coors, p2v, vinds = [], [], [], []
    for b, res in enumerate(points_list):
        res = np.asarray(res, dtype=np.float32)
        c = round_half_even_div_f32(res[:, :3], voxel_size)
        c = c - c.min(axis=0)
        inds, inverse = sparse_quantize_np_unique(c)
        vc = c[inds]
        bcol = np.full((vc.shape[0], 1), b, dtype=np.int32)
        vc = np.concatenate([bcol, vc], 1) if batch_first else np.concatenate([vc, bcol], 1)
        voxels.append(res[inds])
        coors.append(vc.astype(np.int32))
        p2v.append(inverse)
        vinds.append(inds)
    return {
        "voxels": np.concatenate(voxels, 0),
        "coors": np.concatenate(coors, 0),
        "point2voxel_maps": p2v,
        "voxel_inds": vinds,
    }


# --------------------------------------------------------------------------- flavour B
def floor_div(coords: np.ndarray, q) -> np.ndarray:
    """floor(coords / q) in the input dtype, then int32 [ME-upstream, SURVEY 8(a) a6].

    The division is an IEEE division by ``q`` cast to the input dtype (never a multiply by 1/q).
    Integer inputs are divided in float64 like numpy true division does.
    """
    c = np.asarray(coords)
    if c.dtype == np.float32:
        d = np.floor(c / np.float32(q))
    else:
        d = np.floor(c.astype(np.float64, copy=False) / np.float64(q))
    return d.astype(np.int32)


def unique_first_occurrence(icoords: np.ndarray):
    """Unique rows in first-occurrence order.

    Returns (unique_map [M] int64 ascending, inverse_map [N] int64) with
    icoords[unique_map][inverse_map] == icoords.
    """
    n = icoords.shape[0]
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    _, first, inv = np.unique(icoords, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")          # sorted-unique id -> rank by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return first[order].astype(np.int64), rank[inv].astype(np.int64)


def sparse_quantize_me(coordinates: np.ndarray, quantization_size=None):
    """ME.utils.sparse_quantize(return_index=True, return_inverse=True) restatement.

    Returns (discrete_coords[unique_map] int32 [M,D], unique_map [M] int64, inverse_map [N] int64).
    """
    c = np.asarray(coordinates)
    d = floor_div(c, quantization_size) if quantization_size is not None else np.floor(c).astype(np.int32)
    umap, inv = unique_first_occurrence(d)
    return d[umap], umap, inv


def batched_coordinates(coords_list, dtype=np.int32) -> np.ndarray:
    """ME.utils.batched_coordinates: prepend the batch index column."""
    out = []
    for b, c in enumerate(coords_list):
        c = np.asarray(c)
        if dtype == np.int32:
            c = np.floor(c).astype(np.int32)
        col = np.full((c.shape[0], 1), b, dtype=dtype)
        out.append(np.concatenate([col, c.astype(dtype)], 1))
    D = np.asarray(coords_list[0]).shape[1] if len(coords_list) else 3
    return np.concatenate(out, 0) if out else np.zeros((0, 1 + D), dtype)


def dataset_transform(points: np.ndarray, rigid_transformation: np.ndarray) -> np.ndarray:
    """The augmentation transform of the reference's Dataset, verbatim (ref utils/dataset_remission.py:824-833): float32
    points, homogeneous ones of the same dtype, a float64 4x4 matrix -> float64 [N, 3]."""
    coordinates = points
    homo_coords = np.hstack((coordinates, np.ones((coordinates.shape[0], 1), dtype=coordinates.dtype)))
    return homo_coords @ rigid_transformation.T[:, :3]
