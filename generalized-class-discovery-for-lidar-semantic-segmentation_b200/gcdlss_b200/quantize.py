"""Point -> voxel quantisation on the GPU, behind the two interfaces the reference uses.

* ``sparse_quantize`` / ``batched_coordinates``: the ``ME.utils`` functions called by every
  dataset ``__getitem__`` (ref utils/dataset_remission.py:868-873), the collate functions
  (ref utils/collation.py:33) and inline by the Stage-2 step on the LaserMix batch
  (ref modules/exp_merge_mean_teacher.py:2856-2861).  floor(coords / q) in the input dtype,
  voxels in first-occurrence order.
* ``voxelize_minkunet``: the 'minkunet' branch of ``Voxelizer.voxelize``
  (ref models/voxelizer.py:271-302): round-half-even, min shift, voxels in ascending key order.

Inputs may live on the host (numpy / CPU tensors); they are copied to the GPU, quantised and
deduplicated there by the hash kernels, and the maps are returned as CPU int64 tensors, which is
what the reference's callers index with.

DataLoader workers.  The reference calls ``ME.utils.sparse_quantize`` from ``Dataset.__getitem__`` inside
DataLoader workers (ref modules/exp.py:176-202, ``num_workers=8``).  Workers are *forked* by default and
CUDA cannot be initialised in a forked child of a process that already holds a context; there is no CPU
implementation to fall back to (by design).  Two supported arrangements:

* ``DataLoader(..., multiprocessing_context="spawn")``: every worker owns a CUDA context and quantises on the
  GPU; the call works unchanged (tests/test_quantize_workers.py);
* quantise after the loader, on the training process's GPU: ``sparse_quantize_gpu`` on a side stream through
  ``prefetch.BatchPrefetcher`` (what bench.py's e2e loop does) -- the faster of the two.

A call from a forked worker raises a RuntimeError that says so instead of torch's "Cannot re-initialize CUDA in
forked subprocess".
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops
from ._cabi import ROUND_FLOOR, ROUND_HALF_EVEN


_IMPORT_PID = os.getpid()


_forked = [False]
if hasattr(os, "register_at_fork"):
    os.register_at_fork(after_in_child=lambda: _forked.__setitem__(0, True))


def _device(device=None):
    if _forked[0] and getattr(torch.cuda, "_is_in_bad_fork", lambda: False)():
        raise RuntimeError(
            "gcdlss_b200.sparse_quantize was called in a fork()ed worker of a process that already initialised CUDA; CUDA cannot be "
            "used there and there is no CPU fallback.  Use DataLoader(..., multiprocessing_context='spawn'), or return raw points "
            "from the dataset and quantise on the training GPU (gcdlss_b200.quantize.sparse_quantize_gpu / prefetch.BatchPrefetcher).")
    if device is not None and str(device) != "cpu":
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("gcdlss_b200 quantisation runs on CUDA (sm_100a) only; no GPU is visible and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _scalar_q(q):
    if q is None:
        return 1.0
    if isinstance(q, (list, tuple, np.ndarray, torch.Tensor)):
        vals = [float(v) for v in (q.tolist() if hasattr(q, "tolist") else q)]
        if len(set(vals)) != 1:
            raise NotImplementedError("per-axis quantisation sizes are not used by the reference")
        return vals[0]
    return float(q)


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False, return_inverse=False,
                    return_maps_only=False, quantization_size=None, device="cpu"):
    """Drop-in for ``ME.utils.sparse_quantize`` (labels voting is not on the reference's path)."""
    if labels is not None:
        raise NotImplementedError("sparse_quantize(labels=...) is not used by the reference's hot path")
    is_np = isinstance(coordinates, np.ndarray)
    c = torch.from_numpy(np.ascontiguousarray(coordinates)) if is_np else coordinates
    if c.dim() != 2 or c.shape[1] not in (3, 4):
        raise ValueError("coordinates must be [N, 3] or [N, 4]")
    src_device = c.device
    if not c.is_floating_point():
        c = c.to(torch.float64)      # numpy true division of integers is float64
    elif c.dtype not in (torch.float32, torch.float64):
        c = c.to(torch.float32)
    dev = c.device if c.is_cuda else _device()
    cg = c.to(dev, non_blocking=True)
    dims = cg.shape[1]
    icoords = ops.quantize(cg, _scalar_q(quantization_size), dims, ROUND_FLOOR)
    unique_idx, inverse, _ = ops.unique_rows(icoords, order=0)
    if return_maps_only:
        um, im = unique_idx.cpu(), inverse.cpu()
        return (um, im) if return_inverse else um
    vox = icoords.index_select(0, unique_idx)
    um, im = unique_idx.cpu(), inverse.cpu()
    if is_np:
        vox_out = vox.cpu().numpy()
    else:
        vox_out = vox.to(src_device)
    out = [vox_out]
    if features is not None:
        out.append(features[um.numpy() if isinstance(features, np.ndarray) else um.to(features.device)])
    if return_index:
        out.append(um)
    if return_inverse:
        out.append(im)
    return out[0] if len(out) == 1 else tuple(out)


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """Drop-in for ``ME.utils.batched_coordinates``: prepend the batch index column (host glue)."""
    if dtype not in (torch.int32, torch.float32):
        raise ValueError("dtype must be torch.int32 or torch.float32")
    rows = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c)
        if dtype == torch.int32 and c.is_floating_point():
            c = torch.floor(c)
        c = c.to(dtype)
        rows.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device), c], 1))
    d = torch.as_tensor(coords[0]).shape[1] if len(coords) else 3
    out = torch.cat(rows, 0) if rows else torch.zeros((0, d + 1), dtype=dtype)
    return out.to(device) if device is not None else out


def voxelize_minkunet(points, voxel_size, batch_first: bool = True, max_voxels=None, training: bool = False):
    """The 'minkunet' voxel type of ``Voxelizer.voxelize`` (ref models/voxelizer.py:271-302).

    points: list of [N_i, 3+C] float32 tensors.  Returns the reference's ``voxel_dict``:
    'voxels' [sum M, 3+C], 'coors' [sum M, 4] int32, 'point2voxel_maps' (list of [N_i] int64),
    'voxel_inds' (list of [M_i] int64), everything on the GPU.
    """
    dev = _device(points[0].device if points and points[0].is_cuda else None)
    vs = voxel_size if not isinstance(voxel_size, (list, tuple)) else _scalar_q(voxel_size)
    voxels, coors, p2v, vinds = [], [], [], []
    for b, res in enumerate(points):
        res = res.to(dev, torch.float32)
        ic = ops.shift_to_min(ops.quantize(res, float(vs), 3, ROUND_HALF_EVEN))
        inds, inverse, _ = ops.unique_rows(ic, order=1)
        if training and max_voxels is not None and inds.shape[0] > max_voxels:
            keep = torch.randperm(inds.shape[0], device=dev)[:max_voxels]   # ref: np.random.choice without replacement
            inds = inds[keep]
        vc = ic.index_select(0, inds)
        bcol = torch.full((vc.shape[0], 1), b, dtype=torch.int32, device=dev)
        coors.append(torch.cat([bcol, vc], 1) if batch_first else torch.cat([vc, bcol], 1))
        voxels.append(res.index_select(0, inds))
        p2v.append(inverse)
        vinds.append(inds)
    return {"voxels": torch.cat(voxels, 0), "coors": torch.cat(coors, 0), "point2voxel_maps": p2v, "voxel_inds": vinds}


def sparse_quantize_gpu(points: torch.Tensor, quantization_size: float):
    """GPU-resident variant: CUDA points [N, 3|4] in, (voxel coords int32 [M, D], unique_map [M], inverse_map [N])
    out, everything left on the device (no host round trip of the maps)."""
    ic = ops.quantize(points, float(quantization_size), points.shape[1], ROUND_FLOOR)
    unique_idx, inverse, _ = ops.unique_rows(ic, order=0)
    return ic.index_select(0, unique_idx), unique_idx, inverse
