"""Drop-in for the reference ``models/voxelizer.py``: batch voxelisation front-end of the
mmdet3d-style path (ref models/voxelizer.py:25-487), minus the mmdet/mmengine base classes, which
are not installable here.  Same class names, constructor arguments and returned ``voxel_dict``.

Voxel types: 'minkunet' runs on the sm_100a quantise + hash-dedup kernels; 'dynamic' and
'cylindrical' are per-point coordinate formulas (torch elementwise, as in the reference's own
cylindrical branch); 'hard' (fixed points-per-voxel buffers, mmcv ``hard_voxelize``) dedups and groups with the
hash / counting-sort kernels.
"""
from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from gcdlss_b200 import quantize as _q


class VoxelLayer(nn.Module):
    """Holds the voxelisation geometry (ref models/voxelizer.py:362-407)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points: int, max_voxels: Union[tuple, int] = 20000,
                 deterministic: bool = True, grid_shape=None):
        super().__init__()
        if voxel_size and grid_shape:
            raise ValueError('voxel_size and grid_shape cannot be setting at the same time')
        self.point_cloud_range = point_cloud_range
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        if voxel_size:
            self.voxel_size = voxel_size
            grid = torch.round((pcr[3:] - pcr[:3]) / torch.tensor(voxel_size, dtype=torch.float32)).long().tolist()
            self.grid_shape = grid
        elif grid_shape:
            self.grid_shape = grid_shape
            self.voxel_size = ((pcr[3:] - pcr[:3]) / (torch.tensor(grid_shape, dtype=torch.float32) - 1)).tolist()
        else:
            raise ValueError('must assign a value to voxel_size or grid_shape')
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)
        self.deterministic = deterministic

    def forward(self, input: Tensor) -> Tensor:
        """Dynamic voxelisation (max_num_points == -1): per-point (z, y, x) grid coordinates, -1 outside
        the range (mmcv ``dynamic_voxelize`` semantics, ref models/voxelizer.py:453-461)."""
        if self.max_num_points != -1 and self.max_voxels[0] != -1:
            return self.hard_voxelize(input, self.max_voxels[0] if self.training else self.max_voxels[1])
        lo = input.new_tensor(self.point_cloud_range[:3])
        vs = input.new_tensor(self.voxel_size)
        grid = torch.tensor(self.grid_shape, device=input.device)
        c = torch.floor((input[:, :3] - lo) / vs).int()
        bad = ((c < 0) | (c >= grid)).any(1)
        c = c[:, [2, 1, 0]]
        c[bad] = -1
        return c


def _hard_voxelize(self, points: Tensor, max_voxels: int):
    """mmcv ``hard_voxelize`` (deterministic flavour, ref models/voxelizer.py:463-485): voxels in first-occurrence
    order of their points, the first ``max_num_points`` points of each voxel in point order, at most ``max_voxels``
    voxels.  Returns (voxels [M, max_points, C] zero padded, coors [M, 3] as (z, y, x), num_points [M]).
    Dedup and grouping run on the hash / counting-sort kernels; the rest is index arithmetic."""
    from gcdlss_b200 import ops
    lo = points.new_tensor(self.point_cloud_range[:3])
    vs = points.new_tensor(self.voxel_size)
    grid = torch.tensor(self.grid_shape, device=points.device)
    c = torch.floor((points[:, :3] - lo) / vs).int()
    keep = torch.nonzero(((c >= 0) & (c < grid)).all(1)).reshape(-1)
    c, pts = c.index_select(0, keep).contiguous(), points.index_select(0, keep)
    uniq, inv, _ = ops.unique_rows(c, order=0)
    m = min(int(uniq.shape[0]), int(max_voxels))
    seg_off, order = ops.csr_build(inv, uniq.shape[0])
    vox_of_sorted = inv.index_select(0, order.long())
    rank = torch.arange(order.shape[0], device=points.device) - seg_off.long().index_select(0, vox_of_sorted)
    sel = (rank < self.max_num_points) & (vox_of_sorted < m)
    voxels = points.new_zeros((m, self.max_num_points, points.shape[1]))
    voxels[vox_of_sorted[sel], rank[sel]] = pts.index_select(0, order.long()[sel])
    num_points = torch.clamp(seg_off[1:m + 1] - seg_off[:m], max=self.max_num_points).int()
    coors = c.index_select(0, uniq[:m])[:, [2, 1, 0]].contiguous()
    return voxels, coors, num_points


VoxelLayer.hard_voxelize = _hard_voxelize


class Voxelizer(nn.Module):
    """Point-cloud pre-processor: voxelises ``data['inputs']['points']`` (ref models/voxelizer.py:25-126)."""

    def __init__(self, voxel: bool = False, voxel_type: str = 'hard', voxel_layer: Optional[dict] = None, batch_first: bool = True,
                 max_voxels: Optional[int] = None, mean: Sequence = None, std: Sequence = None, pad_size_divisor: int = 1,
                 pad_value=0, pad_mask: bool = False, mask_pad_value: int = 0, pad_seg: bool = False, seg_pad_value: int = 255,
                 bgr_to_rgb: bool = False, rgb_to_bgr: bool = False, boxtype2tensor: bool = True, non_blocking: bool = False,
                 batch_augments: Optional[List[dict]] = None) -> None:
        super().__init__()
        self.voxel = voxel
        self.voxel_type = voxel_type
        self.batch_first = batch_first
        self.max_voxels = max_voxels
        self.non_blocking = non_blocking
        if voxel:
            self.voxel_layer = VoxelLayer(**voxel_layer)

    # -- data plumbing (the mmengine BaseDataPreprocessor part, reduced to what the path needs) --
    def cast_data(self, data):
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if isinstance(data, dict):
            return {k: self.cast_data(v) for k, v in data.items()}
        if isinstance(data, (list, tuple)):
            return type(data)(self.cast_data(v) for v in data)
        if isinstance(data, Tensor) and dev is not None:
            return data.to(dev, non_blocking=self.non_blocking)
        return data

    def collate_data(self, data: dict) -> dict:
        data = self.cast_data(data)
        data.setdefault('data_samples', None)
        return data

    def forward(self, data: Union[dict, List[dict]], training: bool = False):
        if isinstance(data, list):
            return [self.simple_process(d, training) for d in data]
        return self.simple_process(data, training)

    def simple_process(self, data: dict, training: bool = False) -> dict:
        data = self.collate_data(data)
        inputs = data['inputs'] if 'inputs' in data else data
        batch_inputs = dict()
        if 'points' in inputs:
            batch_inputs['points'] = inputs['points']
            if self.voxel:
                batch_inputs['voxels'] = self.voxelize(inputs['points'], data['data_samples'])
        return {'inputs': batch_inputs, 'data_samples': data['data_samples']}

    # -- voxelisation ---------------------------------------------------------------------------
    @torch.no_grad()
    def voxelize(self, points: List[Tensor], data_samples=None) -> Dict[str, Tensor]:
        if self.voxel_type == 'minkunet':
            return _q.voxelize_minkunet(points, self.voxel_layer.voxel_size, self.batch_first, self.max_voxels, self.training)
        voxel_dict = dict()
        if self.voxel_type == 'hard':
            voxels, coors, num_points, voxel_centers = [], [], [], []
            for i, res in enumerate(points):
                res_voxels, res_coors, res_num_points = self.voxel_layer(res)
                res_voxel_centers = (res_coors[:, [2, 1, 0]] + 0.5) * res_voxels.new_tensor(self.voxel_layer.voxel_size) + \
                    res_voxels.new_tensor(self.voxel_layer.point_cloud_range[0:3])
                voxels.append(res_voxels)
                coors.append(F.pad(res_coors, (1, 0), mode='constant', value=i))
                num_points.append(res_num_points)
                voxel_centers.append(res_voxel_centers)
            voxel_dict['num_points'] = torch.cat(num_points, dim=0)
            voxel_dict['voxel_centers'] = torch.cat(voxel_centers, dim=0)
            voxels, coors = torch.cat(voxels, dim=0), torch.cat(coors, dim=0)
        elif self.voxel_type == 'dynamic':
            coors = [F.pad(self.voxel_layer(res), (1, 0), mode='constant', value=i) for i, res in enumerate(points)]
            voxels = torch.cat(points, dim=0)
            coors = torch.cat(coors, dim=0)
        elif self.voxel_type == 'cylindrical':
            voxels, coors = [], []
            for i, res in enumerate(points):
                polar = torch.stack((torch.sqrt(res[:, 0] ** 2 + res[:, 1] ** 2), torch.atan2(res[:, 1], res[:, 0]), res[:, 2]), dim=-1)
                lo = polar.new_tensor(self.voxel_layer.point_cloud_range[:3])
                hi = polar.new_tensor(self.voxel_layer.point_cloud_range[3:])
                c = torch.floor((torch.clamp(polar, lo, hi) - lo) / polar.new_tensor(self.voxel_layer.voxel_size)).int()
                coors.append(F.pad(c, (1, 0), mode='constant', value=i))
                voxels.append(torch.cat((polar, res[:, :2], res[:, 3:]), dim=-1))
            voxels = torch.cat(voxels, dim=0)
            coors = torch.cat(coors, dim=0)
        else:
            raise ValueError(f'Invalid voxelization type {self.voxel_type}')
        voxel_dict['voxels'] = voxels
        voxel_dict['coors'] = coors
        return voxel_dict

    # host-side helpers kept for API compatibility; the device path lives in gcdlss_b200.quantize
    def ravel_hash(self, x):
        import numpy as np
        assert x.ndim == 2, x.shape
        x = (x - np.min(x, axis=0)).astype(np.uint64, copy=False)
        ext = np.max(x, axis=0).astype(np.uint64) + 1
        h = np.zeros(x.shape[0], dtype=np.uint64)
        for k in range(x.shape[1] - 1):
            h = (h + x[:, k]) * ext[k + 1]
        return h + x[:, -1]

    def sparse_quantize(self, coords, return_index: bool = False, return_inverse: bool = False):
        """Same contract as ref models/voxelizer.py:334-360, computed by the GPU kernels."""
        ic = torch.as_tensor(coords).to(torch.int32)
        dev = torch.device("cuda", torch.cuda.current_device())
        inds, inverse, _ = _q.ops.unique_rows(_q.ops.shift_to_min(ic.to(dev).clone()), order=1)
        outputs = []
        if return_index:
            outputs += [inds.cpu().numpy()]
        if return_inverse:
            outputs += [inverse.cpu().numpy()]
        return outputs
