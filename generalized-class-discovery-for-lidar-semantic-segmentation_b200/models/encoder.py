"""Drop-in for the reference ``models/encoder.py``: ``SegVFE`` point->voxel feature encoder of the Cylinder3D-style
path (ref models/encoder.py:23-170) without the mmcv / mmdet3d registries.  Same constructor; the point MLP is plain
torch (Linear + BatchNorm1d + ReLU); the point->voxel reduction that the reference delegates to mmcv's
``DynamicScatter`` (``:121-123, :164``) runs on the sm_100a hash / counting-sort / segmented-reduce kernels.
"""
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from gcdlss_b200 import ops


class _DynamicScatterFunction(torch.autograd.Function):
    """Reduce point features into their voxels.  Forward: ``gcd_unique_rows`` (ascending voxel order, as mmcv's
    ``torch.unique(coors, dim=0)``), ``gcd_csr_build``, ``gcd_segment_reduce``.  Backward: mean -> grad / count gathered back
    to the points; max -> grad routed to the points that attain the maximum."""

    @staticmethod
    def forward(ctx, feats, coors, mode):
        valid = (coors >= 0).all(1)
        idx = torch.nonzero(valid).reshape(-1)
        f = feats.detach().index_select(0, idx).float().contiguous()
        c = coors.index_select(0, idx).to(torch.int32).contiguous()
        uniq, inv, _ = ops.unique_rows(c, order=1)
        m = uniq.shape[0]
        seg_off, order = ops.csr_build(inv, m)
        out = ops.segment_reduce(f, seg_off, order, m, 2 if mode == "max" else 1)
        voxel_coors = c.index_select(0, uniq)
        ctx.mode, ctx.n_points = mode, feats.shape[0]
        counts = (seg_off[1:] - seg_off[:-1]).to(torch.float32)
        ctx.save_for_backward(idx, inv, f, out, counts)
        ctx.mark_non_differentiable(voxel_coors)
        return out.to(feats.dtype), voxel_coors

    @staticmethod
    def backward(ctx, gout, _gcoors):
        idx, inv, f, out, counts = ctx.saved_tensors
        g = ops.rows_gather(gout.float().contiguous(), inv)
        if ctx.mode == "max":
            g = g * (f == ops.rows_gather(out, inv)).to(g.dtype)
        else:
            g = g / counts.index_select(0, inv)[:, None]
        gfeats = g.new_zeros((ctx.n_points, g.shape[1]))
        gfeats.index_copy_(0, idx, g)
        return gfeats, None, None


class DynamicScatter(nn.Module):
    """mmcv.ops.DynamicScatter(voxel_size, point_cloud_range, average_points) equivalent: coordinates are taken as given
    (already voxelised), rows with a negative coordinate are dropped, voxels come out in ascending coordinate order."""

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super().__init__()
        self.voxel_size, self.point_cloud_range, self.average_points = voxel_size, point_cloud_range, average_points

    def forward(self, points: Tensor, coors: Tensor) -> Tuple[Tensor, Tensor]:
        return _DynamicScatterFunction.apply(points, coors, "mean" if self.average_points else "max")


class SegVFE(nn.Module):
    def __init__(self, in_channels: int = 6, feat_channels: Sequence[int] = [], with_voxel_center: bool = False,
                 voxel_size: Optional[Sequence[float]] = None, grid_shape: Sequence[float] = (480, 360, 32),
                 point_cloud_range: Sequence[float] = (0, -3.14159265359, -4, 50, 3.14159265359, 2),
                 norm_cfg: dict = dict(type='BN1d', eps=1e-5, momentum=0.1), mode: bool = 'max', with_pre_norm: bool = True,
                 feat_compression: Optional[int] = None, return_point_feats: bool = False) -> None:
        super().__init__()
        assert mode in ['avg', 'max']
        assert len(feat_channels) > 0
        assert not (voxel_size and grid_shape), 'voxel_size and grid_shape cannot be setting at the same time'
        if with_voxel_center:
            in_channels += 3
        self.in_channels = in_channels
        self._with_voxel_center = with_voxel_center
        self.return_point_feats = return_point_feats
        self.point_cloud_range = point_cloud_range
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        if voxel_size:
            self.voxel_size = voxel_size
            self.grid_shape = torch.round((pcr[3:] - pcr[:3]) / torch.tensor(voxel_size, dtype=torch.float32)).long().tolist()
        elif grid_shape:
            self.voxel_size = ((pcr[3:] - pcr[:3]) / (torch.tensor(grid_shape, dtype=torch.float32) - 1)).tolist()
            self.grid_shape = grid_shape
        else:
            raise ValueError('must assign a value to voxel_size or grid_shape')
        self.vx, self.vy, self.vz = self.voxel_size
        self.x_offset = self.vx / 2 + point_cloud_range[0]
        self.y_offset = self.vy / 2 + point_cloud_range[1]
        self.z_offset = self.vz / 2 + point_cloud_range[2]

        def norm(ch):
            return nn.BatchNorm1d(ch, eps=norm_cfg.get('eps', 1e-5), momentum=norm_cfg.get('momentum', 0.1))

        widths = [self.in_channels] + list(feat_channels)
        self.pre_norm = norm(self.in_channels) if with_pre_norm else None
        layers = []
        for i in range(len(widths) - 1):
            if i == len(widths) - 2:
                layers.append(nn.Linear(widths[i], widths[i + 1]))
            else:
                layers.append(nn.Sequential(nn.Linear(widths[i], widths[i + 1]), norm(widths[i + 1]), nn.ReLU(inplace=True)))
        self.vfe_layers = nn.ModuleList(layers)
        self.vfe_scatter = DynamicScatter(self.voxel_size, self.point_cloud_range, (mode != 'max'))
        self.compression_layers = None
        if feat_compression is not None:
            self.compression_layers = nn.Sequential(nn.Linear(widths[-1], feat_compression), nn.ReLU())

    def forward(self, features: Tensor, coors: Tensor, *args, **kwargs) -> Tuple[Tensor]:
        parts = [features]
        if self._with_voxel_center:
            centre = features.new_zeros(size=(features.size(0), 3))
            centre[:, 0] = features[:, 0] - (coors[:, 1].type_as(features) * self.vx + self.x_offset)
            centre[:, 1] = features[:, 1] - (coors[:, 2].type_as(features) * self.vy + self.y_offset)
            centre[:, 2] = features[:, 2] - (coors[:, 3].type_as(features) * self.vz + self.z_offset)
            parts.append(centre)
        features = torch.cat(parts[::-1], dim=-1)
        if self.pre_norm is not None:
            features = self.pre_norm(features)
        point_feats = []
        for vfe in self.vfe_layers:
            features = vfe(features)
            point_feats.append(features)
        voxel_feats, voxel_coors = self.vfe_scatter(features, coors)
        if self.compression_layers is not None:
            voxel_feats = self.compression_layers(voxel_feats)
        if self.return_point_feats:
            return voxel_feats, voxel_coors, point_feats
        return voxel_feats, voxel_coors
