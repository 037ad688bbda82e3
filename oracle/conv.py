"""Oracle: sparse convolution, batch norm, devoxelisation (CPU, torch).  Test infrastructure only.

Restates MinkowskiEngine's per-offset gather -> GEMM -> scatter-add [ME-upstream, SURVEY 3.4 and
8(a) a12-a14] with differentiable torch ops, so ``torch.autograd`` yields the reference
gradients (dgrad / wgrad) without a second hand-written formula.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def conv_table(feats: torch.Tensor, nbr, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """out[o] = sum_k feats[nbr[o,k]] @ weight[k]  (skipping nbr == -1), + bias.

    feats [Nin, Cin]; nbr [Nout, KV] (numpy or tensor); weight [KV, Cin, Cout].
    """
    nbr = torch.as_tensor(np.asarray(nbr)).long()
    n_out, kv = nbr.shape
    out = feats.new_zeros((n_out, weight.shape[-1]))
    for k in range(kv):
        o = torch.nonzero(nbr[:, k] >= 0).reshape(-1)
        if o.numel() == 0:
            continue
        out = out.index_add(0, o, feats.index_select(0, nbr[o, k]) @ weight[k])
    if bias is not None:
        out = out + bias.reshape(1, -1)
    return out


def conv_1x1(feats: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """kernel_size 1, stride 1: plain ``F.mm(kernel)`` with a 2-D kernel [Cin, Cout] (a12)."""
    out = feats @ weight
    if bias is not None:
        out = out + bias.reshape(1, -1)
    return out


def batch_norm(x, weight, bias, running_mean, running_var, training: bool, momentum: float = 0.1, eps: float = 1e-5):
    """ME.MinkowskiBatchNorm == nn.BatchNorm1d over all rows of F [N,C] (a14)."""
    return F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)


def devox_gather(voxel_feats: torch.Tensor, inverse_map) -> torch.Tensor:
    """Voxel -> point devoxelisation: ``feats[inverse_map]`` (ref models/decoder.py:416-424,
    modules/exp_merge_mean_teacher.py:2845-2846).  Its autograd backward is the segmented sum."""
    return voxel_feats.index_select(0, torch.as_tensor(np.asarray(inverse_map)).long())


def point_to_voxel(point_feats: torch.Tensor, inverse_map, n_voxels: int, mode: str = "mean") -> torch.Tensor:
    """Point -> voxel reduce (mmcv DynamicScatter as used by ref models/encoder.py:121-164)."""
    idx = torch.as_tensor(np.asarray(inverse_map)).long()
    c = point_feats.shape[1]
    if mode == "max":
        out = point_feats.new_full((n_voxels, c), float("-inf"))
        out = out.scatter_reduce(0, idx[:, None].expand(-1, c), point_feats, reduce="amax", include_self=True)
        return out
    out = point_feats.new_zeros((n_voxels, c)).index_add(0, idx, point_feats)
    if mode == "mean":
        cnt = torch.bincount(idx, minlength=n_voxels).clamp(min=1).to(point_feats.dtype)
        out = out / cnt[:, None]
    return out


# ----------------------------------------------------------------------------------------------
# Independent pin: the same convolutions computed on a dense grid with torch's conv3d.
def dense_conv_reference(coords: np.ndarray, feats: torch.Tensor, weight: torch.Tensor, kernel_size: int,
                         tensor_stride: int = 1):
    """Stride-1 sparse conv evaluated densely (one scan, batch column ignored).

    Scatter feats into a dense [1,Cin,X,Y,Z] grid, run F.conv3d with the weight re-laid from
    [K^3 (x fastest), Cin, Cout] to torch's [Cout, Cin, kx, ky, kz], read back at active sites.
    """
    K = kernel_size
    xyz = coords[:, 1:] // tensor_stride
    lo = xyz.min(0)
    p = xyz - lo
    shape = p.max(0) + 1
    cin, cout = weight.shape[1], weight.shape[2]
    grid = feats.new_zeros((1, cin, *shape.tolist()))
    grid[0, :, p[:, 0], p[:, 1], p[:, 2]] = feats.t()
    # weight index k = ix + K*iy + K*K*iz  ->  w[iz,iy,ix] -> permute to [Cout,Cin,ix,iy,iz]
    w = weight.reshape(K, K, K, cin, cout).permute(4, 3, 2, 1, 0).contiguous()
    out = F.conv3d(grid, w, padding=K // 2)
    return out[0, :, p[:, 0], p[:, 1], p[:, 2]].t()
