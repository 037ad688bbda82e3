"""Deterministic synthetic LiDAR scans with SemanticKITTI / nuScenes shapes.

There is no network for datasets, so benchmarks and tests ray-cast a simple street scene: a ground
plane, four "street canyon" walls and random axis-aligned boxes, seen by a spinning multi-beam
sensor.  Seed = 1234 + scan index (``numpy.random.default_rng``).
"""
from __future__ import annotations

import numpy as np

SENSOR_HEIGHT = 1.73
MAX_RANGE = 80.0
SENSOR_CLEARANCE = 2.5      # metres of free space around the sensor in x and y

PRESETS = {
    # name: (beams, elevation top/bottom in degrees, azimuth steps, voxel size, feature scale)
    "kitti": (64, 2.0, -24.8, 1900, 0.05, 1.0),
    "nuscenes": (32, 10.0, -30.0, 1085, 0.10, 255.0),
}


def _ray_dirs(beams, el_top, el_bot, az_steps):
    el = np.deg2rad(np.linspace(el_top, el_bot, beams))
    az = np.linspace(-np.pi, np.pi, az_steps, endpoint=False)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (beams, az_steps))], -1)
    return d.reshape(-1, 3)


def _cast(dirs, rng, n_boxes=40):
    """Nearest hit distance of every ray from origin (0, 0, SENSOR_HEIGHT)."""
    o = np.array([0.0, 0.0, SENSOR_HEIGHT])
    inf = np.full(dirs.shape[0], np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(dirs[:, 2] < 0, -o[2] / dirs[:, 2], inf)                      # ground z = 0
        wy, wx = rng.uniform(8, 15, 2), rng.uniform(40, 70, 2)
        for sign, w in ((1, wy[0]), (-1, wy[1])):                                   # walls y = +-w
            tw = (sign * w - o[1]) / dirs[:, 1]
            t = np.minimum(t, np.where(tw > 0, tw, inf))
        for sign, w in ((1, wx[0]), (-1, wx[1])):                                   # walls x = +-w
            tw = (sign * w - o[0]) / dirs[:, 0]
            t = np.minimum(t, np.where(tw > 0, tw, inf))
        centers = np.stack([rng.uniform(-35, 35, n_boxes), rng.uniform(-7.5, 7.5, n_boxes), np.zeros(n_boxes)], 1)
        sizes = np.stack([rng.uniform(1.5, 5.0, n_boxes), rng.uniform(1.2, 2.5, n_boxes), rng.uniform(1.2, 3.0, n_boxes)], 1)
        # no box on top of the ego vehicle: a box that contains (or hugs) the sensor swallows every ray and leaves a
        # 2-4 k-voxel "scan"; boxes whose footprint comes within SENSOR_CLEARANCE of the origin are pushed out along x
        near = (np.abs(centers[:, 0]) < sizes[:, 0] / 2 + SENSOR_CLEARANCE) & (np.abs(centers[:, 1]) < sizes[:, 1] / 2 + SENSOR_CLEARANCE)
        centers[near, 0] = np.where(centers[near, 0] >= 0, 1.0, -1.0) * (sizes[near, 0] / 2 + SENSOR_CLEARANCE + rng.uniform(0.0, 4.0, int(near.sum())))
        lo = centers - sizes / 2 * np.array([1, 1, 0])
        hi = lo + sizes
        inv = 1.0 / dirs                                                           # slab method, rays x boxes
        t0 = (lo[None, :, :] - o[None, None, :]) * inv[:, None, :]
        t1 = (hi[None, :, :] - o[None, None, :]) * inv[:, None, :]
        tnear = np.nanmax(np.minimum(t0, t1), axis=2)
        tfar = np.nanmin(np.maximum(t0, t1), axis=2)
        hit = (tnear <= tfar) & (tfar > 0)
        tb = np.where(hit, np.where(tnear > 0, tnear, tfar), np.inf).min(axis=1)
        t = np.minimum(t, tb)
    return t


def make_scan(kind: str = "kitti", index: int = 0, n_points: int | None = None):
    """Returns (xyz float32 [N,3], feat float32 [N,1]) of one synthetic sweep.

    ``n_points``: optional random subset of that many points in sorted index order, as the
    reference's training down-sampling does (ref utils/dataset_remission.py:813-816).
    """
    beams, el_top, el_bot, az_steps, _, fscale = PRESETS[kind]
    rng = np.random.default_rng(1234 + index)
    dirs = _ray_dirs(beams, el_top, el_bot, az_steps)
    r = _cast(dirs, rng)
    r = r * (1.0 + rng.normal(0.0, 0.002, r.shape[0]))
    keep = np.isfinite(r) & (r < MAX_RANGE) & (r > 0.5)
    xyz = (dirs[keep] * r[keep, None] + np.array([0.0, 0.0, SENSOR_HEIGHT])).astype(np.float32)
    xyz[:, 2] -= SENSOR_HEIGHT                                                     # sensor frame, like the datasets
    feat = (rng.uniform(0.0, 1.0, (xyz.shape[0], 1)) * fscale).astype(np.float32)
    if n_points is not None and xyz.shape[0] > n_points:
        sel = np.sort(rng.choice(xyz.shape[0], n_points, replace=False))
        xyz, feat = xyz[sel], feat[sel]
    return xyz, feat


def make_dense_scan(index: int = 0, sweeps: int = 10):
    """~1.2 M points: ``sweeps`` KITTI-like sweeps with ego motion 0.5 m * i in x and yaw jitter N(0, 1 deg)."""
    rng = np.random.default_rng(99 + index)
    pts, feats = [], []
    for i in range(sweeps):
        xyz, f = make_scan("kitti", index * sweeps + i)
        yaw = np.deg2rad(rng.normal(0.0, 1.0))
        c, s = np.cos(yaw), np.sin(yaw)
        rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)
        pts.append(xyz @ rot.T + np.array([0.5 * i, 0, 0], np.float32))
        feats.append(f)
    return np.concatenate(pts).astype(np.float32), np.concatenate(feats)


def voxel_size(kind: str) -> float:
    return PRESETS["kitti" if kind == "dense" else kind][4]
