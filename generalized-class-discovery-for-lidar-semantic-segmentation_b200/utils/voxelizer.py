"""Drop-in for the reference ``utils/voxelizer.py``: host-side augmentation matrices.

Despite its name this class does no voxelisation: ``get_transformation_matrix``
returns a 4x4 uniform-scale matrix and a 4x4 rotation(+translation) matrix in float64
(ref utils/voxelizer.py:41-74).  The only change is ``collections.abc.Iterable`` (the reference's
``collections.Iterable`` was removed in Python 3.10).
"""
import collections.abc

import numpy as np
from scipy import linalg


def M(axis, theta):
    """Rotation by ``theta`` about ``axis`` (matrix exponential of the cross-product matrix)."""
    return linalg.expm(np.cross(np.eye(3), axis / linalg.norm(axis) * theta))


class Voxelizer:
    def __init__(self, voxel_size=0.05, clip_bound=None, use_augmentation=False, scale_augmentation_bound=None,
                 rotation_augmentation_bound=None, translation_augmentation_ratio_bound=None, ignore_label=255):
        self.voxel_size = voxel_size
        self.clip_bound = clip_bound
        self.ignore_label = ignore_label if ignore_label is not None else -100
        self.use_augmentation = use_augmentation
        self.scale_augmentation_bound = scale_augmentation_bound
        self.rotation_augmentation_bound = rotation_augmentation_bound
        self.translation_augmentation_ratio_bound = translation_augmentation_ratio_bound

    def get_transformation_matrix(self):
        scale_m, rigid_m = np.eye(4), np.eye(4)
        # The order and number of np.random draws matches the reference so a seeded run reproduces it:
        # one uniform per bounded axis, one shuffle, then the scale, then the translation.
        rot = np.eye(3)
        if self.use_augmentation and self.rotation_augmentation_bound is not None:
            if not isinstance(self.rotation_augmentation_bound, collections.abc.Iterable):
                raise ValueError()
            per_axis = []
            for axis_ind, bound in enumerate(self.rotation_augmentation_bound):
                axis = np.zeros(3)
                axis[axis_ind] = 1
                theta = np.random.uniform(*bound) if bound is not None else 0
                per_axis.append(M(axis, theta))
            np.random.shuffle(per_axis)
            rot = per_axis[0] @ per_axis[1] @ per_axis[2]
        rigid_m[:3, :3] = rot
        scale = 1
        if self.use_augmentation and self.scale_augmentation_bound is not None:
            scale *= np.random.uniform(*self.scale_augmentation_bound)
        np.fill_diagonal(scale_m[:3, :3], scale)
        if self.use_augmentation and self.translation_augmentation_ratio_bound is not None:
            rigid_m[:3, 3] = [np.random.uniform(*t) for t in self.translation_augmentation_ratio_bound]
        return scale_m, rigid_m
