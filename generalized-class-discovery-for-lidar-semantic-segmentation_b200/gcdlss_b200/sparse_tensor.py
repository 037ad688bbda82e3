"""SparseTensor: (coordinates, features) pair bound to a CoordinateManager.

API subset of ``ME.SparseTensor`` that the reference touches: construction from
``features=`` / ``coordinates=`` (ref modules/exp.py:259), ``.F .features .C .coordinates
.coordinate_map_key .coordinate_manager .tensor_stride .device``, ``+`` / ``+=`` on tensors that
share a coordinate map.  Row order is the caller's order (callers rely on it, e.g.
``logits[coords[:,0]==i]``, ref modules/exp_merge_mean_teacher.py:2276).
"""
from __future__ import annotations

import torch

from .coords import CoordinateManager


class CoordinateMapKey:
    """Identifies a coordinate map inside its manager: the tensor stride."""

    def __init__(self, tensor_stride: int, dimension: int = 3):
        self._ts, self._d = int(tensor_stride), dimension

    def get_tensor_stride(self):
        return [self._ts] * self._d

    def get_key(self):
        return (self.get_tensor_stride(), "")

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and other._ts == self._ts and other._d == self._d

    def __hash__(self):
        return hash((self._ts, self._d))

    def __repr__(self):
        return f"CoordinateMapKey(tensor_stride={self.get_tensor_stride()})"


class SparseTensor:
    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None, coordinate_manager=None,
                 quantization_mode=None, device=None, requires_grad=None, **_ignored):
        if not isinstance(features, torch.Tensor) or features.dim() != 2:
            raise ValueError("features must be a 2-D tensor [N, C]")
        if device is not None:
            features = features.to(device)
        if isinstance(tensor_stride, (list, tuple)):
            if len(set(tensor_stride)) != 1:
                raise NotImplementedError("anisotropic tensor strides are not on the MinkUNet path")
            tensor_stride = tensor_stride[0]
        if coordinate_manager is None:
            if coordinates is None:
                raise ValueError("coordinates or (coordinate_map_key, coordinate_manager) required")
            if not features.is_cuda:
                raise RuntimeError("gcdlss_b200 runs on CUDA (sm_100a) only: move features to the GPU; there is no CPU fallback")
            if coordinates.shape[0] != features.shape[0]:
                raise ValueError(f"{coordinates.shape[0]} coordinates for {features.shape[0]} feature rows")
            if int(tensor_stride) != 1:
                raise NotImplementedError("new coordinate managers start at tensor stride 1")
            coordinate_manager = CoordinateManager(coordinates.to(features.device))
            coordinate_map_key = CoordinateMapKey(1)
        elif coordinate_map_key is None:
            coordinate_map_key = CoordinateMapKey(tensor_stride)
        if requires_grad:
            features = features.requires_grad_(True)
        self._F = features
        self.coordinate_manager = coordinate_manager
        self.coordinate_map_key = coordinate_map_key

    # ---- ME-style accessors ---------------------------------------------------------------
    @property
    def F(self) -> torch.Tensor:
        """Features as the caller sees them: always fp32 (bf16 storage is an internal detail)."""
        return self._F if self._F.dtype == torch.float32 else self._F.float()

    features = F

    @property
    def C(self) -> torch.Tensor:
        self.coordinate_manager.check()
        return self.coordinate_manager.get_map(self.tensor_stride_int).coords

    coordinates = C

    @property
    def tensor_stride_int(self) -> int:
        return self.coordinate_map_key.get_tensor_stride()[0]

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def D(self) -> int:
        return 3

    dimension = D

    @property
    def device(self):
        return self._F.device

    @property
    def dtype(self):
        return torch.float32

    @property
    def shape(self):
        return self._F.shape

    def size(self, *a):
        return self._F.size(*a)

    def __len__(self):
        return self._F.shape[0]

    @property
    def requires_grad(self):
        return self._F.requires_grad

    def detach(self):
        return self._like(self._F.detach())

    def _like(self, feats) -> "SparseTensor":
        return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key, coordinate_manager=self.coordinate_manager)

    def _check_same_map(self, other):
        if not isinstance(other, SparseTensor):
            raise TypeError("expected a SparseTensor")
        if other.coordinate_manager is not self.coordinate_manager or other.coordinate_map_key != self.coordinate_map_key:
            raise ValueError("SparseTensors do not share a coordinate map")

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            self._check_same_map(other)
            a, b = self._F, other._F
            if a.dtype != b.dtype:
                a, b = a.float(), b.float()
            return self._like(a + b)
        return self._like(self._F + other)

    def __iadd__(self, other):
        return self.__add__(other)   # ME's in-place add; a fresh tensor keeps autograd simple

    def features_at(self, batch_index: int):
        return self.F[self.C[:, 0] == batch_index]

    def coordinates_at(self, batch_index: int):
        c = self.C
        return c[c[:, 0] == batch_index][:, 1:]

    @property
    def decomposed_features(self):
        c = self.C
        nb = int(c[:, 0].max().item()) + 1 if c.shape[0] else 0
        return [self.F[c[:, 0] == b] for b in range(nb)]

    def __repr__(self):
        return f"SparseTensor(features={tuple(self._F.shape)}, dtype={self._F.dtype}, tensor_stride={self.tensor_stride})"
