"""gcdlss_b200 — B200-native (sm_100a) point->voxel quantisation + MinkUNet sparse-conv backbone.

Host-side mirror of the MinkowskiEngine surface the GCDLSS reference calls; all arithmetic runs in
hand-written CUDA kernels behind the C ABI of ``libgcdlss_sm100a.so`` (include/gcdlss_b200.h).
"""
from . import _cabi
from .config import get_kmap_search, get_math_mode, get_tile_sort, set_kmap_search, set_math_mode, set_tile_sort
from .coords import CoordinateManager, KernelMap
from .functional import devoxelize, voxelize_reduce
from .nn import (BasicBlock, Bottleneck, MinkowskiBatchNorm, MinkowskiConvolution, MinkowskiConvolutionTranspose, MinkowskiDropout,
                 MinkowskiLinear, MinkowskiReLU, cat, kaiming_normal_)
from .quantize import batched_coordinates, sparse_quantize, voxelize_minkunet
from .sparse_tensor import CoordinateMapKey, SparseTensor

__version__ = "0.1.0"


def build(force: bool = False) -> str:
    """Compile libgcdlss_sm100a.so for sm_100a (nvcc; works without a GPU)."""
    return _cabi.build(force)
