"""Global math-mode switch of the sparse-conv path.

``fp32``: every kernel computes and stores fp32 (SIMT FMA convolution); meets the 1e-4 parity
bound against the oracle.  ``bf16``: features are stored in bf16 between layers and the
convolutions run on tcgen05 tensor cores with fp32 accumulation in TMEM; parameters, BN
statistics, weight gradients and everything a caller sees through ``SparseTensor.F`` stay fp32.
"""
import os

_MODES = ("fp32", "bf16")
_state = {"mode": os.environ.get("GCDLSS_MATH", "fp32")}
if _state["mode"] not in _MODES:
    raise ValueError(f"GCDLSS_MATH must be one of {_MODES}")


def set_math_mode(mode: str) -> None:
    if mode not in _MODES:
        raise ValueError(f"math mode must be one of {_MODES}")
    _state["mode"] = mode


def get_math_mode() -> str:
    return _state["mode"]


# Kernel-map search of the stride-1 convolutions.  "points": one hash slot per voxel (gcd_hash_build + gcd_kmap_subm).
# "runs": one 32-byte slot per run of four x-adjacent cells (gcd_runtable_build + gcd_kmap_subm_runs, csrc/runtable.cuh):
# the same maps from 2.4-2.8x fewer scattered loads.  The run table's logic is held to the oracle on the CPU by
# tests/test_emulated_kernels.py; it becomes the default once tests/test_gpu_zz_runtable.py has passed on a B200.
_KMAP = ("points", "runs")
_state["kmap"] = os.environ.get("GCDLSS_KMAP", "points")
if _state["kmap"] not in _KMAP:
    raise ValueError(f"GCDLSS_KMAP must be one of {_KMAP}")


def set_kmap_search(kind: str) -> None:
    if kind not in _KMAP:
        raise ValueError(f"kernel-map search must be one of {_KMAP}")
    _state["kmap"] = kind


def get_kmap_search() -> str:
    return _state["kmap"]
