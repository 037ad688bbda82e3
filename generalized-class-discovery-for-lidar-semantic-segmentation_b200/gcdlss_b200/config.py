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
# the same maps from 2.4-2.8x fewer scattered loads.  Default since the whole GPU suite passed under it on a B200 (round 2,
# profiles/r2_call1_tests.txt); the run table's logic is also held to the oracle on the CPU by tests/test_emulated_kernels.py.
_KMAP = ("points", "runs")
_state["kmap"] = os.environ.get("GCDLSS_KMAP", "runs")
if _state["kmap"] not in _KMAP:
    raise ValueError(f"GCDLSS_KMAP must be one of {_KMAP}")


def set_kmap_search(kind: str) -> None:
    if kind not in _KMAP:
        raise ValueError(f"kernel-map search must be one of {_KMAP}")
    _state["kmap"] = kind


def get_kmap_search() -> str:
    return _state["kmap"]


# Tile sort of the 3x3x3 neighbour tables for the tcgen05 convolution (csrc/tilesort.cuh): columns sorted by the mask of
# present neighbours so that a 128-column tile visits 8-12 kernel offsets instead of 21-25.  On by default (every rank, every
# workload: the whole GPU suite passes with it, profiles/r2_call1_tests.txt); maps with fewer rows than
# ``tile_sort_min_rows`` are left in scan order (the sort is ~20 small launches per map).
_state["tile_sort"] = os.environ.get("GCDLSS_TILE_SORT", "1") not in ("0", "")
_state["tile_sort_min_rows"] = int(os.environ.get("GCDLSS_TILE_SORT_MIN_ROWS", "16384"))


def set_tile_sort(enabled: bool, min_rows: int | None = None) -> None:
    _state["tile_sort"] = bool(enabled)
    if min_rows is not None:
        _state["tile_sort_min_rows"] = int(min_rows)


def get_tile_sort() -> bool:
    return _state["tile_sort"]


def tile_sort_min_rows() -> int:
    return _state["tile_sort_min_rows"]


# NVTX ranges around the phases of a step (map build on the prefetch stream, trunk forward / backward, loss, gradient exchange,
# optimiser): off by default (a range costs two driver calls), on with GCDLSS_NVTX=1 or set_nvtx(True) for nsys / ncu --nvtx
# timelines.  (the reference has no tracing of its own; this is the hook a profiler run needs.)
_state["nvtx"] = os.environ.get("GCDLSS_NVTX", "0") not in ("0", "")


def set_nvtx(enabled: bool) -> None:
    _state["nvtx"] = bool(enabled)


class nvtx_range:
    """``with nvtx_range("name"):`` -- an NVTX range when tracing is on, nothing otherwise."""

    __slots__ = ("name", "on")

    def __init__(self, name: str):
        self.name, self.on = name, _state["nvtx"]

    def __enter__(self):
        if self.on:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if self.on:
            import torch
            torch.cuda.nvtx.range_pop()
        return False
