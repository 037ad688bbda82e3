"""Coordinate manager: per-resolution coordinate maps, their hash tables and cached kernel maps.

Plays the role of MinkowskiEngine's CoordinateManager for the layer types MinkUNet uses
(ref models/minkunet.py:62-128): one instance is created by ``SparseTensor`` and shared by every
tensor derived from it, so a teacher and a student that consume the same input tensor
(ref modules/exp_merge_mean_teacher.py:2802-2805) share all maps, as they do under ME.
"""
from __future__ import annotations

import weakref

import torch

from . import config, ops


class CoordMap:
    __slots__ = ("coords", "n", "table", "runs", "tensor_stride", "parent", "code")

    def __init__(self, coords, table, tensor_stride, runs=None):
        self.coords, self.n, self.table, self.tensor_stride = coords, coords.shape[0], table, tensor_stride
        self.runs = runs     # run table (config "runs" kernel-map search); `table` is the point-wise one (may be None then)
        self.parent = None   # [n] row of the 2x coarser map (filled when that map is created)
        self.code = None     # [n] child offset index dx + 2dy + 4dz


class KernelMap:
    """A dense neighbour table plus, lazily, its per-offset pair lists.

    nbr [kv, n_out] int32 feeds the output-stationary forward; ``back_key`` names the kernel map whose
    table drives the matching dgrad (see include/gcdlss_b200.h, gcd_conv_args).  The manager is held
    weakly: maps must die by reference counting when the batch's tensors die (a manager <-> map cycle
    would park hundreds of MB per step until the cyclic GC runs)."""

    def __init__(self, nbr, n_in, n_out, kv, manager, back_key, back_mirror):
        self.nbr, self.n_in, self.n_out, self.kv = nbr, n_in, n_out, kv
        self._mgr = weakref.ref(manager)
        self._back_key, self.back_mirror = back_key, back_mirror
        self._pairs = None
        self._sorted = None          # (tile-sorted table, its column -> output row permutation), built on first use

    def tc_table(self):
        """(table, out_rows) for the tcgen05 convolution: the tile-sorted copy when tile sorting is on and this is a
        3x3x3 or 2x2x2 map with enough rows to pay for the sort, else (nbr, None).  Pair lists always come from ``nbr``."""
        if self.nbr is None or self.kv not in (8, 27) or not config.get_tile_sort() or self.n_out < config.tile_sort_min_rows():
            return self.nbr, None
        if self._sorted is None:
            self._sorted = ops.kmap_tile_sort(self.nbr)
        return self._sorted

    def tc_back_table(self):
        """The same for the table that drives the matching dgrad."""
        if self._back_key is None:
            return None, None
        if self._back_key == "self":
            return self.tc_table()
        mgr = self._mgr()
        if mgr is None:
            raise RuntimeError("the coordinate manager of this kernel map no longer exists")
        return mgr.kernel_map(*self._back_key).tc_table()

    @property
    def pairs(self):
        if self._pairs is None:
            self._pairs = ops.pairs_from_table(self.nbr) if self.nbr is not None else None
        return self._pairs

    @property
    def back_nbr(self):
        if self._back_key is None:
            return None
        if self._back_key == "self":
            return self.nbr
        mgr = self._mgr()
        if mgr is None:
            raise RuntimeError("the coordinate manager of this kernel map no longer exists")
        return mgr.kernel_map(*self._back_key).nbr

    def num_pairs(self) -> int:
        """Exact pair count (one device read; used for FLOP accounting only)."""
        if self.nbr is None:
            return self.n_out
        return int(self.pairs[2][-1].item())


class CoordinateManager:
    def __init__(self, coords: torch.Tensor):
        if coords.dim() != 2 or coords.shape[1] != 4:
            raise ValueError("coordinates must be [N, 4] (batch, x, y, z)")
        coords = coords.to(torch.int32).contiguous()
        ops.new_batch()
        self.device = coords.device
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.runs = config.get_kmap_search() == "runs"
        if self.runs:    # the run-table build reports duplicates / range errors like gcd_hash_build does
            self.maps = {1: CoordMap(coords, None, 1, ops.runtable_build(coords, 1, self.status))}
        else:
            self.maps = {1: CoordMap(coords, ops.hash_build(coords, self.status), 1)}
        self._kmaps = {}

    # -- coordinate maps -------------------------------------------------------------------
    def check(self):
        """Raise if any device-side build reported a key-range / duplicate / overflow condition."""
        ops._status_check(self.status, "SparseTensor coordinates")

    def get_map(self, tensor_stride: int) -> CoordMap:
        if tensor_stride not in self.maps:
            if tensor_stride < 2 or tensor_stride % 2:
                raise KeyError(f"no coordinate map at tensor stride {tensor_stride}")
            fine = self.get_map(tensor_stride // 2)
            coarse, parent, code, table = ops.coords_stride2(fine.coords, fine.tensor_stride, self.status)
            self.check()  # the call above already synchronised to learn the coarse voxel count
            fine.parent, fine.code = parent, code
            self.maps[tensor_stride] = CoordMap(coarse, table, tensor_stride)
        return self.maps[tensor_stride]

    def prebuild_unet(self, n_levels: int = 5, stem_kernel: int = 5, with_pairs: bool = True):
        """Build everything one MinkUNet forward+backward will ask for (coordinate maps at tensor strides
        1..2^(n_levels-1), the stem / 3x3x3 / stride-2 / transposed kernel maps and, for training, their pair
        lists).  Used by the batch prefetcher so the host syncs of the build happen off the training stream."""
        strides = [1 << l for l in range(n_levels)]
        for ts in strides:
            self.get_map(ts)
        keys = [(1, stem_kernel, 1, False)] + [(ts, 3, 1, False) for ts in strides] + [(ts, 1, 1, False) for ts in strides]
        keys += [(ts, 2, 2, False) for ts in strides[:-1]] + [(ts, 2, 2, True) for ts in strides[1:]]
        for key in keys:
            km = self.kernel_map(*key)
            if with_pairs and km.nbr is not None and km.kv <= 27:
                km.pairs
            if config.get_math_mode() == "bf16":
                km.tc_table()        # tile-sorted copy for the tcgen05 convolution; no-op unless tile sorting is on
        return self

    def device_tensors(self):
        """Every device tensor this manager owns (for cross-stream hand-over)."""
        out = [self.status]
        for m in self.maps.values():
            out.append(m.coords)
            if m.table is not None:
                out += [m.table.keys, m.table.vals]
            if m.runs is not None:
                out.append(m.runs.slots)
            out += [t for t in (m.parent, m.code) if t is not None]
        for km in self._kmaps.values():
            if km.nbr is not None:
                out.append(km.nbr)
            if km._pairs is not None:
                out += list(km._pairs)
            if km._sorted is not None:
                out += list(km._sorted)
        return out

    # -- kernel maps -----------------------------------------------------------------------
    def kernel_map(self, ts_in: int, kernel_size: int, stride: int, transposed: bool) -> KernelMap:
        key = (ts_in, kernel_size, stride, transposed)
        km = self._kmaps.get(key)
        if km is not None:
            return km
        if kernel_size == 1 and stride == 1:
            m = self.get_map(ts_in)
            km = KernelMap(None, m.n, m.n, 1, self, None, False)
        elif stride == 1 and kernel_size in (3, 5) and not transposed:
            m = self.get_map(ts_in)
            if self.runs:
                if m.runs is None:       # coarse maps: built on first use from the map's unique coordinates
                    m.runs = ops.runtable_build(m.coords, ts_in, self.status)
                nbr = ops.kmap_subm_runs(m.coords, m.runs, kernel_size, ts_in)
            else:
                nbr = ops.kmap_subm(m.coords, m.table, kernel_size, ts_in)
            # stride-1 symmetric kernel: the transposed map is the same table with mirrored offsets
            km = KernelMap(nbr, m.n, m.n, kernel_size ** 3, self, "self", True)
        elif stride == 2 and kernel_size == 2 and not transposed:
            fine = self.get_map(ts_in)
            coarse = self.get_map(ts_in * 2)
            nbr = ops.kmap_down2(fine.parent, fine.code, coarse.n)
            km = KernelMap(nbr, fine.n, coarse.n, 8, self, (ts_in * 2, 2, 2, True), False)
        elif stride == 2 and kernel_size == 2 and transposed:
            if ts_in % 2 or ts_in // 2 not in self.maps:
                raise RuntimeError("transposed convolution needs the finer coordinate map to exist already "
                                   "(MinkUNet decoders only upsample onto encoder maps)")
            fine = self.get_map(ts_in // 2)
            coarse = self.get_map(ts_in)
            nbr = ops.kmap_up2(fine.parent, fine.code)
            km = KernelMap(nbr, coarse.n, fine.n, 8, self, (ts_in // 2, 2, 2, False), False)
        else:
            raise NotImplementedError(f"kernel_size={kernel_size}, stride={stride}, transposed={transposed} is not on the "
                                      "MinkUNet path (supported: 1/1, 3/1, 5/1, 2/2 and transposed 2/2)")
        self._kmaps[key] = km
        return km
