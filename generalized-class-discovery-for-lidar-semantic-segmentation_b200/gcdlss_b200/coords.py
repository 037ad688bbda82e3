"""Coordinate manager: per-resolution coordinate maps, their hash tables and cached kernel maps.

Plays the role of MinkowskiEngine's CoordinateManager for the layer types MinkUNet uses
(ref models/minkunet.py:62-128): one instance is created by ``SparseTensor`` and shared by every
tensor derived from it, so a teacher and a student that consume the same input tensor
(ref modules/exp_merge_mean_teacher.py:2802-2805) share all maps, as they do under ME.
"""
from __future__ import annotations

import torch

from . import config, ops


TABLE_LOG = None      # instrumentation (bench.py): when a list, every NeighbourTable built is appended to it


class CoordMap:
    __slots__ = ("coords", "n", "table", "runs", "tensor_stride", "parent", "code")

    def __init__(self, coords, table, tensor_stride, runs=None):
        self.coords, self.n, self.table, self.tensor_stride = coords, coords.shape[0], table, tensor_stride
        self.runs = runs     # run table (config "runs" kernel-map search); `table` is the point-wise one (may be None then)
        self.parent = None   # [n] row of the 2x coarser map (filled when that map is created)
        self.code = None     # [n] child offset index dx + 2dy + 4dz


class NeighbourTable:
    """One dense neighbour table ``nbr [kv, n_out]`` (int32, -1 = none) plus what is derived from it on first use:
    the tile-sorted copy for the tcgen05 convolution and the per-offset pair lists for wgrad.  A table references
    nothing but device tensors, so kernel maps can share tables (a stride-2 map's dgrad table is the matching
    transposed map's forward table) without reference cycles and without going back to the manager."""
    __slots__ = ("nbr", "kv", "n_out", "_sorted", "_pairs")

    def __init__(self, nbr, kv, n_out):
        self.nbr, self.kv, self.n_out = nbr, kv, n_out
        self._sorted = None          # (tile-sorted table, its column -> output row permutation, per-tile offset masks)
        self._pairs = None

    def tc_table(self):
        """(table, out_rows, tile_masks) for the tcgen05 convolution: the tile-sorted copy when tile sorting is on and this
        is a 3x3x3 or 2x2x2 table with enough rows to pay for the sort, else (nbr, None, None)."""
        if self.nbr is None or self.kv not in (8, 27) or not config.get_tile_sort() or self.n_out < config.tile_sort_min_rows():
            return self.nbr, None, None
        if self._sorted is None:
            self._sorted = ops.kmap_tile_sort(self.nbr)
        return self._sorted

    @property
    def pairs(self):
        if self._pairs is None and self.nbr is not None:
            self._pairs = ops.pairs_from_table(self.nbr)
        return self._pairs

    def device_tensors(self):
        out = [self.nbr] if self.nbr is not None else []
        if self._pairs is not None:
            out += list(self._pairs)
        if self._sorted is not None:
            out += list(self._sorted)
        return out


class KernelMap:
    """Forward table of a convolution plus the table that drives the matching dgrad (see include/gcdlss_b200.h,
    gcd_conv_args).  Both are held STRONGLY (as NeighbourTable objects, i.e. device tensors): an autograd node that
    keeps its KernelMap keeps everything backward will read, whatever happened to the SparseTensors and their
    coordinate manager in the meantime (a Lightning ``training_step`` returns only the loss,
    ref modules/exp.py:249-267)."""
    __slots__ = ("fwd", "back", "n_in", "n_out", "kv", "back_mirror")

    def __init__(self, fwd: NeighbourTable, n_in, n_out, back, back_mirror):
        self.fwd, self.back = fwd, back
        self.n_in, self.n_out, self.kv, self.back_mirror = n_in, n_out, fwd.kv, back_mirror

    @property
    def nbr(self):
        return self.fwd.nbr

    def tc_table(self):
        return self.fwd.tc_table()

    def tc_back_table(self):
        """The same for the table that drives the matching dgrad."""
        return self.back.tc_table() if self.back is not None else (None, None, None)

    @property
    def pairs(self):
        return self.fwd.pairs

    @property
    def back_nbr(self):
        return self.back.nbr if self.back is not None else None

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
        self._tables = {}      # kernel-map key -> NeighbourTable (a stride-2 map and its transpose share theirs)

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
        for table in self._tables.values():
            out += table.device_tensors()
        return out

    # -- kernel maps -----------------------------------------------------------------------
    def _table(self, key, build) -> NeighbourTable:
        t = self._tables.get(key)
        if t is None:
            nbr, kv, n_out = build()
            t = self._tables[key] = NeighbourTable(nbr, kv, n_out)
            if TABLE_LOG is not None:
                TABLE_LOG.append(t)
        return t

    def kernel_map(self, ts_in: int, kernel_size: int, stride: int, transposed: bool) -> KernelMap:
        key = (ts_in, kernel_size, stride, transposed)
        km = self._kmaps.get(key)
        if km is not None:
            return km
        if kernel_size == 1 and stride == 1:
            m = self.get_map(ts_in)
            km = KernelMap(self._table(key, lambda: (None, 1, m.n)), m.n, m.n, None, False)
        elif stride == 1 and kernel_size in (3, 5) and not transposed:
            m = self.get_map(ts_in)

            def build():
                if self.runs:
                    if m.runs is None:       # coarse maps: built on first use from the map's unique coordinates
                        m.runs = ops.runtable_build(m.coords, ts_in, self.status)
                    return ops.kmap_subm_runs(m.coords, m.runs, kernel_size, ts_in), kernel_size ** 3, m.n
                return ops.kmap_subm(m.coords, m.table, kernel_size, ts_in), kernel_size ** 3, m.n
            t = self._table(key, build)
            # stride-1 symmetric kernel: the transposed map is the same table with mirrored offsets
            km = KernelMap(t, m.n, m.n, t, True)
        elif stride == 2 and kernel_size == 2:
            # A stride-2 map and the transposed map between the same two resolutions are each other's dgrad table;
            # both tables are built now (three small launches) so that neither map ever needs the manager again.
            ts_fine = ts_in // 2 if transposed else ts_in
            if transposed and (ts_in % 2 or ts_fine not in self.maps):
                raise RuntimeError("transposed convolution needs the finer coordinate map to exist already "
                                   "(MinkUNet decoders only upsample onto encoder maps)")
            fine = self.get_map(ts_fine)
            coarse = self.get_map(ts_fine * 2)
            down = self._table((ts_fine, 2, 2, False), lambda: (ops.kmap_down2(fine.parent, fine.code, coarse.n), 8, coarse.n))
            up = self._table((ts_fine * 2, 2, 2, True), lambda: (ops.kmap_up2(fine.parent, fine.code), 8, fine.n))
            km = KernelMap(up, coarse.n, fine.n, down, False) if transposed else KernelMap(down, fine.n, coarse.n, up, False)
        else:
            raise NotImplementedError(f"kernel_size={kernel_size}, stride={stride}, transposed={transposed} is not on the "
                                      "MinkUNet path (supported: 1/1, 3/1, 5/1, 2/2 and transposed 2/2)")
        self._kmaps[key] = km
        return km
