"""Oracle: coordinate maps and kernel maps (CPU, numpy).  Test infrastructure only.

Restates what MinkowskiEngine's CoordinateManager computes for the layer types the reference's
MinkUNet uses (``models/minkunet.py:62-128``) [ME-upstream semantics, SURVEY 8(a) a8, a12, a13]:

* kernel offsets: odd K centred (-K//2..K//2), even K one-sided (0..K-1), first spatial axis (x)
  fastest in the offset index k;
* stride-1 conv (K=3,5): out coords == in coords, ``out[o] = sum_k in[o + off_k*ts] W[k]``;
* stride-2 conv (K=2): out coords = unique ``floor(c / (2 ts)) * 2 ts``; tensor stride doubles;
* transposed stride-2 conv (K=2): out coords = the finer, already existing map; one pair per
  fine voxel: ``out[f] = in[parent(f)] W[code(f)]``.

Canonical forms defined here (ME's own orders are hash-internal, so they cannot be followed):
* coarse voxels are numbered in **first-occurrence order of their children** in the fine map;
* a kernel map is the dense neighbour table ``nbr[Nout, K^3]`` (int32, -1 = no input);
  the per-offset pair lists are derived from it, pairs of one offset sorted by output row.

The implementation deliberately avoids hashing (sort + searchsorted on packed keys), so it is
independent of the open-addressing hash the CUDA path uses.
"""
from __future__ import annotations

import numpy as np

from .quantize import unique_first_occurrence

# 64-bit key packing shared with the CUDA side: 10 bits batch | 3 x 18 bits biased coords.
COORD_BITS = 18
COORD_BIAS = 1 << (COORD_BITS - 1)
BATCH_BITS = 10


def pack_keys(coords: np.ndarray) -> np.ndarray:
    """[N,4] int32 (b,x,y,z) -> [N] uint64; numeric order == lexicographic (b,x,y,z) order."""
    c = coords.astype(np.int64)
    b, x, y, z = c[:, 0], c[:, 1] + COORD_BIAS, c[:, 2] + COORD_BIAS, c[:, 3] + COORD_BIAS
    if c.shape[0]:
        if b.min() < 0 or b.max() >= (1 << BATCH_BITS):
            raise ValueError("batch index out of the 10-bit key range")
        for v in (x, y, z):
            if v.min() < 0 or v.max() >= (1 << COORD_BITS):
                raise ValueError("voxel coordinate out of the 18-bit key range")
    return ((b << (3 * COORD_BITS)) | (x << (2 * COORD_BITS)) | (y << COORD_BITS) | z).astype(np.uint64)


def kernel_offsets(kernel_size: int) -> np.ndarray:
    """[K^3, 3] int32 offsets, x fastest."""
    r = np.arange(kernel_size, dtype=np.int32)
    if kernel_size % 2 == 1:
        r = r - kernel_size // 2
    oz, oy, ox = np.meshgrid(r, r, r, indexing="ij")
    return np.stack([ox.reshape(-1), oy.reshape(-1), oz.reshape(-1)], 1).astype(np.int32)


class KeyIndex:
    """Sorted-key lookup: coordinates -> row index (or -1)."""

    def __init__(self, coords: np.ndarray):
        keys = pack_keys(coords)
        self.order = np.argsort(keys, kind="stable")
        self.sorted = keys[self.order]
        if self.sorted.shape[0] > 1 and np.any(self.sorted[1:] == self.sorted[:-1]):
            raise ValueError("duplicate coordinates in a coordinate map")

    def lookup(self, coords: np.ndarray) -> np.ndarray:
        c = coords.astype(np.int64)
        ok = np.ones(c.shape[0], bool)
        for d in (1, 2, 3):
            ok &= (c[:, d] + COORD_BIAS >= 0) & (c[:, d] + COORD_BIAS < (1 << COORD_BITS))
        q = np.zeros(c.shape[0], np.uint64)
        q[ok] = pack_keys(coords[ok])
        pos = np.searchsorted(self.sorted, q)
        pos_c = np.minimum(pos, max(self.sorted.shape[0] - 1, 0))
        hit = ok & (self.sorted.shape[0] > 0)
        if self.sorted.shape[0]:
            hit &= self.sorted[pos_c] == q
        out = np.full(c.shape[0], -1, np.int32)
        out[hit] = self.order[pos_c[hit]].astype(np.int32)
        return out


def kmap_subm(coords: np.ndarray, kernel_size: int, tensor_stride: int) -> np.ndarray:
    """Stride-1 kernel map: nbr[N, K^3] int32, nbr[o,k] = row of coords[o] + off_k*ts or -1."""
    offs = kernel_offsets(kernel_size)
    n = coords.shape[0]
    nbr = np.full((n, offs.shape[0]), -1, np.int32)
    if n == 0:
        return nbr
    index = KeyIndex(coords)
    for k, off in enumerate(offs):
        q = coords.copy()
        q[:, 1:] += off[None, :] * tensor_stride
        nbr[:, k] = index.lookup(q)
    return nbr


def stride2(coords: np.ndarray, tensor_stride: int):
    """Coarse map of a stride-2 conv.

    Returns (coarse [M,4] int32, parent [N] int32, code [N] int32) where
    coarse[parent[f]] == floor(coords[f] / (2 ts)) * (2 ts) (floor division, so negatives round
    down) and code = dx + 2 dy + 4 dz with d* = (c - parent_c) / ts in {0,1}.
    """
    s = 2 * tensor_stride
    pc = coords.copy()
    pc[:, 1:] = np.floor_divide(coords[:, 1:], s) * s
    umap, inv = unique_first_occurrence(pc)
    d = (coords[:, 1:] - pc[:, 1:]) // tensor_stride
    code = d[:, 0] + 2 * d[:, 1] + 4 * d[:, 2]
    return pc[umap].astype(np.int32), inv.astype(np.int32), code.astype(np.int32)


def kmap_down2(parent: np.ndarray, code: np.ndarray, n_coarse: int) -> np.ndarray:
    """Stride-2 K=2 conv kernel map as a table nbr[M, 8]: child row with offset code k, or -1."""
    nbr = np.full((n_coarse, 8), -1, np.int32)
    nbr[parent, code] = np.arange(parent.shape[0], dtype=np.int32)
    return nbr


def kmap_up2(parent: np.ndarray, code: np.ndarray) -> np.ndarray:
    """Transposed stride-2 K=2 conv: nbr[Nfine, 8] with the single entry nbr[f, code[f]] = parent[f]."""
    n = parent.shape[0]
    nbr = np.full((n, 8), -1, np.int32)
    nbr[np.arange(n), code] = parent
    return nbr


def pairs_from_table(nbr: np.ndarray):
    """Per-offset pair lists derived from a neighbour table.

    Returns (pair_in [P], pair_out [P], offsets [KV+1]); the pairs of offset k are
    pair_*[offsets[k]:offsets[k+1]], sorted by output row.
    """
    kv = nbr.shape[1]
    ins, outs, offs = [], [], [0]
    for k in range(kv):
        o = np.nonzero(nbr[:, k] >= 0)[0]
        ins.append(nbr[o, k])
        outs.append(o.astype(np.int32))
        offs.append(offs[-1] + o.shape[0])
    cat = lambda l: np.concatenate(l).astype(np.int32) if l else np.zeros(0, np.int32)
    return cat(ins), cat(outs), np.asarray(offs, np.int64)


def brute_force_subm(coords: np.ndarray, kernel_size: int, tensor_stride: int) -> np.ndarray:
    """O(N^2) neighbour search, used only by property tests on tiny inputs."""
    offs = kernel_offsets(kernel_size)
    n = coords.shape[0]
    nbr = np.full((n, offs.shape[0]), -1, np.int32)
    for o in range(n):
        for k, off in enumerate(offs):
            tgt = coords[o].copy()
            tgt[1:] += off * tensor_stride
            hit = np.nonzero((coords == tgt[None, :]).all(1))[0]
            if hit.shape[0]:
                nbr[o, k] = hit[0]
    return nbr


class CoordLevels:
    """All coordinate maps and kernel maps one MinkUNet forward needs (5 resolutions)."""

    def __init__(self, coords: np.ndarray, n_levels: int = 5):
        self.coords = [np.ascontiguousarray(coords, dtype=np.int32)]
        self.parent, self.code = [], []
        for lvl in range(n_levels - 1):
            c, p, k = stride2(self.coords[lvl], 1 << lvl)
            self.coords.append(c)
            self.parent.append(p)
            self.code.append(k)
        self._subm = {}

    def subm(self, level: int, kernel_size: int) -> np.ndarray:
        key = (level, kernel_size)
        if key not in self._subm:
            self._subm[key] = kmap_subm(self.coords[level], kernel_size, 1 << level)
        return self._subm[key]

    def down(self, level: int) -> np.ndarray:
        """Map of the stride-2 conv from ``level`` to ``level+1``."""
        return kmap_down2(self.parent[level], self.code[level], self.coords[level + 1].shape[0])

    def up(self, level: int) -> np.ndarray:
        """Map of the transposed conv from ``level+1`` back to ``level``."""
        return kmap_up2(self.parent[level], self.code[level])
