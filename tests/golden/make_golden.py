"""Generates tests/golden/*.npz.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Part 1 executes the REFERENCE's own code and freezes its outputs (this is what pins the oracle):
  * ``Voxelizer.ravel_hash`` / ``Voxelizer.sparse_quantize`` are pure-numpy methods of
    /root/reference/models/voxelizer.py (:312-360).  The module cannot be imported (mmdet3d is not
    installable), so the two method definitions are taken from the file by AST and executed
    unchanged against numpy.
  * ``utils.voxelizer.Voxelizer.get_transformation_matrix`` is imported from
    /root/reference/utils/voxelizer.py (with ``collections.Iterable`` aliased, which Python 3.10
    removed).
  * the 'minkunet' voxel branch (:271-302) is re-executed line by line with torch CPU ops and the
    reference's sparse_quantize (``.cuda()`` calls dropped).
Part 2 freezes ORACLE outputs for the MinkowskiEngine-defined parts (floor quantiser, coordinate /
kernel maps, a small conv forward/backward, a tiny MinkUNet): "parity unpinned" fixtures, so the
CUDA tests compare against files instead of a moving oracle.
"""
import ast
import collections
import collections.abc
import importlib.util
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import conv as oc  # noqa: E402
from oracle import coords as ocd  # noqa: E402
from oracle import quantize as oq  # noqa: E402


def reference_voxelizer_methods():
    src = open(os.path.join(REF, "models", "voxelizer.py")).read()
    tree = ast.parse(src)
    wanted = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "Voxelizer":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in ("ravel_hash", "sparse_quantize"):
                    wanted[item.name] = textwrap.dedent(ast.get_source_segment(src, item))
    ns = {"np": np, "List": list}
    for code in wanted.values():
        exec(code, ns)

    class Holder:
        ravel_hash = ns["ravel_hash"]
        sparse_quantize = ns["sparse_quantize"]
    return Holder()


def reference_aug_voxelizer():
    collections.Iterable = collections.abc.Iterable
    spec = importlib.util.spec_from_file_location("ref_utils_voxelizer", os.path.join(REF, "utils", "voxelizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synthetic_points(seed, n, spread=20.0):
    rng = np.random.default_rng(seed)
    pts = rng.normal(0, spread, (n, 3)).astype(np.float32)
    pts[:, 2] = rng.normal(0, 1.5, n)
    dup = rng.integers(0, n, n // 5)             # exact duplicates and near-duplicates
    pts[rng.integers(0, n, n // 5)] = pts[dup]
    half = (np.round(pts[: n // 10] / 0.05) + 0.5) * 0.05   # values on rounding boundaries
    pts[: n // 10] = half.astype(np.float32)
    feat = rng.uniform(0, 1, (n, 1)).astype(np.float32)
    return pts, feat


def part1_reference():
    out = {}
    ref = reference_voxelizer_methods()
    for tag, (seed, n) in {"a": (1, 3000), "b": (2, 257), "c": (3, 1)}.items():
        pts, feat = synthetic_points(seed, n)
        # --- models/voxelizer.py:271-302, executed with the reference's own sparse_quantize ---
        res = torch.from_numpy(np.concatenate([pts, feat], 1))
        voxel_size = res.new_tensor([0.05, 0.05, 0.05])
        res_coors = torch.round(res[:, :3] / voxel_size).int()
        res_coors -= res_coors.min(0)[0]
        res_coors_numpy = res_coors.cpu().numpy()
        h = ref.ravel_hash(res_coors_numpy)
        inds, point2voxel_map = ref.sparse_quantize(res_coors_numpy, return_index=True, return_inverse=True)
        out[f"voxA_{tag}_points"] = res.numpy()
        out[f"voxA_{tag}_hash"] = h
        out[f"voxA_{tag}_coors"] = res_coors.numpy()
        out[f"voxA_{tag}_inds"] = np.asarray(inds, np.int64)
        out[f"voxA_{tag}_inverse"] = np.asarray(point2voxel_map, np.int64).reshape(-1)
    mod = reference_aug_voxelizer()
    np.random.seed(1234)
    v = mod.Voxelizer(voxel_size=0.05, use_augmentation=True, scale_augmentation_bound=(0.95, 1.05),
                      rotation_augmentation_bound=((-np.pi / 20, np.pi / 20), (-np.pi / 20, np.pi / 20), (-np.pi, np.pi)),
                      translation_augmentation_ratio_bound=((-3, 3), (-3, 3), (-0.5, 0.5)))
    mats = [v.get_transformation_matrix() for _ in range(3)]
    out["aug_scale"] = np.stack([m[0] for m in mats])
    out["aug_rigid"] = np.stack([m[1] for m in mats])
    np.savez_compressed(os.path.join(HERE, "reference_pinned.npz"), **out)
    print("reference_pinned.npz:", {k: v.shape for k, v in out.items()})


def part2_oracle():
    out = {}
    # floor quantiser, fp32 and fp64 inputs, 3 and 4 columns (LaserMix quirk: batch column divided too)
    pts, _ = synthetic_points(11, 4000)
    for tag, arr in {"f32": pts, "f64": pts.astype(np.float64) @ np.array([[0.99, 0.01, 0], [-0.01, 0.99, 0], [0, 0, 1.0]])}.items():
        c, um, inv = oq.sparse_quantize_me(arr, 0.05)
        out[f"me_{tag}_in"], out[f"me_{tag}_coords"], out[f"me_{tag}_umap"], out[f"me_{tag}_inv"] = arr, c, um, inv
    p4 = np.concatenate([np.repeat(np.arange(4, dtype=np.float32), 1000)[:, None], pts], 1)
    c, um, inv = oq.sparse_quantize_me(p4, 0.05)
    out["me_b4_in"], out["me_b4_coords"], out["me_b4_umap"], out["me_b4_inv"] = p4, c, um, inv

    # coordinate + kernel maps of a two-scan batch with negative coordinates
    scans = []
    for s in range(2):
        p, _ = synthetic_points(20 + s, 1500, spread=1.2)
        scans.append(oq.sparse_quantize_me(p, 0.05)[0])
    bc = oq.batched_coordinates(scans)
    lv = ocd.CoordLevels(bc)
    out["map_coords0"] = bc
    for l in range(1, 5):
        out[f"map_coords{l}"], out[f"map_parent{l - 1}"], out[f"map_code{l - 1}"] = lv.coords[l], lv.parent[l - 1], lv.code[l - 1]
    for l in range(5):
        out[f"map_subm3_{l}"] = lv.subm(l, 3)
    out["map_subm5_0"] = lv.subm(0, 5)

    # one 3x3x3 conv + BN forward/backward in fp64-accurate fp32
    torch.manual_seed(5)
    n = bc.shape[0]
    x = torch.randn(n, 16, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(27, 16, 32, dtype=torch.float64) * 0.1).requires_grad_(True)
    y = oc.conv_table(x, lv.subm(0, 3), w)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    out["conv_x"], out["conv_w"], out["conv_g"] = x.detach().numpy(), w.detach().numpy(), g.numpy()
    out["conv_y"], out["conv_dx"], out["conv_dw"] = y.detach().numpy(), x.grad.numpy(), w.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "oracle_frozen.npz"), **out)
    print("oracle_frozen.npz:", len(out), "arrays")


def part3_tile_sort():
    """Frozen permutations of the tile sort (csrc/tilesort.cuh) for the tables of oracle_frozen.npz: columns ordered, stably,
    by the presence mask with offsets ranked (number of non-zero components, k) for 3x3x3 and bit k for 2x2x2.
    Plain numpy, written out here independently of the test helpers."""
    frozen = np.load(os.path.join(HERE, "oracle_frozen.npz"))
    offs = ocd.kernel_offsets(3)
    rank = np.lexsort((np.arange(27), (offs != 0).sum(1)))
    bit27 = np.empty(27, np.uint64)
    bit27[rank] = np.arange(27, dtype=np.uint64)
    out = {}
    for name, table, bits in (("subm3_0", frozen["map_subm3_0"], bit27), ("subm3_1", frozen["map_subm3_1"], bit27),
                              ("down_0", ocd.kmap_down2(frozen["map_parent0"], frozen["map_code0"], frozen["map_coords1"].shape[0]),
                               np.arange(8, dtype=np.uint64)),
                              ("up_0", ocd.kmap_up2(frozen["map_parent0"], frozen["map_code0"]), np.arange(8, dtype=np.uint64))):
        keys = ((table >= 0).astype(np.uint64) << bits[None, :]).sum(1)
        out[f"rows_{name}"] = np.argsort(keys, kind="stable").astype(np.int32)
        out[f"keys_{name}"] = keys
    np.savez_compressed(os.path.join(HERE, "tile_sort_frozen.npz"), **out)
    print("tile_sort_frozen.npz:", {k: v.shape for k, v in out.items()})


def reference_lasermix_methods():
    """``mix_transform`` and ``laser_mix_transform`` of ExpMergeDiscover_LaserMix_MeanTeacher
    (/root/reference/modules/exp_merge_mean_teacher.py:1577-1787), taken from the file by AST and executed unchanged (the module
    itself cannot be imported: pytorch_lightning / MinkowskiEngine are not installable)."""
    src = open(os.path.join(REF, "modules", "exp_merge_mean_teacher.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "ExpMergeDiscover_LaserMix_MeanTeacher":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in ("mix_transform", "laser_mix_transform"):
                    exec(textwrap.dedent(ast.get_source_segment(src, item)), ns)

    class Holder:
        semi_train_cfg = dict(pitch_angles=[-25, 3], num_areas=[3, 4, 5, 6])      # ref :1448-1452
        mix_transform = ns["mix_transform"]
        laser_mix_transform = ns["laser_mix_transform"]
    return Holder()


def lasermix_inputs(seed, n_sup_scans, n_unsup_scans, n_points):
    """Point-level dicts as the Stage-2 collate hands them to ``mix_transform`` (batched float coords [P, 4], feats [P, 1],
    labels [P]); scans look like a ground plane + clutter seen from 1.7 m so that every pitch band is populated."""
    rng = np.random.default_rng(seed)

    def scans(n_scans):
        c, f = [], []
        for b in range(n_scans):
            n = n_points + int(rng.integers(0, 200))
            az = rng.uniform(-np.pi, np.pi, n)
            pitch = np.deg2rad(rng.uniform(-30.0, 6.0, n))           # a little beyond [-25, 3]: exercises the clamp
            r = rng.uniform(2.0, 60.0, n)
            xyz = np.stack([r * np.cos(pitch) * np.cos(az), r * np.cos(pitch) * np.sin(az), r * np.sin(pitch)], 1).astype(np.float32)
            c.append(np.concatenate([np.full((n, 1), b, np.float32), xyz], 1))
            f.append(rng.uniform(0, 1, (n, 1)).astype(np.float32))
        return np.concatenate(c), np.concatenate(f)

    sc, sf = scans(n_sup_scans)
    uc, uf = scans(n_unsup_scans)
    sup_labels = rng.integers(0, 17, sc.shape[0]).astype(np.int64)
    pseudo = rng.integers(-1, 17, uc.shape[0]).astype(np.int64)
    return sc, sf, sup_labels, uc, uf, pseudo


def part4_lasermix():
    """Freezes the reference's LaserMix outputs (row order, band edges, labels) and the quantisation of the mixed batch."""
    ref = reference_lasermix_methods()
    out = {}
    for tag, (seed, ns, nu, n) in {"b22": (7, 2, 2, 4000), "b21": (8, 2, 1, 2500), "b22_small": (9, 2, 2, 37)}.items():
        sc, sf, sl, uc, uf, pl = lasermix_inputs(seed, ns, nu, n)
        np.random.seed(100 + seed)
        areas = [int(np.random.choice([3, 4, 5, 6], size=1)[0]) for _ in range(2)]     # what the calls below will draw
        np.random.seed(100 + seed)
        sup = {"coords": torch.from_numpy(sc), "feats": torch.from_numpy(sf), "mapped_labels": torch.from_numpy(sl)}
        unsup = {"coords": torch.from_numpy(uc), "feats": torch.from_numpy(uf)}
        bcoords, feats, labels = ref.mix_transform(sup, unsup, torch.from_numpy(pl))
        # the Stage-2 step quantises the mixed batch inline, batch column included (ref :2856-2861)
        qc, um, inv = oq.sparse_quantize_me(bcoords.numpy(), 0.05)
        for k, v in dict(sup_coords=sc, sup_feats=sf, sup_labels=sl, unsup_coords=uc, unsup_feats=uf, pseudo=pl, areas=np.asarray(areas),
                         mix_bcoords=bcoords.numpy(), mix_feats=feats.numpy(), mix_labels=labels.numpy(), q_coords=qc, q_umap=um, q_inv=inv).items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "lasermix_pinned.npz"), **out)
    print("lasermix_pinned.npz:", {k: v.shape for k, v in out.items() if "mix_bcoords" in k})


if __name__ == "__main__":
    if sys.argv[1:] == ["lasermix"]:
        part4_lasermix()
    elif sys.argv[1:] == ["tile_sort"]:           # derived from oracle_frozen.npz alone: does not need the reference checkout
        part3_tile_sort()
    else:
        part1_reference()
        part2_oracle()
        part3_tile_sort()
        part4_lasermix()
