"""ctypes binding of libgcdlss_sm100a.so (the C ABI declared in include/gcdlss_b200.h).

There is no CPU fallback: if the shared library is missing and cannot be built, importing this
module raises, and every call raises ``RuntimeError`` with the library's error text on failure.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.normpath(os.path.join(_HERE, "..", "csrc"))
LIB_PATH = os.path.join(_HERE, "libgcdlss_sm100a.so")
_LIB_OVERRIDE = os.environ.get("GCDLSS_LIB_PATH")       # tuning only: the PROFILE=1 build (csrc/Makefile), loaded as is

F32, BF16 = 0, 1
MATH_FP32_SIMT, MATH_BF16_TC = 0, 1
ROUND_FLOOR, ROUND_HALF_EVEN = 0, 1
DEV_KEY_RANGE, DEV_DUPLICATE, DEV_TABLE_FULL = 1, 2, 4
OPT_PAIRS_FUSED, OPT_GATHER_FLAT, OPT_TC_STAGES, OPT_TC_GROUP, OPT_WG_CHUNK_MIN, OPT_TC_WARPS, OPT_BN_FUSED, OPT_KMAP_COOP, OPT_PDL, OPT_WGRAD_SIDE, OPT_DYN_TILES, OPT_DYN_AHEAD, OPT_BN_MASK_FROM_X, OPT_SCAN_LOOKBACK = range(14)


def build(force: bool = False) -> str:
    """Compile every .cu for sm_100a with nvcc (see csrc/Makefile). Returns the library path."""
    if force or not os.path.exists(LIB_PATH) or _stale():
        if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
            raise RuntimeError("libgcdlss_sm100a.so is missing and nvcc is not available to build it")
        env = dict(os.environ)
        env["PATH"] = env.get("PATH", "") + ":/usr/local/cuda/bin"
        cmd = ["make", "-C", _CSRC, "-j", str(min(8, os.cpu_count() or 1))]
        # one builder at a time: under torchrun every rank imports the package, and N concurrent `make`s in the same build
        # directory could leave a rank loading a half-written library.  The others wait for the lock and then find it fresh.
        import fcntl
        os.makedirs(os.path.join(_CSRC, "build"), exist_ok=True)
        with open(os.path.join(_CSRC, "build", ".lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or not os.path.exists(LIB_PATH) or _stale():
                    if force:
                        subprocess.run(["make", "-C", _CSRC, "clean"], env=env, check=True, capture_output=True)
                        os.makedirs(os.path.join(_CSRC, "build"), exist_ok=True)
                    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
                    if r.returncode != 0:
                        raise RuntimeError("building libgcdlss_sm100a.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def _stale() -> bool:
    try:
        t = os.path.getmtime(LIB_PATH)
        srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))]
        srcs.append(os.path.normpath(os.path.join(_HERE, "..", "..", "include", "gcdlss_b200.h")))
        return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))
    except OSError:
        return True


class ConvArgs(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("ld_in", C.c_int64), ("n_in", C.c_int64),
        ("nbr", C.c_void_p), ("kv", C.c_int32), ("n_out", C.c_int64),
        ("c_in", C.c_int32), ("c_out", C.c_int32),
        ("w", C.c_void_p), ("w_packed", C.c_void_p),
        ("w_stride_k", C.c_int64), ("w_stride_c", C.c_int64), ("w_stride_n", C.c_int64),
        ("mirror", C.c_int32),
        ("bias", C.c_void_p), ("out", C.c_void_p), ("ld_out", C.c_int64),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("stats", C.c_void_p), ("math_mode", C.c_int32),
        ("out_rows", C.c_void_p), ("tile_masks", C.c_void_p), ("sched", C.c_void_p),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("ld_in", C.c_int64),
        ("gout", C.c_void_p), ("ld_gout", C.c_int64),
        ("pair_in", C.c_void_p), ("pair_out", C.c_void_p), ("pair_off", C.c_void_p),
        ("n_pairs", C.c_int64),
        ("kv", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32),
        ("dw", C.c_void_p), ("dbias", C.c_void_p), ("n_out", C.c_int64),
        ("in_dtype", C.c_int32), ("gout_dtype", C.c_int32), ("math_mode", C.c_int32),
        ("sched", C.c_void_p),
    ]


class ConvBnUnit(C.Structure):
    """gcd_convbn: one conv -> batch-norm unit of a fused block."""
    _fields_ = [
        ("nbr", C.c_void_p), ("n_in", C.c_int64), ("n_out", C.c_int64),
        ("back_nbr", C.c_void_p), ("back_mirror", C.c_int32),
        ("pair_in", C.c_void_p), ("pair_out", C.c_void_p), ("pair_off", C.c_void_p), ("n_pairs", C.c_int64),
        ("kv", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32),
        ("w", C.c_void_p), ("w_packed_fwd", C.c_void_p), ("w_packed_bwd", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
        ("eps", C.c_float), ("momentum", C.c_float),
        ("stats", C.c_void_p), ("sums", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
        ("dw", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("out_rows", C.c_void_p), ("back_out_rows", C.c_void_p),
        ("tile_masks", C.c_void_p), ("back_tile_masks", C.c_void_p),
    ]


class BlockArgs(C.Structure):
    """gcd_block_args."""
    _fields_ = [
        ("u1", ConvBnUnit), ("u2", ConvBnUnit), ("ud", ConvBnUnit),
        ("has_u2", C.c_int32), ("has_ud", C.c_int32), ("relu1", C.c_int32), ("dtype", C.c_int32),
        ("x", C.c_void_p), ("ld_x", C.c_int64),
        ("y1", C.c_void_p), ("a1", C.c_void_p), ("y2", C.c_void_p), ("yd", C.c_void_p), ("rd", C.c_void_p), ("out", C.c_void_p),
        ("gout", C.c_void_p), ("dy2", C.c_void_p), ("dres", C.c_void_p), ("da1", C.c_void_p), ("dy1", C.c_void_p),
        ("dyd", C.c_void_p), ("dx", C.c_void_p), ("dxd", C.c_void_p),
        ("need_dx", C.c_int32), ("launches", C.c_int32), ("ld_gout", C.c_int64),
    ]


OP_BLOCK_FORWARD, OP_BLOCK_BACKWARD, OP_COPY_COLS, OP_ADD_COLS, OP_RECORD_EVENT = range(5)


class Op(C.Structure):
    """gcd_op: one entry of a gcd_run_ops program."""
    _fields_ = [
        ("op", C.c_int32), ("dtype", C.c_int32), ("block", C.POINTER(BlockArgs)),
        ("dst", C.c_void_p), ("ld_dst", C.c_int64), ("src", C.c_void_p), ("ld_src", C.c_int64),
        ("n", C.c_int64), ("c", C.c_int32), ("reserved", C.c_int32),
    ]


_i32, _i64, _vp, _sz, _f32, _f64 = C.c_int32, C.c_int64, C.c_void_p, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes).  Every symbol declared in include/gcdlss_b200.h is listed here;
# tests/test_cabi_symbols.py checks the two against each other.
PROTOTYPES = {
    "gcd_last_error_string": (C.c_char_p, []),
    "gcd_abi_version": (_i32, []),
    "gcd_has_tcgen05": (_i32, []),
    "gcd_set_option": (_i32, [_i32, _i32]),
    "gcd_get_option": (_i32, [_i32]),
    "gcd_quantize_f32": (_i32, [_vp, _i64, _i64, _i32, _f32, _i32, _vp, _vp]),
    "gcd_quantize_f64": (_i32, [_vp, _i64, _i64, _i32, _f64, _i32, _vp, _vp]),
    "gcd_affine_f64": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "gcd_colmin_i32": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "gcd_sub_cols_i32": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "gcd_hash_capacity": (_i64, [_i64]),
    "gcd_unique_workspace_bytes": (_sz, [_i64]),
    "gcd_unique_rows": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gcd_hash_build": (_i32, [_vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "gcd_stride2_workspace_bytes": (_sz, [_i64]),
    "gcd_coords_stride2": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gcd_kmap_subm": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "gcd_kmap_down2": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "gcd_kmap_up2": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "gcd_runtable_slot_bytes": (_sz, []),
    "gcd_runtable_build": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp, _vp]),
    "gcd_kmap_subm_runs": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp]),
    "gcd_tile_sort_workspace_bytes": (_sz, [_i64]),
    "gcd_kmap_tile_sort": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gcd_pairs_workspace_bytes": (_sz, [_i64, _i32]),
    "gcd_pairs_from_table": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gcd_conv_forward": (_i32, [C.POINTER(ConvArgs), _vp]),
    "gcd_conv_packed_weight_bytes": (_sz, [_i32, _i32, _i32]),
    "gcd_conv_pack_weights": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "gcd_conv_wgrad": (_i32, [C.POINTER(WgradArgs), _vp]),
    "gcd_conv_pack_weights_batched": (_i32, [_vp, _i32, _i32, _vp]),
    "gcd_conv_tc_supported": (_i32, [_i32, _i32, _i32]),
    "gcd_im2col": (_i32, [_vp, _i64, _i32, _vp, _i32, _i64, _vp, _i64, _i32, _i32, _vp]),
    "gcd_bn_stats": (_i32, [_vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "gcd_bn_finalize": (_i32, [_vp, _i64, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcd_bn_apply_train": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp]),
    "gcd_bn_fold_eval": (_i32, [_i32, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp]),
    "gcd_bn_apply": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp]),
    "gcd_bn_backward_reduce": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "gcd_bn_backward_apply": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _i32,
                                     _vp, _i64, _vp, _i64, _vp, _vp, _i32, _vp]),
    "gcd_relu": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "gcd_relu_backward": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "gcd_rows_gather": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "gcd_csr_workspace_bytes": (_sz, [_i64, _i64]),
    "gcd_csr_build": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "gcd_consistency_rows": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "gcd_block_forward": (_i32, [C.POINTER(BlockArgs), _vp]),
    "gcd_block_backward": (_i32, [C.POINTER(BlockArgs), _vp]),
    "gcd_run_ops": (_i32, [C.POINTER(Op), _i32, _vp, C.POINTER(_i32)]),
    "gcd_exec_create": (_i32, [C.POINTER(_vp)]),
    "gcd_exec_destroy": (_i32, [_vp]),
    "gcd_run_ops_exec": (_i32, [_vp, C.POINTER(Op), _i32, _vp, C.POINTER(_i32)]),
    "gcd_segment_reduce": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp]),
}

_lib = None


def lib():
    """The loaded library (built on first use if the .so is missing or older than its sources)."""
    global _lib
    if _lib is None:
        path = _LIB_OVERRIDE or LIB_PATH
        if not _LIB_OVERRIDE and (not os.path.exists(path) or (_stale() and shutil.which("nvcc"))):
            path = build()
        handle = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().gcd_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libgcdlss_sm100a {what} failed (code {rc}): {msg}")


_fn_cache = {}


def call(name: str, *args):
    """Call an int32-status entry point and raise on a non-zero return."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(lib(), name)
    rc = fn(*args)
    if rc != 0:
        check(rc, name)
