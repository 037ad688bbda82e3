"""The C-ABI library loads without a GPU and exports exactly what include/gcdlss_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

import _paths
from gcdlss_b200 import _cabi

HEADER = os.path.join(_paths.ROOT, "include", "gcdlss_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(gcd_[a-z0-9_]+)\s*\(", text))


def test_header_and_binding_agree():
    assert declared_symbols() == set(_cabi.PROTOTYPES)


def test_library_loads_and_exports_every_symbol():
    lib = _cabi.lib()          # builds with nvcc if the .so is missing
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.gcd_abi_version() == 1
    assert lib.gcd_has_tcgen05() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gcd_[a-z0-9_]+)", out))
    assert declared_symbols() <= exported


def test_pure_host_queries():
    lib = _cabi.lib()
    assert lib.gcd_hash_capacity(0) >= 1024
    cap = lib.gcd_hash_capacity(100000)
    assert cap >= 200000 and cap & (cap - 1) == 0
    assert lib.gcd_unique_workspace_bytes(1000) > 0 and lib.gcd_stride2_workspace_bytes(1000) > 0
    assert lib.gcd_conv_tc_supported(96, 96, 27) == 1
    assert lib.gcd_conv_tc_supported(96, 17, 1) == 0 and lib.gcd_conv_tc_supported(1, 32, 125) == 0
    assert lib.gcd_conv_packed_weight_bytes(27, 96, 96) == 27 * 2 * 96 * 128


def test_argument_validation_reports_text():
    lib = _cabi.lib()
    rc = lib.gcd_quantize_f32(None, 3, 10, 5, ctypes.c_float(0.05), 0, None, None)   # dims = 5 is invalid
    assert rc == -1
    assert b"gcd_quantize_f32" in lib.gcd_last_error_string()
    with pytest.raises(RuntimeError, match="kernel_size must be 3 or 5"):
        _cabi.call("gcd_kmap_subm", None, 10, None, None, 1024, 4, 1, None, None)


def test_struct_layout_matches_c():
    # offsets computed by hand from include/gcdlss_b200.h (LP64)
    a = _cabi.ConvArgs
    assert (a.inp.offset, a.nbr.offset, a.kv.offset, a.n_out.offset, a.c_in.offset, a.w.offset) == (0, 24, 32, 40, 48, 56)
    assert (a.mirror.offset, a.bias.offset, a.out.offset, a.in_dtype.offset, a.stats.offset, a.math_mode.offset) == (96, 104, 112, 128, 136, 144)
    assert (a.out_rows.offset, a.tile_masks.offset, a.sched.offset) == (152, 160, 168) and ctypes.sizeof(a) == 176
    w = _cabi.WgradArgs
    assert (w.pair_in.offset, w.n_pairs.offset, w.kv.offset, w.dw.offset, w.n_out.offset, w.math_mode.offset) == (32, 56, 64, 80, 96, 112)
    assert w.sched.offset == 120 and ctypes.sizeof(w) == 128


def test_block_struct_layout_against_gcc(tmp_path):
    """sizeof / offsetof of the fused-block structs as gcc lays them out == the ctypes mirror."""
    fields_u = [f[0] for f in _cabi.ConvBnUnit._fields_]
    fields_b = [f[0] for f in _cabi.BlockArgs._fields_]
    fields_o = [f[0] for f in _cabi.Op._fields_]
    prog = ["#include <stdio.h>", "#include <stddef.h>", f'#include "{HEADER}"', "int main(void){",
            'printf("%zu %zu %zu\\n", sizeof(gcd_convbn), sizeof(gcd_block_args), sizeof(gcd_op));']
    prog += [f'printf("%zu\\n", offsetof(gcd_convbn, {f}));' for f in fields_u]
    prog += [f'printf("%zu\\n", offsetof(gcd_block_args, {f}));' for f in fields_b]
    prog += [f'printf("%zu\\n", offsetof(gcd_op, {f}));' for f in fields_o]
    prog += ["return 0;}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    sizes, offs = [int(v) for v in out[:3]], [int(v) for v in out[3:]]
    assert sizes == [ctypes.sizeof(_cabi.ConvBnUnit), ctypes.sizeof(_cabi.BlockArgs), ctypes.sizeof(_cabi.Op)]
    assert offs[:len(fields_u)] == [getattr(_cabi.ConvBnUnit, f).offset for f in fields_u]
    assert offs[len(fields_u):len(fields_u) + len(fields_b)] == [getattr(_cabi.BlockArgs, f).offset for f in fields_b]
    assert offs[len(fields_u) + len(fields_b):] == [getattr(_cabi.Op, f).offset for f in fields_o]


def test_product_code_never_imports_the_oracle():
    pkg = _paths.PKG_DIR
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
