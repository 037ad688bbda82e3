"""Host logic of bench.py that does not need a GPU: the reference arm (the CPU oracle) must run without importing the
product package, on the same scans as the GPU arm, and the oracle's own parameter table must describe the same network
as the product's module tree."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_oracle_parameter_table_equals_the_module_tree():
    from models import minkunet as mu
    from oracle.minkunet import random_params
    for arch, cls in (("MinkUNet34C", mu.MinkUNet34C), ("MinkUNet14A", mu.MinkUNet14A), ("MinkUNet50", mu.MinkUNet50)):
        sd = cls(1, 17).state_dict()
        p = random_params(arch, 1, 17)
        assert set(p) == set(sd), (arch, sorted(set(p) ^ set(sd))[:6])
        assert all(tuple(p[k].shape) == tuple(sd[k].shape) for k in sd), arch
    k = random_params("MinkUNet34C", 1, 17)["block1.0.conv1.kernel"]
    assert abs(float(k.std()) - (2.0 / (27 * 32)) ** 0.5) < 0.1 * (2.0 / (27 * 32)) ** 0.5      # kaiming fan_out (ref models/resnet.py:62-69)


def test_both_arms_draw_the_same_scans():
    assert [bench.scan_index(0, 0, s, 3, 4) for s in range(4)] == [0, 1, 2, 3]
    assert bench.scan_index(1, 0, 0, 3, 4) == 12 and bench.scan_index(0, 2, 3, 3, 4) == 11
    synth = bench.load_synth()
    from gcdlss_b200 import synth as product_synth
    a, b = synth.make_scan("kitti", 3, n_points=2000), product_synth.make_scan("kitti", 3, n_points=2000)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_no_scan_loses_the_sensor_to_a_box():
    # a box containing the sensor used to swallow every ray (scans 3 and 8 had 2-4 k voxels)
    synth = bench.load_synth()
    from oracle import quantize as oq
    for i in (3, 8):
        xyz, _ = synth.make_scan("kitti", i)
        assert xyz.shape[0] > 100_000 and oq.sparse_quantize_me(xyz, 0.05)[0].shape[0] > 40_000


def test_reference_arm_runs_without_the_product(tmp_path):
    code = ("import sys, json; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--workload','nuscenes_b16'];"
            "import bench; bench.WORKLOADS['nuscenes_b16']=('nuscenes',16,3000,14); bench.main()")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["product_modules_imported"] == [] and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
