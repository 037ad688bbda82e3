"""Host logic of bench.py that does not need a GPU: the tile-sort self-check gate must never raise and never let a
NaN into the bench's JSON line, and the self-check script must always print a verdict."""
import json
import os
import subprocess
import sys
import types

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _fake_run(stdout="", returncode=0, raises=None):
    def run(cmd, **kw):
        if raises is not None:
            raise raises
        return types.SimpleNamespace(stdout=stdout, stderr="boom", returncode=returncode)
    return run


def test_gate_parses_the_last_json_line(monkeypatch):
    monkeypatch.setattr(bench.subprocess, "run", _fake_run('noise\n{"ok": true, "reason": "fine", "ms_sorted": 9.5, "logits_rel_diff": NaN}\n'))
    v = bench.tile_sort_selfcheck(0, 4, "kitti", 17)
    assert v["ok"] is True and v["ms_sorted"] == 9.5 and v["logits_rel_diff"] is None
    json.loads(json.dumps(v, allow_nan=False))                      # strict JSON


def test_gate_reads_every_failure_as_off(monkeypatch):
    monkeypatch.setattr(bench.subprocess, "run", _fake_run("Traceback ...\n", returncode=1))
    assert bench.tile_sort_selfcheck(0, 4, "kitti", 17)["ok"] is False
    monkeypatch.setattr(bench.subprocess, "run", _fake_run(raises=subprocess.TimeoutExpired("x", 120)))
    v = bench.tile_sort_selfcheck(0, 4, "kitti", 17)
    assert v["ok"] is False and "TimeoutExpired" in v["reason"]
    monkeypatch.setattr(bench.subprocess, "run", _fake_run("{not json}\n"))
    assert bench.tile_sort_selfcheck(0, 4, "kitti", 17)["ok"] is False


def test_selfcheck_script_always_prints_a_verdict():
    # no GPU here: the script must still exit 0 with {"ok": false, "reason": ...}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "selfcheck_tilesort.py"), "0", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0
    verdict = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    import torch
    if not torch.cuda.is_available():
        assert verdict["ok"] is False and verdict["reason"]


def test_selfcheck_key_bits_match_the_device_code():
    # the self-check's torch reference of the sort key must use the bit order of csrc/tilesort.cuh
    import importlib.util
    from test_tile_sort_model import bit_order
    spec = importlib.util.spec_from_file_location("selfcheck_tilesort", os.path.join(ROOT, "tools", "selfcheck_tilesort.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.presence_bits(27).tolist() == list(bit_order())
    assert mod.presence_bits(8).tolist() == list(range(8))
