"""Host plumbing of the run-table mode (GCDLSS_KMAP=runs) without a GPU: the real ``CoordinateManager`` runs on CPU tensors
with the library calls replaced by the host emulation of csrc/runtable.cuh (tests/emu) and the oracle's stride-2 maps.
Checks that every stride-1 kernel map it hands out -- level 1 from the table built at construction, coarse levels from
tables built on first use with the right tensor stride -- equals the oracle's, and that duplicates are reported."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import small_cloud
from oracle import coords as ocd
from test_emulated_kernels import build as emu_build, emu, kmap as emu_kmap  # noqa: F401  (emu is a fixture)


@pytest.fixture()
def runs_manager(emu, monkeypatch):  # noqa: F811
    import gcdlss_b200
    from gcdlss_b200 import coords, ops
    built = []

    class FakeRunTable:
        def __init__(self, slots, cap):
            self.slots, self.cap = slots, cap

    def runtable_build(c, ts, status):
        slots, cap, st = emu_build(emu, c.numpy(), ts)
        status |= st
        built.append(ts)
        return FakeRunTable(slots, cap)

    def kmap_subm_runs(c, table, kernel_size, ts):
        return torch.from_numpy(emu_kmap(emu, c.numpy(), table.slots, table.cap, kernel_size, ts))

    def coords_stride2(c, ts, status):
        coarse, parent, code = ocd.stride2(c.numpy(), ts)
        return torch.from_numpy(coarse), torch.from_numpy(parent), torch.from_numpy(code), None      # no point-wise coarse table needed

    as_cols = lambda t: torch.from_numpy(np.ascontiguousarray(t.T))
    monkeypatch.setattr(ops, "kmap_down2", lambda parent, code, n_coarse: as_cols(ocd.kmap_down2(parent.numpy(), code.numpy(), n_coarse)))
    monkeypatch.setattr(ops, "kmap_up2", lambda parent, code: as_cols(ocd.kmap_up2(parent.numpy(), code.numpy())))
    monkeypatch.setattr(ops, "new_batch", lambda: None)
    monkeypatch.setattr(ops, "runtable_build", runtable_build)
    monkeypatch.setattr(ops, "kmap_subm_runs", kmap_subm_runs)
    monkeypatch.setattr(ops, "coords_stride2", coords_stride2)
    monkeypatch.setattr(ops, "hash_build", lambda *a: pytest.fail("runs mode must not build the point-wise table"))
    monkeypatch.setattr(ops, "kmap_subm", lambda *a: pytest.fail("runs mode must not search the point-wise table"))
    prev = gcdlss_b200.get_kmap_search()
    gcdlss_b200.set_kmap_search("runs")
    yield coords.CoordinateManager, built
    gcdlss_b200.set_kmap_search(prev)


def test_all_levels_equal_the_oracle(runs_manager):
    CoordinateManager, built = runs_manager
    bc = np.concatenate([small_cloud(41, 2500, spread=0.5, batch=0), small_cloud(42, 1200, spread=0.5, batch=1)])
    mgr = CoordinateManager(torch.from_numpy(bc))
    assert mgr.runs and mgr.maps[1].table is None and built == [1]
    lv = ocd.CoordLevels(bc)
    for level in range(4):
        ts = 1 << level
        km = mgr.kernel_map(ts, 3, 1, False)
        np.testing.assert_array_equal(km.nbr.numpy().T, lv.subm(level, 3))
        assert mgr.kernel_map(ts, 3, 1, False) is km                     # cached
    np.testing.assert_array_equal(mgr.kernel_map(1, 5, 1, False).nbr.numpy().T, lv.subm(0, 5))
    assert built == [1, 2, 4, 8]                                          # one run table per level, built once, with its stride
    mgr.check()
    assert any(t is mgr.maps[2].runs.slots for t in mgr.device_tensors())  # handed over to the consumer stream with the rest


def test_duplicates_are_reported(runs_manager):
    CoordinateManager, _ = runs_manager
    c = torch.tensor([[0, 1, 2, 3], [0, 4, 4, 4], [0, 1, 2, 3]], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="duplicate"):
        CoordinateManager(c).check()


def test_kernel_maps_outlive_their_manager(runs_manager):
    """A Lightning training_step returns only the loss (ref modules/exp.py:249-267): by the time backward runs the
    SparseTensors and their coordinate manager are gone.  Every map must carry its own dgrad table."""
    import gc
    import weakref
    CoordinateManager, _ = runs_manager
    bc = small_cloud(43, 2500, spread=0.5, batch=0)
    lv = ocd.CoordLevels(bc)
    mgr = CoordinateManager(torch.from_numpy(bc))
    km3 = mgr.kernel_map(1, 3, 1, False)
    km_down = mgr.kernel_map(1, 2, 2, False)       # asked for alone: an encoder-only network never builds the transposed map
    km_1x1 = mgr.kernel_map(2, 1, 1, False)
    ref = weakref.ref(mgr)
    del mgr
    gc.collect()
    assert ref() is None, "kernel maps must not keep the manager alive (no map <-> manager cycle)"
    assert km3.back_nbr is km3.nbr and km3.tc_back_table()[0] is km3.nbr
    parent, code = lv.parent[0], lv.code[0]
    np.testing.assert_array_equal(km_down.nbr.numpy().T, ocd.kmap_down2(parent, code, lv.coords[1].shape[0]))
    np.testing.assert_array_equal(km_down.back_nbr.numpy().T, ocd.kmap_up2(parent, code))
    assert km_down.tc_back_table()[0] is km_down.back_nbr
    assert km_1x1.back_nbr is None and km_1x1.tc_back_table() == (None, None, None)
