"""``ME.utils.sparse_quantize`` inside DataLoader workers (ref modules/exp.py:176-202: num_workers=8, quantisation in
``Dataset.__getitem__``).  Spawned workers own a CUDA context and quantise on the GPU, bit-exact against the oracle; forked
workers of a CUDA-initialised parent cannot use CUDA and must say so clearly (there is no CPU fallback)."""
import numpy as np
import pytest
import torch

import _paths  # noqa: F401
from oracle import quantize as oq
from worker_dataset import QuantisingDataset, collate

pytestmark = pytest.mark.gpu


def test_spawned_workers_quantise_on_the_gpu(cuda):
    torch.zeros(1, device="cuda")                 # the parent holds a CUDA context, as a training process does
    ds = QuantisingDataset()
    loader = torch.utils.data.DataLoader(ds, batch_size=2, num_workers=2, collate_fn=collate, multiprocessing_context="spawn")
    seen = 0
    for batch in loader:
        for item in batch:
            c0, um0, inv0 = oq.sparse_quantize_me(ds.points(item["i"]), ds.q)
            assert np.array_equal(item["coords"].numpy(), c0)
            assert np.array_equal(item["unique_map"].numpy(), um0) and np.array_equal(item["inverse_map"].numpy(), inv0)
            seen += 1
    assert seen == len(ds)


def test_forked_workers_fail_with_a_clear_message(cuda):
    torch.zeros(1, device="cuda")
    loader = torch.utils.data.DataLoader(QuantisingDataset(2), batch_size=1, num_workers=1, collate_fn=collate, multiprocessing_context="fork")
    with pytest.raises(RuntimeError, match="multiprocessing_context='spawn'"):
        for _ in loader:
            pass
