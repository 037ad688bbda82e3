"""Batch prefetcher: prepares the next batch (H2D copy, GPU quantisation, coordinate hash, kernel maps) on a side
CUDA stream while the current step trains on the main stream.

The map build needs a handful of device->host reads (voxel counts per resolution).  Issued on the training stream
they would wait for the whole previous backward pass and leave the GPU idle until the host catches up again; on the
side stream they only wait for the few small kernels of the build itself.  This is the GPU-resident counterpart of the
reference's DataLoader workers, which quantise the next batch on CPU cores while the GPU trains
(ref utils/dataset_remission.py:868-873, modules/exp.py:176-202).
"""
from __future__ import annotations

import torch

from .config import nvtx_range
from .sparse_tensor import SparseTensor


class PreparedBatch:
    def __init__(self, tensor: SparseTensor, extras, event, owned):
        self.tensor, self.extras, self._event, self._owned = tensor, extras, event, owned

    def get(self):
        """Hand the batch over to the current stream: wait for the build, tell the allocator about the new user."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._event)
        for t in self._owned:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(cur)
        return self.tensor, self.extras


class BatchPrefetcher:
    def __init__(self, device=None, n_levels: int = 5, stem_kernel: int = 5, training: bool = True):
        # high priority: the build is a chain of small kernels with host reads in between (voxel counts per level); behind
        # the training stream's persistent 148-CTA kernels each read would otherwise wait for a free SM slot
        self.stream = torch.cuda.Stream(device=device, priority=-1)
        self.n_levels, self.stem_kernel, self.training = n_levels, stem_kernel, training

    def submit(self, make_inputs, wait_for_current_stream: bool = False) -> PreparedBatch:
        """``make_inputs()`` runs under the side stream and returns (features [N,C], coordinates [N,4], extras);
        it may itself copy from pinned host memory and quantise.  Set ``wait_for_current_stream`` when the inputs
        are produced by work already queued on the current stream."""
        if wait_for_current_stream:
            self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream), nvtx_range("gcd:prefetch (inputs, hash, kernel maps, tile sort)"):
            feats, coords, extras = make_inputs()
            st = SparseTensor(features=feats, coordinates=coords)
            st.coordinate_manager.prebuild_unet(self.n_levels, self.stem_kernel, with_pairs=self.training)
            event = self.stream.record_event()
        owned = [feats, coords] + st.coordinate_manager.device_tensors()
        if isinstance(extras, (list, tuple)):
            owned += [e for e in extras if isinstance(e, torch.Tensor)]
        elif isinstance(extras, torch.Tensor):
            owned.append(extras)
        return PreparedBatch(st, extras, event, owned)
