"""A dataset that quantises inside ``__getitem__`` the way the reference's datasets do (ref utils/dataset_remission.py:868-873);
imported by DataLoader worker processes of tests/test_gpu_quantize_workers.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _paths  # noqa: E402,F401


class QuantisingDataset(torch.utils.data.Dataset):
    def __init__(self, n_items=4, n_points=3000, q=0.05):
        self.n_items, self.n_points, self.q = n_items, n_points, q

    def __len__(self):
        return self.n_items

    def points(self, i):
        return np.random.default_rng(100 + i).normal(0, 2.0, (self.n_points, 3)).astype(np.float32)

    def __getitem__(self, i):
        import MinkowskiEngine as ME
        pts = self.points(i)
        c, um, inv = ME.utils.sparse_quantize(coordinates=pts, return_index=True, return_inverse=True, quantization_size=self.q)
        return {"i": i, "coords": torch.from_numpy(c), "unique_map": um, "inverse_map": inv}


def collate(items):
    return items
