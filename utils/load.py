"""Dataset loading with the reference's signature (utils/load.py:11-37 upstream): an HDF5 file
with `input` (N,1,H,W) and `output` (N,3,H,W) -> shuffling DataLoader(drop_last=True)."""
import json
from argparse import Namespace

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

try:
    import h5py
except ImportError:  # container without h5py: npz-backed stand-in with the same File[...] surface
    from pde_surrogate_b200._shims import h5py


def load_args(run_dir):
    with open(run_dir + '/args.txt') as f:
        return Namespace(**json.load(f))


class ResidentLoader(object):
    """DataLoader(TensorDataset(...), shuffle=True, drop_last=True) semantics (utils/load.py:34-35 upstream) over
    tensors that live on ONE device: the whole dataset (64 MB at 4096 x 64 x 64) is copied to the GPU once, every
    epoch draws a fresh device permutation and yields device-resident batches, so the script's per-step
    `input.to(device)` (train_codec_mixed_residual.py:225) is a no-op instead of a host-to-device copy.
    Selected with PDES_DATA_DEVICE=cuda[:i]; the default stays the reference's CPU DataLoader."""

    def __init__(self, tensors, batch_size, device, shuffle=True, drop_last=True):
        self.device = torch.device(device)
        self.tensors = [t.to(self.device) for t in tensors]
        self.dataset = TensorDataset(*tensors)          # host view: the script reads .dataset[0][1].numel()
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.n = self.tensors[0].shape[0]

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else -(-self.n // self.batch_size)

    def __iter__(self):
        idx = torch.randperm(self.n, device=self.device) if self.shuffle else torch.arange(self.n, device=self.device)
        for i in range(len(self)):
            sel = idx[i * self.batch_size:(i + 1) * self.batch_size]
            yield tuple(t.index_select(0, sel) for t in self.tensors)


def load_data(hdf5_file, ndata, batch_size, only_input=True, return_stats=False):
    with h5py.File(hdf5_file, 'r') as f:
        x = np.asarray(f['input'][:ndata])
        print(f'x_data: {x.shape}')
        y = None
        if not only_input:
            y = np.asarray(f['output'][:ndata])
            print(f'y_data: {y.shape}')
    stats = {}
    if return_stats:
        if y is None:
            raise ValueError('return_stats=True needs only_input=False')
        stats['y_variation'] = ((y - y.mean(0, keepdims=True)) ** 2).sum(axis=(0, 2, 3))
    tensors = [torch.as_tensor(x, dtype=torch.float32)]
    if y is not None:
        tensors.append(torch.as_tensor(y, dtype=torch.float32))
    import os
    dev = os.environ.get("PDES_DATA_DEVICE", "")
    if dev.startswith("cuda") and torch.cuda.is_available():
        loader = ResidentLoader(tensors, batch_size, dev)
    else:
        loader = DataLoader(TensorDataset(*tensors), batch_size=batch_size, shuffle=True, drop_last=True)
    print(f'Loaded dataset: {hdf5_file}')
    return loader, stats
