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
    loader = DataLoader(TensorDataset(*tensors), batch_size=batch_size, shuffle=True, drop_last=True)
    print(f'Loaded dataset: {hdf5_file}')
    return loader, stats
