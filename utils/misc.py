"""Host helpers with the reference's names (utils/misc.py:6-34 upstream)."""
import os

import numpy as np
import torch


def mkdir(path):
    os.makedirs(path, exist_ok=True)


def mkdirs(*paths):
    for p in paths:
        mkdir(p)


def to_numpy(x):
    if isinstance(x, np.ndarray):
        return x
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    raise TypeError('Unknown type of input, expected torch.Tensor or np.ndarray, but got {}'.format(type(x)))


def module_size(module):
    from pde_surrogate_b200.codec import module_size as _ms
    return _ms(module)
