"""Synthetic datasets shaped like the reference's HDF5 files (utils/load.py:18-24 upstream:
`input` (N,1,H,W), `output` (N,3,H,W), float).  The reference only downloads data
(scripts/download_datasets.sh); there is no network here, so inputs are generated:

  GRF KLE-m:   K = exp(G),  G = sum_{i<m} sqrt(lambda_i) phi_i xi_i,  (lambda, phi) the leading
               eigenpairs of the exponential covariance exp(-|s-s'|/l) on the unit-square grid,
               xi ~ N(0,1).  l is not stated anywhere upstream; l = 0.1 follows the only hint
               (a developer path `grf_exp/ls0.1_...`, utils/image_gradient.py:312).
  channelized: two-valued positive field from thresholded smooth noise.
"""
import os

import numpy as np
import torch


def kle_basis(imsize, n_kle=512, length_scale=0.1, device=None):
    """sqrt(lambda_i) * phi_i for the leading n_kle modes, shape (n_kle, imsize*imsize), float64."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    g = (torch.arange(imsize, dtype=torch.float64, device=device) + 0.5) / imsize
    yy, xx = torch.meshgrid(g, g, indexing="ij")
    pts = torch.stack([yy.reshape(-1), xx.reshape(-1)], 1)
    cov = torch.exp(-torch.cdist(pts, pts) / length_scale)
    lam, phi = torch.linalg.eigh(cov)
    n_kle = min(n_kle, lam.numel())
    lam, phi = lam[-n_kle:].flip(0).clamp_min(0), phi[:, -n_kle:].flip(1)
    return (phi * lam.sqrt()).t().contiguous()


def grf_kle(n, imsize, n_kle=512, length_scale=0.1, seed=1, device=None, chunk=1024):
    """(n,1,imsize,imsize) float32 log-normal permeability fields on the CPU."""
    basis = kle_basis(imsize, n_kle, length_scale, device)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out = torch.empty(n, 1, imsize, imsize, dtype=torch.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        xi = torch.randn(e - s, basis.shape[0], generator=gen, dtype=torch.float64).to(basis.device)
        out[s:e] = torch.exp(xi @ basis).reshape(e - s, 1, imsize, imsize).float().cpu()
    return out


def channelized(n, imsize, seed=1, lo=1.0, hi=float(np.exp(2.5))):
    """(n,1,imsize,imsize) float32 two-valued fields from thresholded low-pass noise."""
    rs = np.random.RandomState(seed)
    z = rs.standard_normal((n, imsize, imsize))
    f = np.fft.rfft2(z)
    ky = np.fft.fftfreq(imsize)[:, None]
    kx = np.fft.rfftfreq(imsize)[None, :]
    f *= np.exp(-((ky * 3.0) ** 2 + (kx * 12.0) ** 2) * 40.0)
    s = np.fft.irfft2(f, s=(imsize, imsize))
    s = (s - s.mean((1, 2), keepdims=True)) / (s.std((1, 2), keepdims=True) + 1e-12)
    return torch.tensor(np.where(s > 0.3, hi, lo)[:, None], dtype=torch.float32)


def write_hdf5(path, inputs, outputs=None):
    """Write the reference's HDF5 layout (npz-backed stand-in when h5py is not installed)."""
    try:
        import h5py
    except ImportError:
        from ._shims import h5py
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with h5py.File(path, 'w') as f:
        f.create_dataset('input', data=np.asarray(inputs, dtype=np.float32))
        if outputs is not None:
            f.create_dataset('output', data=np.asarray(outputs, dtype=np.float32))
