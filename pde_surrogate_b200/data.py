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


def darcy_fv_solve(K):
    """Cell-centred finite-volume solve of the reference's Darcy problem on one (H, W) permeability field:
    -div(K grad u) = 0, u = 1 on the first column, u = 0 on the last column, no flux through the top and
    bottom rows (the boundary conditions models/darcy.py:226-233 upstream penalises), harmonic-mean face
    transmissibilities, cell size 1/W x 1/H.  Returns (3, H, W) float64: pressure u and the flux
    sigma = -K grad u at the cell centres (mean of the two face fluxes; one-sided at Dirichlet columns).

    Stand-in for the FEniCS mixed-FEM solver that produced the reference's test labels
    (utils/fenics.py:13-91 upstream; dolfin is not installable here): it provides `output` fields for
    the validation files so that the script's r^2 / NRMSE columns are meaningful."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    K = np.asarray(K, dtype=np.float64)
    H, W = K.shape
    idx = np.arange(H * W).reshape(H, W)
    tx = 2.0 * K[:, :-1] * K[:, 1:] / (K[:, :-1] + K[:, 1:])   # square cells: the cell size cancels
    ty = 2.0 * K[:-1, :] * K[1:, :] / (K[:-1, :] + K[1:, :])
    rows = np.concatenate([idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()])
    cols = np.concatenate([idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()])
    vals = np.concatenate([tx.ravel(), tx.ravel(), ty.ravel(), ty.ravel()])
    A = sp.coo_matrix((-vals, (rows, cols)), shape=(H * W, H * W)).tocsr()
    A = A - sp.diags(np.asarray(A.sum(1)).ravel())
    fixed = np.zeros((H, W), dtype=bool)
    fixed[:, 0] = fixed[:, -1] = True
    ufix = np.zeros((H, W))
    ufix[:, 0] = 1.0
    f = fixed.ravel()
    rhs = -(A[~f][:, f] @ ufix.ravel()[f])
    u = ufix.ravel().copy()
    u[~f] = spla.spsolve(A[~f][:, ~f].tocsc(), rhs)
    u = u.reshape(H, W)
    fx = -tx * (u[:, 1:] - u[:, :-1]) * W           # flux through the W-1 interior x-faces
    fy = -ty * (u[1:, :] - u[:-1, :]) * H
    s1 = np.empty((H, W))
    s1[:, 1:-1] = 0.5 * (fx[:, :-1] + fx[:, 1:])
    s1[:, 0], s1[:, -1] = fx[:, 0], fx[:, -1]
    s2 = np.zeros((H, W))
    s2[1:-1, :] = 0.5 * (fy[:-1, :] + fy[1:, :])
    s2[0, :], s2[-1, :] = 0.5 * fy[0, :], 0.5 * fy[-1, :]   # the outer faces carry no flux
    return np.stack([u, s1, s2])


def darcy_fv_dataset(K):
    """(N,1,H,W) permeability -> (N,3,H,W) float32 reference fields, one sparse solve per sample."""
    K = np.asarray(K, dtype=np.float64)
    return np.stack([darcy_fv_solve(k[0]) for k in K]).astype(np.float32)


def write_script_datasets(data_dir, imsize, ntrain, ntest, kind="grf_kle512", seed=1, device=None, solve=True):
    """Write the two files train_codec_mixed_residual.py:127-139 upstream opens for `--data kind`:
    training inputs only, validation inputs + finite-volume reference outputs."""
    d = os.path.join(data_dir, "%dx%d" % (imsize, imsize))
    if kind == "grf_kle512":
        x = grf_kle(ntrain + ntest, imsize, 512, 0.1, seed=seed, device=device).numpy()
        names = ("kle512_lhs10000_train.hdf5", "kle512_lhs1000_val.hdf5")
    elif kind == "grf_kle100":   # train_cglow_reverse_kl.py:113-121 (--kle 100, 32x32)
        x = grf_kle(ntrain + ntest, imsize, 100, 0.1, seed=seed, device=device).numpy()
        names = ("kle100_lhs10000_train.hdf5", "kle100_lhs1000_val.hdf5")
    elif kind == "channelized":
        x = channelized(ntrain + ntest, imsize, seed=seed).numpy()
        names = ("channel_ng64_n4096_train.hdf5", "channel_ng64_n512_test.hdf5")
    else:
        raise ValueError("unknown dataset kind %r" % (kind,))
    xt = x[ntrain:]
    y = darcy_fv_dataset(xt) if solve else np.zeros((ntest, 3, imsize, imsize), np.float32)
    write_hdf5(os.path.join(d, names[0]), x[:ntrain])
    write_hdf5(os.path.join(d, names[1]), xt, y)
    return os.path.join(d, names[0]), os.path.join(d, names[1])
