"""Stand-in for utils/fenics.py upstream (FEniCS mixed-FEM solve of the nonlinear Darcy problem,
lines 13-91; `dolfin` is not installable here).  `solve_nonlinear_poisson` returns reference fields
(u, sigma1, sigma2) for ONE permeability field by a Picard iteration on the cell-centred finite-volume
discretisation of pde_surrogate_b200.data.darcy_fv_solve: the nonlinear law
    -K grad(u) = g(sigma),   g(s) = s + alpha1 sqrt(K) s^2 + alpha2 K s^3
is inverted face by face (Newton on the scalar g), which defines an effective face transmissibility for
the next linear solve.  The solver script only plots / saves this field next to the network's solution
(solve_conv_mixed_residual.py:103-113)."""
import numpy as np


def _invert_g(rhs, b1, b2, iters=30):
    s = rhs.copy()
    for _ in range(iters):
        g = s + b1 * s * s + b2 * s ** 3 - rhs
        dg = 1.0 + 2.0 * b1 * s + 3.0 * b2 * s * s
        s = s - g / np.where(np.abs(dg) > 1e-12, dg, 1e-12)
    return s


def solve_nonlinear_poisson(K, alpha1, alpha2, run_dir=None, iters=40, tol=1e-10):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from pde_surrogate_b200.data import darcy_fv_solve
    K = np.asarray(K, dtype=np.float64)
    H, W = K.shape
    out = darcy_fv_solve(K)
    if alpha1 == 0 and alpha2 == 0:
        return out.astype(np.float32)
    u = out[0]
    kx = 2.0 * K[:, :-1] * K[:, 1:] / (K[:, :-1] + K[:, 1:])      # face permeabilities (harmonic means)
    ky = 2.0 * K[:-1, :] * K[1:, :] / (K[:-1, :] + K[1:, :])
    idx = np.arange(H * W).reshape(H, W)
    fixed = np.zeros((H, W), dtype=bool)
    fixed[:, 0] = fixed[:, -1] = True
    ufix = np.zeros((H, W))
    ufix[:, 0] = 1.0
    f = fixed.ravel()
    for _ in range(iters):
        # flux through every face from the current pressure: sigma = g^-1(-k du/dn)
        gx = -kx * (u[:, 1:] - u[:, :-1]) * W
        gy = -ky * (u[1:, :] - u[:-1, :]) * H
        sx = _invert_g(gx, alpha1 * np.sqrt(kx), alpha2 * kx)
        sy = _invert_g(gy, alpha1 * np.sqrt(ky), alpha2 * ky)
        tx = np.where(np.abs(gx) > 1e-14, kx * sx / np.where(np.abs(gx) > 1e-14, gx, 1.0), kx)   # effective transmissibility
        ty = np.where(np.abs(gy) > 1e-14, ky * sy / np.where(np.abs(gy) > 1e-14, gy, 1.0), ky)
        rows = np.concatenate([idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()])
        cols = np.concatenate([idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()])
        vals = np.concatenate([tx.ravel(), tx.ravel(), ty.ravel(), ty.ravel()])
        A = sp.coo_matrix((-vals, (rows, cols)), shape=(H * W, H * W)).tocsr()
        A = A - sp.diags(np.asarray(A.sum(1)).ravel())
        rhs = -(A[~f][:, f] @ ufix.ravel()[f])
        un = ufix.ravel().copy()
        un[~f] = spla.spsolve(A[~f][:, ~f].tocsc(), rhs)
        un = un.reshape(H, W)
        done = np.abs(un - u).max() < tol
        u = un
        if done:
            break
    gx = -kx * (u[:, 1:] - u[:, :-1]) * W
    gy = -ky * (u[1:, :] - u[:-1, :]) * H
    fx = _invert_g(gx, alpha1 * np.sqrt(kx), alpha2 * kx)
    fy = _invert_g(gy, alpha1 * np.sqrt(ky), alpha2 * ky)
    s1 = np.empty((H, W))
    s1[:, 1:-1] = 0.5 * (fx[:, :-1] + fx[:, 1:])
    s1[:, 0], s1[:, -1] = fx[:, 0], fx[:, -1]
    s2 = np.zeros((H, W))
    s2[1:-1, :] = 0.5 * (fy[:-1, :] + fy[1:, :])
    s2[0, :], s2[-1, :] = 0.5 * fy[0, :], 0.5 * fy[-1, :]
    return np.stack([u, s1, s2]).astype(np.float32)
