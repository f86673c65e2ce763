"""TEST INFRASTRUCTURE ONLY (imported by tests/ -- never by geobo_b200).

CPU restatement of the separable form of the squared-exponential covariance blocks.

On the voxel grid of ``kernels.calcGridPoints3D`` (``geobo/kernels.py:27-42``: coordinates ``(1..n) * voxel size``, voxel
index ``(iy * xN + ix) * zN + iz``) the squared distance of ``calcDistanceMatrix`` (``kernels.py:45-61``) is a sum of three
per-axis terms, and both exp kernels (``gpkernel`` ``kernels.py:81-88``, ``gpkernel2`` ``:90-99``) are ``coef * exp(-D2 / s)``.
Hence every block of ``create_cov`` (``kernels.py:158-195``) is

    K_rc = t0 * Ky (x) Kx (x) Kz,      K_axis[a, b] = block(D2 = ((a - b) * voxel size)^2) / t0,      t0 = block(D2 = 0)

and ``A . K_rc`` is three Toeplitz mode products.  Nothing here is a new formula: the factor lines are the oracle's own
``cov_block`` evaluated on the three coordinate axes, exactly what the device reads off its stationary tables.
Parity: pinned through ``numpy_oracle.pt_panel`` / ``predict_lean`` (themselves pinned against the reference's golden
vectors), see tests/test_kron.py.
"""
import numpy as np

from . import numpy_oracle as o


def factor_lines(c, params, w, amp, cb, r):
    """(t0, fy, fx, fz): value of block (data block cb, property block r) at zero offset and the block's values along the
    y / x / z axis for offsets -(n-1) .. n-1 (the axis lines of the device's stationary tables)."""
    vox = {"y": c.yvoxsize, "x": c.xvoxsize, "z": c.zvoxsize}
    n = {"y": c.yNcube, "x": c.xNcube, "z": c.zNcube}
    lines = []
    for ax in "yxz":
        d = np.arange(-(n[ax] - 1), n[ax]) * vox[ax]
        lines.append(amp * o.cov_block(d ** 2, params, w, "exp", cb, r) * np.ones(d.size))
    t0 = float(amp * np.asarray(o.cov_block(np.zeros(1), params, w, "exp", cb, r)).ravel()[0] * 1.0)
    return t0, lines[0], lines[1], lines[2]


def toeplitz(line, n):
    """n x n matrix T[a, b] = line[(a - b) + n - 1]."""
    a = np.arange(n)
    return line[a[:, None] - a[None, :] + n - 1]


def apply_block(c, params, w, amp, cb, r, X):
    """X (rows, N) -> X . K_(cb, r)  through the three mode products (y first, then z, then x -- the device's order)."""
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    t0, fy, fx, fz = factor_lines(c, params, w, amp, cb, r)
    if t0 == 0.0:
        return np.zeros_like(X)
    T = X.reshape(-1, yN, xN, zN)
    T = np.einsum("ab,sbxz->saxz", toeplitz(fy / t0 / t0, yN), T)
    T = np.einsum("ab,syxb->syxa", toeplitz(fz, zN), T)
    T = np.einsum("ab,sybz->syaz", toeplitz(fx, xN), T)
    return T.reshape(X.shape)


def pt_kron(c, params, w, amp, A_list, didx):
    """Pt = Asens3 . kcov as (M, 3, N), like ``numpy_oracle.pt_panel`` over all columns: the two survey blocks through the
    mode products, the drill rows as gathers of the covariance (unchanged)."""
    Ns = A_list[0].shape[0]
    N = A_list[0].shape[1]
    nd = didx.size
    out = np.zeros((2 * Ns + nd, 3, N))
    for cb in range(2):
        for r in range(3):
            out[cb * Ns:(cb + 1) * Ns, r, :] = apply_block(c, params, w, amp, cb, r, A_list[cb])
    if nd:
        pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
        out[2 * Ns:] = o.pt_panel(c, params, w, amp, A_list, didx, pts, np.arange(N))[2 * Ns:]
    return out


def kw_kron(c, params, w, amp, W):
    """z[r] = sum_cb K_(cb, r) w[cb]  for W (3, N): the covariance block matrix times one vector (refinement, mean)."""
    return np.stack([sum(apply_block(c, params, w, amp, cb, r, W[cb][None, :])[0] for cb in range(3)) for r in range(3)])
