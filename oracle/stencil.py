"""TEST INFRASTRUCTURE ONLY (imported by tests/ -- never by geobo_b200).

CPU restatement of the compact-support form of the 'sparse' covariance blocks.

The Melkumyan kernels ``gpkernel_sparse`` / ``gpkernel_sparse2`` (``geobo/kernels.py:101-138``) are zero for ``d >= gamma``
(same property) and ``d > (l1 + l2) / 2`` (cross; ``l2`` bumped by 1e-3 when the scales coincide, ``:125-126``), and on the
voxel grid of ``calcGridPoints3D`` every block of ``create_cov`` only depends on the integer offset between two voxels.  So

    (X . K_rc)[s, (jy, jx, jz)] = sum over offsets (dy, dx, dz) inside the window of  K_rc(offset) * X[s, (jy-dy, jx-dx, jz-dz)]

Nothing here is a new formula: the taps are the oracle's own ``cov_block`` evaluated on the offset lattice.  Parity: pinned
through ``numpy_oracle.pt_panel`` / ``create_cov`` (see tests/test_compact.py).
"""
import numpy as np

from . import numpy_oracle as o


def window(c, params):
    """Half widths (ry, rx, rz) of the tap window: offsets k with k * voxel size within 1.001 * the largest length scale."""
    r = 1.001 * float(np.max(params))
    out = []
    for vox, n in ((c.yvoxsize, c.yNcube), (c.xvoxsize, c.xNcube), (c.zvoxsize, c.zNcube)):
        out.append(int(min(n - 1, np.floor(r / vox + 1e-9))))
    return tuple(out)


def taps(c, params, w, amp, cb, r):
    """Block (cb, r) on the offset lattice of the window: array (2ry+1, 2rx+1, 2rz+1), entry [dy+ry, dx+rx, dz+rz]."""
    ry, rx, rz = window(c, params)
    dy, dx, dz = np.meshgrid(np.arange(-ry, ry + 1) * c.yvoxsize, np.arange(-rx, rx + 1) * c.xvoxsize, np.arange(-rz, rz + 1) * c.zvoxsize,
                             indexing="ij")
    D2 = dx ** 2 + dy ** 2 + dz ** 2                   # summed x, y, z like kernels.py:46,54-58
    return amp * o.cov_block(D2, params, w, "sparse", cb, r) * np.ones(D2.shape)


def apply_block(c, params, w, amp, cb, r, X):
    """X (rows, N) -> X . K_(cb, r)  as the tap sum."""
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    ry, rx, rz = window(c, params)
    T = taps(c, params, w, amp, cb, r)
    src = np.asarray(X, dtype=float).reshape(-1, yN, xN, zN)
    pad = np.zeros((src.shape[0], yN + 2 * ry, xN + 2 * rx, zN + 2 * rz))
    pad[:, ry:ry + yN, rx:rx + xN, rz:rz + zN] = src
    out = np.zeros_like(src)
    for a in range(2 * ry + 1):
        for b in range(2 * rx + 1):
            for d in range(2 * rz + 1):
                t = T[a, b, d]
                if t != 0.0:
                    # offset (dy, dx, dz) = (a - ry, b - rx, d - rz); input voxel = output voxel - offset
                    out += t * pad[:, 2 * ry - a:2 * ry - a + yN, 2 * rx - b:2 * rx - b + xN, 2 * rz - d:2 * rz - d + zN]
    return out.reshape(np.asarray(X).shape)


def pt_compact(c, params, w, amp, A_list, didx):
    """Pt = Asens3 . kcov as (M, 3, N), like ``numpy_oracle.pt_panel`` over all columns."""
    Ns, N = A_list[0].shape
    nd = didx.size
    out = np.zeros((2 * Ns + nd, 3, N))
    for cb in range(2):
        for r in range(3):
            out[cb * Ns:(cb + 1) * Ns, r, :] = apply_block(c, params, w, amp, cb, r, A_list[cb])
    if nd:
        pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
        out[2 * Ns:] = o.pt_panel(c, params, w, amp, A_list, didx, pts, np.arange(N))[2 * Ns:]
    return out


def kw_compact(c, params, w, amp, W):
    """z[r] = sum_cb K_(cb, r) w[cb]  for W (3, N)."""
    return np.stack([sum(apply_block(c, params, w, amp, cb, r, W[cb][None, :])[0] for cb in range(3)) for r in range(3)])
