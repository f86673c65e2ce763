"""TEST INFRASTRUCTURE ONLY (imported by tests/ -- never by geobo_b200).

CPU restatement of the block-Toeplitz (FFT) form of the covariance blocks, any kernel family.

On the voxel grid of ``kernels.calcGridPoints3D`` (``geobo/kernels.py:27-42``) every block of ``create_cov``
(``kernels.py:158-195``) only depends on the integer offset between two voxels, so ``X . K_rc`` is a 3-D linear convolution
of every row with the block's values on the offset lattice -- evaluated here by circulant embedding with ``numpy.fft``: pad
each axis to a power of two ``P >= 2n - 1``, wrap the offsets into the padded lattice, multiply the transforms, keep the
first ``n`` outputs per axis.  Nothing here is a new formula: the offset values are the oracle's own ``cov_block``.
Parity: pinned through ``numpy_oracle.pt_panel`` / ``create_cov`` (see tests/test_fftconv.py).
"""
import numpy as np

from . import numpy_oracle as o


def padded(c):
    """(Py, Px, Pz): powers of two >= 2n - 1."""
    return tuple(int(2 ** np.ceil(np.log2(max(1, 2 * n - 1)))) for n in (c.yNcube, c.xNcube, c.zNcube))


def spectrum(c, params, w, amp, kernel, cb, r):
    """Real spectrum of block (cb, r) wrapped into the padded lattice."""
    ns = (c.yNcube, c.xNcube, c.zNcube)
    vox = (c.yvoxsize, c.xvoxsize, c.zvoxsize)
    P = padded(c)
    offs = [np.arange(-(n - 1), n) for n in ns]
    dy, dx, dz = np.meshgrid(offs[0] * vox[0], offs[1] * vox[1], offs[2] * vox[2], indexing="ij")
    with np.errstate(all="ignore"):
        T = amp * o.cov_block(dx ** 2 + dy ** 2 + dz ** 2, params, w, kernel, cb, r) * np.ones(dx.shape)
    wrapped = np.zeros(P)
    wrapped[np.ix_(offs[0] % P[0], offs[1] % P[1], offs[2] % P[2])] = T
    S = np.fft.fftn(wrapped)
    assert np.abs(S.imag).max() <= 1e-9 * max(1.0, np.abs(S.real).max())      # even table -> real spectrum
    return S.real


def apply_block(c, params, w, amp, kernel, cb, r, X):
    """X (rows, N) -> X . K_(cb, r)."""
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    P = padded(c)
    S = spectrum(c, params, w, amp, kernel, cb, r)
    src = np.asarray(X, dtype=float).reshape(-1, yN, xN, zN)
    pad = np.zeros((src.shape[0],) + P)
    pad[:, :yN, :xN, :zN] = src
    out = np.fft.ifftn(np.fft.fftn(pad, axes=(1, 2, 3)) * S, axes=(1, 2, 3)).real[:, :yN, :xN, :zN]
    return out.reshape(np.asarray(X).shape)


def pt_fft(c, params, w, amp, kernel, A_list, didx):
    """Pt = Asens3 . kcov as (M, 3, N), like ``numpy_oracle.pt_panel`` over all columns."""
    Ns, N = A_list[0].shape
    nd = didx.size
    out = np.zeros((2 * Ns + nd, 3, N))
    for cb in range(2):
        for r in range(3):
            out[cb * Ns:(cb + 1) * Ns, r, :] = apply_block(c, params, w, amp, kernel, cb, r, A_list[cb])
    if nd:
        pts = o.grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
        out[2 * Ns:] = o.pt_panel(c, params, w, amp, A_list, didx, pts, np.arange(N))[2 * Ns:]
    return out


def kw_fft(c, params, w, amp, kernel, W):
    """z[r] = sum_cb K_(cb, r) w[cb]  for W (3, N)."""
    return np.stack([sum(apply_block(c, params, w, amp, kernel, cb, r, W[cb][None, :])[0] for cb in range(3)) for r in range(3)])
