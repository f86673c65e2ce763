"""CPU restatement of the zero-digit culling of the projection kernel (csrc/ozaki.cu: table_extent_kernel, tile_krange) and the proof
of its claim on the oracle's own numbers: every K step (32 consecutive contraction voxels) that a tile of 128 output voxels skips
multiplies ONLY zero covariance digits, so skipping it cannot change any int32 accumulator.

The covariance values come from the oracle (oracle/numpy_oracle.cov_block, pinned against the reference), the digits from the CPU
restatement of the device's digit extraction (oracle/digit_slices.py, bit-identical to csrc/ozaki.cuh, tests/test_digit_slices.py);
the K-range arithmetic below restates tile_krange line by line.  No GPU."""
import numpy as np
import pytest

from oracle import digit_slices as ds
from oracle import numpy_oracle as o


def table_digits(c, kernel, gl, S, cb, r):
    """Digits (EY, EX, EZ, S) of the stationary table of block (cb, r) on the extended offset lattice, one exponent per table
    (cov_tables_kernel + slice_table_kernel)."""
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    dy = (np.arange(2 * yN - 1) - (yN - 1)) * c.yvoxsize
    dx = (np.arange(2 * xN - 1) - (xN - 1)) * c.xvoxsize
    dz = (np.arange(2 * zN - 1) - (zN - 1)) * c.zvoxsize
    D2 = dy[:, None, None] ** 2 + dx[None, :, None] ** 2 + dz[None, None, :] ** 2
    with np.errstate(all="ignore"):
        tab = np.asarray(o.cov_block(D2, o.dedup_lengths(np.array(gl, dtype=float)), np.asarray(c.gp_coeff, dtype=float), kernel, cb, r), dtype=float)
    tab = tab * np.ones_like(D2)
    e = ds.scale_exp(np.abs(tab).max())
    return ds.balanced_digits(np.ldexp(tab, -e), S)


def extents(dig, xN, yN):
    """table_extent_kernel: largest |dy| / |dx| offset with a non-zero digit in any plane."""
    nz = (dig != 0).any(axis=(2, 3))                      # (EY, EX)
    ys, xs = np.nonzero(nz)
    if ys.size == 0:
        return 0, 0
    return int(np.abs(ys - (yN - 1)).max()), int(np.abs(xs - (xN - 1)).max())


def tile_krange(xN, yN, zN, c0, ncol, itile, ey, ex):
    """tile_krange (csrc/ozaki.cu), culling branch: (jya, nrows, koff, w, rowsteps)."""
    XZ = xN * zN
    assert XZ % 32 == 0
    g0 = c0 + itile * 128
    g1 = min(g0 + 127, c0 + ncol - 1)
    iya, iyb = g0 // XZ, g1 // XZ
    ixa, ixb = 0, xN - 1
    if iya == iyb:
        ixa, ixb = (g0 % XZ) // zN, (g1 % XZ) // zN
    jya, jyb = max(0, iya - ey), min(yN - 1, iyb + ey)
    jxa, jxb = max(0, ixa - ex), min(xN - 1, ixb + ex)
    rowsteps = XZ // 32
    koff = (jxa * zN) // 32
    w = ((jxb + 1) * zN + 31) // 32 - koff
    return jya, jyb - jya + 1, koff, w, rowsteps


@pytest.mark.parametrize("shape,kernel,gl_mult,S", [((8, 10, 16), "sparse", (1.0, 1.0, 1.0), 5), ((16, 6, 16), "exp", (1.0, 1.0, 1.0), 5),
                                                  ((4, 12, 32), "exp", (1.0, 1.0, 1.0), 4), ((6, 9, 16), "matern32", (1.0, 1.01, 1.02), 5)])
def test_every_skipped_k_step_holds_only_zero_digits(shape, kernel, gl_mult, S):
    from geobo_b200 import synth
    xN, yN, zN = shape
    cfg = synth.settings(xN, yN, zN, kernelfunc=kernel, gp_lengthscale={"sparse": 0.6, "exp": 0.6 if xN > 4 else 0.15, "matern32": 0.25}[kernel])
    c = o.make_config(cfg)
    gl = c.gp_lengthscale * c.xvoxsize * np.asarray(gl_mult)
    N = xN * yN * zN
    ksteps = N // 32
    vox = np.arange(N)
    vy, vx, vz = vox // (xN * zN), (vox // zN) % xN, vox % zN
    culled_any = False
    for cb in range(2):
        for r in range(3):
            dig = table_digits(c, kernel, gl, S, cb, r)
            ey, ex = extents(dig, xN, yN)
            nzmask = (dig != 0).any(axis=3)                                   # (EY, EX, EZ): any non-zero digit at this offset
            for c0, ncol in ((0, N), (128 * ((N // 128) // 2), N - 128 * ((N // 128) // 2))):      # whole cube and an upper shard
                for itile in range((ncol + 127) // 128):
                    jya, nrows, koff, w, rowsteps = tile_krange(xN, yN, zN, c0, ncol, itile, ey, ex)
                    kept = np.zeros(ksteps, dtype=bool)
                    for jj in range(nrows):
                        kept[(jya + jj) * rowsteps + koff:(jya + jj) * rowsteps + koff + w] = True
                    assert kept.any()
                    i = np.arange(c0 + itile * 128, min(c0 + itile * 128 + 128, c0 + ncol))
                    skipped = np.flatnonzero(~kept)
                    culled_any |= skipped.size > 0
                    if skipped.size == 0:
                        continue
                    j = (skipped[:, None] * 32 + np.arange(32)[None, :]).ravel()
                    # digits of K[(cb, j), (r, i)] = table at offset L(i) - L(j)
                    oy = vy[i][None, :] - vy[j][:, None] + (yN - 1)
                    ox = vx[i][None, :] - vx[j][:, None] + (xN - 1)
                    oz = vz[i][None, :] - vz[j][:, None] + (zN - 1)
                    assert not nzmask[oy, ox, oz].any(), (kernel, cb, r, itile)
    assert culled_any or kernel == "matern32"          # the compact and the short exp kernels really cull on these cubes
