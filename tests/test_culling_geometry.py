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


def tile_krange_chunk(xN, yN, zN, c0, ncol, itile, ey, ex, cy0, cy1):
    """tile_krange with the streamed-contraction clamp (rows [cy0, cy1) of one launch): None when the tile has no step in the chunk."""
    XZ = xN * zN
    g0 = c0 + itile * 128
    g1 = min(g0 + 127, c0 + ncol - 1)
    iya, iyb = g0 // XZ, g1 // XZ
    ixa, ixb = 0, xN - 1
    if iya == iyb:
        ixa, ixb = (g0 % XZ) // zN, (g1 % XZ) // zN
    jya, jyb = max(0, iya - ey), min(yN - 1, iyb + ey)
    jya, jyb = max(jya, cy0), min(jyb, cy1 - 1)
    if jyb < jya:
        return None
    jxa, jxb = max(0, ixa - ex), min(xN - 1, ixb + ex)
    koff = (jxa * zN) // 32
    return jya, jyb - jya + 1, koff, ((jxb + 1) * zN + 31) // 32 - koff, XZ // 32


@pytest.mark.parametrize("chunk_rows", [1, 2, 3, 5])
def test_streamed_chunks_partition_the_kept_steps(chunk_rows):
    """Streamed contraction: one launch per chunk of voxel rows; the steps a tile visits over all launches are exactly the steps of
    the unchunked launch, each once (the launches add into Pt)."""
    xN, yN, zN, ey, ex = 8, 11, 16, 3, 2
    N = xN * yN * zN
    for c0, ncol in ((0, N), (512, N - 512)):
        for itile in range((ncol + 127) // 128):
            jya, nrows, koff, w, rowsteps = tile_krange(xN, yN, zN, c0, ncol, itile, ey, ex)
            full = np.zeros(N // 32, dtype=int)
            for jj in range(nrows):
                full[(jya + jj) * rowsteps + koff:(jya + jj) * rowsteps + koff + w] += 1
            acc = np.zeros(N // 32, dtype=int)
            for y0 in range(0, yN, chunk_rows):
                y1 = min(yN, y0 + chunk_rows)
                kr = tile_krange_chunk(xN, yN, zN, c0, ncol, itile, ey, ex, y0, y1)
                if kr is None:
                    continue
                a, n, ko, ww, rs = kr
                ks_base, ksteps_local = y0 * rs, (y1 - y0) * rs
                for jj in range(n):
                    lo = (a + jj) * rs + ko
                    assert ks_base <= lo and lo + ww <= ks_base + ksteps_local          # inside the chunk's digit blocks
                    acc[lo:lo + ww] += 1
            assert np.array_equal(acc, full) and full.max() == 1


def test_alternating_sorted_tile_order_is_a_permutation_with_smooth_round_boundaries():
    """tile_at: position -> voxel-column tile through the order sorted by K-step count, descending for even sensor-row tiles and
    ascending for odd ones: every tile once per sensor-row tile, and consecutive positions across a sensor-row-tile boundary hold tiles
    of equal cost."""
    rng = np.random.default_rng(0)
    n_itile, n_stile = 37, 5
    keys = rng.integers(10, 60, n_itile)
    rank = np.array([np.sum((keys > keys[i]) | ((keys == keys[i]) & (np.arange(n_itile) < i))) for i in range(n_itile)])   # tile_rank_kernel
    perm = np.empty(n_itile, dtype=int)
    perm[rank] = np.arange(n_itile)
    assert sorted(perm) == list(range(n_itile)) and (np.diff(keys[perm]) <= 0).all()
    seq = []
    for rem in range(n_stile * n_itile):                       # tile_at
        stile, pos = divmod(rem, n_itile)
        if stile & 1:
            pos = n_itile - 1 - pos
        seq.append((stile, perm[pos]))
    for st in range(n_stile):
        assert sorted(t for s_, t in seq if s_ == st) == list(range(n_itile))
    cost = np.array([keys[t] for _, t in seq])
    for st in range(1, n_stile):
        assert cost[st * n_itile - 1] == cost[st * n_itile]     # the same tile cost on both sides of the boundary
