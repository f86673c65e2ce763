"""NumPy/SciPy restatement of GeoBO's joint-inversion hot path.  TEST INFRASTRUCTURE ONLY.

See ``oracle/__init__.py`` for who may import this.  Every function cites the
reference lines it restates (paths relative to the reference repo).  Two paths:

* ``cubing_literal`` -- the algorithm exactly as the reference runs it: dense
  ``D2`` (N x N), dense ``kcov`` (3N x 3N), dense zero-padded ``Asens3`` (M x 3N),
  full posterior covariance.  Needs ~30 N^2 x 8 B; small cubes only.
* ``cubing_lean``    -- the same arithmetic with the structure exploited
  (``Asens3`` is block-diagonal, ``kcov`` symmetric, only diag(cov) is used):
  ``Pt = Asens3 . kcov`` assembled block by block in row panels, ``AkA = Asens3 . Pt^T``,
  ``var = amp - colsumsq(L^-1 Pt)``.  Scales to the benchmark cubes; this is the
  CPU baseline ``bench.py`` times.

Parity status: pinned against the reference's committed result cubes and live
reference runs (``tests/test_oracle_golden.py``).
"""
import json
import math
import time

import numpy as np
from scipy.linalg import cholesky, solve_triangular

ALONG_WAY = 1e6  # sensormodel.py:64 ("aLongWay")


# --------------------------------------------------------------------------- config
class Config(dict):
    """Settings + derived constants, as ``geobo/config_loader.py:33-59`` injects them as globals."""

    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def make_config(cfg):
    """``cfg``: dict of YAML keys (or JSON string).  Adds the derived names of config_loader.py:41-59."""
    if isinstance(cfg, (str, bytes)):
        cfg = json.loads(cfg)
    c = Config(cfg)
    c.xLcube = c.xmax - c.xmin                         # config_loader.py:41
    c.yLcube = c.ymax - c.ymin                         # :42
    c.zmin = c.zmax - c.zLcube                         # :44
    c.magneticField = np.asarray([c.XMAG, c.YMAG, c.ZMAG]) * 1e-3   # :46
    c.c_MILLIGALS_UNITS = c.c_G * c.c_SI_TO_MILLIGALS * c.c_GCM3_TO_SI  # :53
    c.xvoxsize = c.xLcube / c.xNcube * 1.0             # :56-58
    c.yvoxsize = c.yLcube / c.yNcube * 1.0
    c.zvoxsize = c.zLcube / c.zNcube * 1.0
    c.Nsensor = c.xNcube * c.yNcube                    # :59
    return c


def load_yaml(path, **overrides):
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f)
    cfg.update(overrides)
    return make_config(cfg)


# --------------------------------------------------------------------------- kernels.py
def grid_points(Lpix, pixscale):
    """kernels.py:27-42 -- points ((ix+1)sx, (iy+1)sy, (iz+1)sz), row = (iy*xN+ix)*zN+iz."""
    xr = np.arange(1, Lpix[0] + 1) * pixscale[0]
    yr = np.arange(1, Lpix[1] + 1) * pixscale[1]
    zr = np.arange(1, Lpix[2] + 1) * pixscale[2]
    xg, yg, zg = np.meshgrid(xr, yr, zr)  # default 'xy' indexing -> shape (yN, xN, zN)
    return np.stack([xg.ravel(), yg.ravel(), zg.ravel()], axis=1)


def sqdist(points, rows=None):
    """kernels.py:45-61 -- D2[i,j] = sum_d (p[j,d]-p[i,d])^2, summed in dimension order starting from 0."""
    p = np.asarray(points, dtype=float)
    q = p if rows is None else p[rows]
    acc = 0  # np.sum(generator) == builtin sum: starts from int 0, adds d = 0, 1, 2 in order
    for d in range(p.shape[1]):
        delta = p[:, d][None, :] - q[:, d][:, None]
        acc = acc + delta ** 2
    return acc


def k_exp(D2, g):                       # kernels.py:81-88
    return np.exp(-0.5 * D2 / g ** 2)


def k_exp2(D2, l1, l2):                 # kernels.py:90-99
    return np.sqrt(2.0 * l1 * l2 / (l1 ** 2 + l2 ** 2)) * np.exp(-D2 / (l1 ** 2 + l2 ** 2))


def k_sparse(D2, g):                    # kernels.py:101-114
    d = np.sqrt(D2)
    res = np.zeros_like(d)
    m = d < g
    dm = d[m]
    res[m] = (2 + np.cos(2 * np.pi * dm / g)) / 3.0 * (1 - dm / g) + 1 / (2.0 * np.pi) * np.sin(2 * np.pi * dm / g)
    res[res < 0.0] = 0.0
    return res


def k_sparse2(D2, l1, l2):              # kernels.py:116-138 (Q3: cosine sits inside the sine in branch A)
    d = np.sqrt(D2)
    if l1 == l2:
        l2 = l2 + 1e-3 * l2
    lmean = np.mean([l1, l2])
    lmin = np.min([l1, l2])
    lmax = np.max([l1, l2])
    res = np.zeros_like(d)
    mA = d <= abs(l2 - l1) / 2.0
    res[mA] = 2.0 / (3 * np.sqrt(l1 * l2)) * (
        lmin + 1 / np.pi * lmax ** 3 / (lmax ** 2 - lmin ** 2)
        * np.sin(np.pi * lmin / lmax * np.cos(2 * np.pi * d[mA] / lmax)))
    mB = (d >= abs(l2 - l1) / 2.0) & (d <= (l1 + l2) / 2.0)   # assigned second: wins ties with A
    dB = d[mB]
    res[mB] = 2.0 / (3 * np.sqrt(l1 * l2)) * (
        lmean - dB
        + l1 ** 3 * np.sin(np.pi * (l2 - 2.0 * dB) / l1) / (2 * np.pi * (l1 ** 2 - l2 ** 2))
        - l2 ** 3 * np.sin(np.pi * (l1 - 2.0 * dB) / l2) / (2 * np.pi * (l1 ** 2 - l2 ** 2)))
    res[res < 0.0] = 0.0
    return res


def k_matern32(D2, g):                  # kernels.py:140-146
    nu = np.sqrt(3) * np.sqrt(D2) / g
    return (1 + nu) * np.exp(-nu)


def k_matern32_2(D2, l1, l2):           # kernels.py:148-156 (singular for l1 == l2)
    norm = 2 * np.sqrt(l1 * l2) / (l1 ** 2 - l2 ** 2)
    return norm * (l1 * np.exp(-np.sqrt(3 * D2) / l1) - l2 * np.exp(-np.sqrt(3 * D2) / l2))


_SAME = {"exp": k_exp, "sparse": k_sparse, "matern32": k_matern32}
_CROSS = {"exp": k_exp2, "sparse": k_sparse2, "matern32": k_matern32_2}


def dedup_lengths(params):
    """kernels.py:174-180, *as coded* (Q1): operates in place on an ndarray; the second test
    rewrites element 1, not 2, so [L,L,L] -> [L, 1.02 L, L]."""
    if params[1] == params[0]:
        params[1] = 1.01 * params[0]
    if params[2] == params[0]:
        params[1] = 1.02 * params[0]
    if params[2] == params[1]:
        params[2] = 1.01 * params[1]
    return params


def cross_weight(w, r, c):
    """kernels.py:181-194: w1 density-drill (0,2), w2 magsus-drill (1,2), w3 density-magsus (0,1)."""
    w1, w2, w3 = w
    return {(0, 1): w3, (0, 2): w1, (1, 2): w2}[(min(r, c), max(r, c))]


def cov_block(D2, params, w, fkernel, r, c):
    """Block at row-block r, column-block c of kernels.py:183-195 (column strip c, vstack slot r):
    same-property kernel on the diagonal, else w * cross(params[[c, r]])."""
    if r == c:
        return _SAME[fkernel](D2, params[c])
    return cross_weight(w, r, c) * _CROSS[fkernel](D2, params[c], params[r])


def create_cov(D2, gplength, crossweights=(1, 1, 1), fkernel="sparse"):
    """kernels.py:158-195.  ``np.asarray`` aliases an ndarray argument, so the de-dup mutates the caller's array."""
    params = dedup_lengths(np.asarray(gplength))
    w = np.asarray(crossweights)
    strips = [np.vstack([cov_block(D2, params, w, fkernel, r, c) for r in range(3)]) for c in range(3)]
    return np.hstack(strips)


# --------------------------------------------------------------------------- inversion.py geometry
def cube_geometry(c):
    """inversion.py:54-74: edge lattice (3, yN+1, xN+1, zN+1) with depth-positive z; voxel centres (3, N)."""
    xedge = np.linspace(0, c.xNcube, c.xNcube + 1) * c.xvoxsize
    yedge = np.linspace(0, c.yNcube, c.yNcube + 1) * c.yvoxsize
    zedge = np.linspace(0, -c.zNcube, c.zNcube + 1) * c.zvoxsize + c.zmax
    xE, yE, zE = np.meshgrid(xedge, yedge, zedge)
    Edges = np.asarray([xE, yE, -zE])
    xnew = np.arange(c.xvoxsize / 2.0, c.xLcube + c.xvoxsize / 2.0, c.xvoxsize)
    ynew = np.arange(c.yvoxsize / 2.0, c.yLcube + c.yvoxsize / 2.0, c.yvoxsize)
    znew = c.zmax - np.arange(c.zvoxsize / 2.0, c.zLcube + c.zvoxsize / 2.0, c.zvoxsize)
    xxx, yyy, zzz = np.meshgrid(xnew, ynew, znew)
    voxelpos = np.vstack([xxx.flatten(), yyy.flatten(), zzz.flatten()])
    return Edges, voxelpos


def sensor_grid(c):
    """run_geobo.py:61-65: sensors over voxel-column centres at height zmax+zoff, row n = iy*xN+ix."""
    x_s = np.linspace(0.5, c.xNcube - 0.5, c.xNcube) * c.xvoxsize
    y_s = np.linspace(0.5, c.yNcube - 0.5, c.yNcube) * c.yvoxsize
    z_s = c.zmax + c.zoff
    xs, ys, zs = np.meshgrid(x_s, y_s, z_s)
    return np.asarray([xs.flatten(), ys.flatten(), zs.flatten()]).T


# --------------------------------------------------------------------------- sensormodel.py
def grav_corner(x, y, z):               # sensormodel.py:96-110
    eps = 1e-9
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    return x * np.log(y + r) + y * np.log(x + r) - z * np.arctan((x * y) / (z * r + eps))


def magn_corner(x, y, z, bx, by, bz):   # sensormodel.py:113-133
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    normB = np.sqrt(bx * bx + by * by + bz * bz)
    f = 1.0 / normB * ((2.0 * by * bz * np.log(x + r)) + (2.0 * bz * bx * np.log(y + r)) + (2.0 * by * bx * np.log(z + r))
                       + (bz * bz - by * by) * np.arctan((x * z) / (y * r))
                       + (bz * bz - bx * bx) * np.arctan((y * z) / (x * r)))
    return -f


def a_sens(c, magneticField, locations, Edges, func, sensors=None):
    """sensormodel.py:29-93, vectorised over the voxel loops with the loop's own association order
    (so it is bit-identical to the triple loop); one sensor per outer iteration like the reference.
    ``sensors``: optional subset of sensor rows (for sampled CPU timing)."""
    Edges = np.asarray(Edges)
    xE, yE, zE = Edges[0], Edges[1], Edges[2]
    bx, by, bz = magneticField[0], magneticField[1], magneticField[2]
    nsens = c.xNcube * c.yNcube               # hard-wired sensor count, sensormodel.py:54,58 (Q7)
    rows = range(nsens) if sensors is None else sensors
    sens = np.zeros((len(rows), c.xNcube * c.yNcube * c.zNcube))
    for out_row, n in enumerate(rows):
        x0 = xE - locations[n, 0]
        y0 = yE - locations[n, 1]
        z0 = zE - locations[n, 2]
        x0[0] -= ALONG_WAY                    # :64-68 -- axis 0 is the y-edge index (Q4)
        y0[0] -= ALONG_WAY
        x0[-1] += ALONG_WAY
        y0[-1] += ALONG_WAY
        with np.errstate(all="ignore"):
            eZ = grav_corner(x0, y0, z0) if func == "grav" else magn_corner(x0, y0, z0, bx, by, bz)
        hi, lo = slice(1, None), slice(None, -1)
        s = -((eZ[hi, hi, hi] - eZ[hi, hi, lo] - eZ[hi, lo, hi] + eZ[hi, lo, lo])
              - (eZ[lo, hi, hi] - eZ[lo, hi, lo] - eZ[lo, lo, hi] + eZ[lo, lo, lo]))   # :83-84
        sens[out_row] = s.reshape(-1)
    if func == "grav":
        sens = c.c_MILLIGALS_UNITS * sens / c.fcor_grav   # :88-89
    else:
        sens = sens / c.fcor_mag                          # :90-91
    return sens


def a_drill(loc, voxelpos):
    """sensormodel.py:136-153: one-hot rows by exact coordinate equality."""
    x, y, z = voxelpos[0].flatten(), voxelpos[1].flatten(), voxelpos[2].flatten()
    sens = np.zeros((loc.shape[1], x.size))
    for i in range(loc.shape[1]):
        sel = np.where((x == loc[0, i]) & (y == loc[1, i]) & (z == loc[2, i]))
        sens[i, sel] = 1
    return sens


# --------------------------------------------------------------------------- inversion.py
def _normalise(c, grav, mag, drillfield):
    """inversion.py:208-214 (population std)."""
    with np.errstate(all="ignore"):
        gm, gs = grav.mean(), grav.std()
        mm, ms = mag.mean(), mag.std()
        dm, ds = (drillfield.mean(), drillfield.std()) if drillfield.size else (np.nan, np.nan)
        y = np.hstack(((grav - gm) / gs, (mag - mm) / ms, (drillfield - dm) / ds))
    return y, (gs, ms, ds)


def _gp_setup(c, gp_length=None):
    """inversion.py:46-51 (Q6: xvoxsize for all three properties)."""
    gl = c.gp_lengthscale * np.asarray([c.xvoxsize, c.xvoxsize, c.xvoxsize]) if gp_length is None \
        else np.array(gp_length, dtype=float)
    return gl, np.asarray(c.gp_err, dtype=float), np.asarray(c.gp_coeff, dtype=float), 1.0


def _finish(c, mu, var, stds):
    """inversion.py:237-248."""
    gs, ms, ds = stds
    shp = (3, c.yNcube, c.xNcube, c.zNcube)
    rec, rv = mu.reshape(shp), var.reshape(shp)
    return (rec[0] * gs, rec[1] * ms, rec[2] * ds, rv[0] * gs ** 2, rv[1] * ms ** 2, rv[2] * ds ** 2)


def cubing_literal(c, grav, mag, drillfield, sensor_locations, drilldata0, gp_length=None):
    """inversion.py:182-248 + predict3 :77-122, dense and literal.  Returns (six cubes, extras)."""
    Edges, voxelpos = cube_geometry(c)
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    xxx, yyy, zzz = (v.reshape(xN, yN, zN) for v in voxelpos)            # run_geobo.py:399-403
    gl, sig, w, amp = _gp_setup(c, gp_length)
    y, stds = _normalise(c, grav, mag, drillfield)
    pts = grid_points((xN, yN, zN), (c.xvoxsize, c.yvoxsize, c.zvoxsize))  # :216
    D2 = sqdist(pts)                                                       # :217
    mask = drilldata0 != 0
    vd = np.vstack([xxx[mask], yyy[mask], zzz[mask]])                      # :219
    Ag = a_sens(c, c.magneticField * 0.0, sensor_locations, Edges, "grav")  # :223
    Am = a_sens(c, c.magneticField, sensor_locations, Edges, "magn")        # :224
    Ad = a_drill(vd, voxelpos)                                              # :225
    Z = np.zeros_like
    A3 = np.hstack([np.vstack([Ag, Z(Am), Z(Ad)]), np.vstack([Z(Ag), Am, Z(Ad)]), np.vstack([Z(Ag), Z(Am), Ad])])
    kcov = amp * create_cov(D2, gl, w, fkernel=c.kernelfunc)               # :92 (mutates gl)
    yerr = np.hstack((grav * 0.0 + sig[0], mag * 0.0 + sig[1], drillfield * 0.0 + sig[2]))
    AkA = A3 @ (kcov @ A3.T) + np.diag(yerr ** 2)                          # :96
    L = cholesky(AkA, lower=True)                                          # :100
    u = solve_triangular(L, y, lower=True)                                 # :105
    logl = -0.5 * (u @ u + np.log(np.diag(L) ** 2).sum() + xN * yN * zN * np.log(2 * np.pi))  # :108-110 (Q5)
    V = solve_triangular(L, A3 @ kcov, lower=True)                         # :114
    mu = V.T @ u                                                           # :115
    cov = kcov - V.T @ V                                                   # :117
    var = np.diag(cov)
    extras = dict(logl=logl, gl_after=gl, mu=mu, var=var, y=y, A3=A3, kcov=kcov, AkA=AkA, V=V, stds=stds)
    return _finish(c, mu, var, stds), extras


def drill_indices(drilldata0):
    """Flat voxel indices selected by A_drill given how inversion.py:219 builds ``voxelpos_drill``:
    boolean mask of ``drilldata0 != 0`` in flat (C) order -- each picks exactly its own voxel."""
    return np.flatnonzero(np.asarray(drilldata0).ravel() != 0)


def pt_panel(c, params, w, amp, A_list, didx, pts, cols, jchunk=2048, timers=None, jstarts=None):
    """Columns ``cols`` (voxel indices i) of Pt = Asens3 . kcov for all three property blocks r.
    ``jstarts``: optional subset of contraction chunks (bounded timing samples only; result then partial).

    Returns array (M, 3, len(cols)):  Pt[(cb, s), r, i] = sum_j A_cb[s, j] * K[(cb, j), (r, i)].
    """
    Ns = A_list[0].shape[0]
    nd = didx.size
    M = 2 * Ns + nd
    ncol = len(cols)
    out = np.zeros((M, 3, ncol))
    N = pts.shape[0]
    fk = c.kernelfunc
    for cb in range(2):
        A = A_list[cb]
        for j0 in (range(0, N, jchunk) if jstarts is None else jstarts):
            J = np.arange(j0, min(N, j0 + jchunk))
            t0 = time.perf_counter()
            # D2[j, i] for j in J (rows), i in cols: same arithmetic as sqdist()
            acc = 0
            for d in range(3):
                delta = pts[cols, d][None, :] - pts[J, d][:, None]
                acc = acc + delta ** 2
            blocks = [amp * cov_block(acc, params, w, fk, cb, r) for r in range(3)]  # K[(cb,J),(r,cols)] (symmetric)
            t1 = time.perf_counter()
            for r in range(3):
                out[cb * Ns:(cb + 1) * Ns, r, :] += A[:, J] @ blocks[r]
            t2 = time.perf_counter()
            if timers is not None:
                timers["kernel_eval"] = timers.get("kernel_eval", 0.0) + (t1 - t0)
                timers["dgemm_proj"] = timers.get("dgemm_proj", 0.0) + (t2 - t1)
    if nd:
        acc = 0
        for d in range(3):
            delta = pts[cols, d][None, :] - pts[didx, d][:, None]
            acc = acc + delta ** 2
        for r in range(3):
            out[2 * Ns:, r, :] = amp * cov_block(acc, params, w, fk, 2, r)
    return out


def predict_lean(c, A_list, didx, y, gl, sig, w, amp, panel=4096, timers=None):
    """predict3 (inversion.py:77-122) with block structure; returns mu (3N), var (3N), logl, extras."""
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    N = xN * yN * zN
    Ns = A_list[0].shape[0]
    nd = didx.size
    M = 2 * Ns + nd
    params = dedup_lengths(gl)                         # mutates like kernels.py:174-180
    pts = grid_points((xN, yN, zN), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    Pt = np.empty((M, 3, N))
    for i0 in range(0, N, panel):
        cols = np.arange(i0, min(N, i0 + panel))
        Pt[:, :, i0:i0 + len(cols)] = pt_panel(c, params, w, amp, A_list, didx, pts, cols, timers=timers)
    t0 = time.perf_counter()
    AkA = np.empty((M, M))
    AkA[:Ns] = A_list[0] @ Pt[:, 0, :].T               # rows of Asens3 in block 0 only touch property 0
    AkA[Ns:2 * Ns] = A_list[1] @ Pt[:, 1, :].T
    if nd:
        AkA[2 * Ns:] = Pt[:, 2, didx].T
    yerr2 = np.hstack((np.full(Ns, sig[0] ** 2), np.full(Ns, sig[1] ** 2), np.full(nd, sig[2] ** 2)))
    AkA[np.diag_indices(M)] += yerr2
    t1 = time.perf_counter()
    L = cholesky(AkA, lower=True)
    t2 = time.perf_counter()
    u = solve_triangular(L, y, lower=True)
    logl = -0.5 * (u @ u + np.log(np.diag(L) ** 2).sum() + N * np.log(2 * np.pi))
    Pt2 = Pt.reshape(M, 3 * N)
    V = solve_triangular(L, Pt2, lower=True, overwrite_b=True, check_finite=False)
    t3 = time.perf_counter()
    mu = V.T @ u
    var = amp - np.einsum("ij,ij->j", V, V)            # diag(kcov) = amp (Q10)
    t4 = time.perf_counter()
    if timers is not None:
        timers.update(aka=t1 - t0, chol=t2 - t1, trsm=t3 - t2, mean_var=t4 - t3)
    return mu, var, logl, dict(AkA_chol=L, u=u)


def cubing_lean(c, grav, mag, drillfield, sensor_locations, drilldata0, gp_length=None, timers=None):
    """Same result as ``cubing_literal`` (to rounding) without any N x N or 3N x 3N array."""
    Edges, voxelpos = cube_geometry(c)
    gl, sig, w, amp = _gp_setup(c, gp_length)
    y, stds = _normalise(c, grav, mag, drillfield)
    t0 = time.perf_counter()
    Ag = a_sens(c, c.magneticField * 0.0, sensor_locations, Edges, "grav")
    Am = a_sens(c, c.magneticField, sensor_locations, Edges, "magn")
    if timers is not None:
        timers["a_sens"] = time.perf_counter() - t0
    didx = drill_indices(drilldata0)
    mu, var, logl, ex = predict_lean(c, [Ag, Am], didx, y, gl, sig, w, amp, timers=timers)
    extras = dict(logl=logl, gl_after=gl, mu=mu, var=var, y=y, stds=stds, A_grav=Ag, A_magn=Am, didx=didx)
    return _finish(c, mu, var, stds), extras


def calc_logl(c, A_list, didx, y, params5):
    """inversion.py:125-152: negative log marginal likelihood for [amp, length-multiplier, w1, w2, w3]
    (no N log 2pi term; any failure -> +inf)."""
    try:
        amp = params5[0]
        gl = params5[1] * np.asarray([c.xvoxsize, c.xvoxsize, c.xvoxsize])
        w = np.asarray(params5[2:])
        sig = np.asarray(c.gp_err, dtype=float)
        N = c.xNcube * c.yNcube * c.zNcube
        Ns = A_list[0].shape[0]
        nd = didx.size
        M = 2 * Ns + nd
        params = dedup_lengths(gl)
        pts = grid_points((c.xNcube, c.yNcube, c.zNcube), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
        Pt = pt_panel(c, params, w, amp, A_list, didx, pts, np.arange(N))
        AkA = np.empty((M, M))
        AkA[:Ns] = A_list[0] @ Pt[:, 0, :].T
        AkA[Ns:2 * Ns] = A_list[1] @ Pt[:, 1, :].T
        if nd:
            AkA[2 * Ns:] = Pt[:, 2, didx].T
        AkA[np.diag_indices(M)] += np.hstack((np.full(Ns, sig[0] ** 2), np.full(Ns, sig[1] ** 2), np.full(nd, sig[2] ** 2)))
        L = cholesky(AkA, lower=True)
        u = solve_triangular(L, y, lower=True)
        logl = -0.5 * (u @ u + np.log(np.diag(L) ** 2).sum())
        if not np.isfinite(logl):
            raise FloatingPointError
    except Exception:
        logl = -np.inf
    return -logl


# --------------------------------------------------------------------------- bounded CPU baseline (bench.py only)
class TiledSens:
    """Stand-in for a sensitivity matrix in TIMING runs only: ``nrows`` rows made by repeating a few real rows of ``A_sens``
    (each scaled by a row factor), materialised chunk by chunk -- the timed stages use it only as a dgemm operand, whose cost
    does not depend on the values, and a dense (Ns, N) array would be 4.3 GB per survey at 64x64x32."""

    def __init__(self, rows, nrows, seed=0):
        self.rows = np.ascontiguousarray(rows)
        self.shape = (int(nrows), self.rows.shape[1])
        self.scale = 1.0 + 0.01 * np.random.default_rng(seed).standard_normal((self.shape[0], 1))

    def __getitem__(self, key):
        rsel, csel = key
        assert isinstance(rsel, slice) and rsel == slice(None)
        sub = self.rows[:, csel]
        reps = -(-self.shape[0] // sub.shape[0])
        return np.tile(sub, (reps, 1))[:self.shape[0]] * self.scale


def _timing_operands(cfg, gp_length):
    c = make_config(cfg)
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    Ns = xN * yN
    E, _ = cube_geometry(c)
    loc = sensor_grid(c)
    rows = np.unique(np.linspace(0, Ns - 1, min(Ns, 8)).astype(int))
    Ag = TiledSens(a_sens(c, c.magneticField * 0, loc, E, "grav", sensors=list(rows)), Ns, seed=1)
    Am = TiledSens(a_sens(c, c.magneticField, loc, E, "magn", sensors=list(rows)), Ns, seed=2)
    gl, sig, w, amp = _gp_setup(c, None if gp_length is None else np.array(gp_length, dtype=float))
    params = dedup_lengths(gl)
    pts = grid_points((xN, yN, zN), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    return c, [Ag, Am], params, sig, w, amp, pts


_WORKER_CACHE = {}


def json_key(cfg):
    import json
    return json.dumps(cfg, sort_keys=True, default=str)


def _pair_worker(job):
    """One worker process of the CPU timing sample: ``n`` (panel, chunk) pairs of the projection with ONE BLAS thread, like the
    workers of tests/golden/make_fullsize_golden.py.  Returns its per-pair times and kernel-evaluation / dgemm split."""
    cfg, gp_length, didx, cols, jchunk, jstarts, warm = job
    from threadpoolctl import threadpool_limits
    with threadpool_limits(1):
        key = (json_key(cfg), None if gp_length is None else tuple(gp_length))
        if _WORKER_CACHE.get("key") != key:                   # a pool that serves several steps builds its operands once
            _WORKER_CACHE.clear()
            _WORKER_CACHE.update(key=key, ops=_timing_operands(cfg, gp_length))
        c, A_list, params, sig, w, amp, pts = _WORKER_CACHE["ops"]
        didx = np.asarray(didx, dtype=int)
        if warm:
            pt_panel(c, params, w, amp, A_list, didx, pts, cols, jchunk=jchunk, jstarts=jstarts[:1])     # warm-up pair, not timed
        timers, times = {}, []
        for j0 in jstarts:
            t0 = time.perf_counter()
            pt_panel(c, params, w, amp, A_list, didx, pts, cols, jchunk=jchunk, timers=timers, jstarts=[j0])
            times.append(time.perf_counter() - t0)
    return times, timers.get("kernel_eval", 0.0), timers.get("dgemm_proj", 0.0)


def cpu_baseline_sample(cfg, didx, y, gp_length=None, panel_cols=2048, jchunk=2048, target_seconds=20.0, workers=None, pool=None,
                        probe_seconds=None, state=None):
    """Time the lean CPU path of predict3 on a bounded, deterministic sample and scale to the whole cube.

    Arrangement = the one that produced the full-size fixtures (tests/golden/make_fullsize_golden.py): the projection Pt = A.K is
    split into independent (column panel, contraction chunk) pairs handed to ``workers`` processes with one BLAS thread each, so
    EVERY host core is busy in the single-threaded NumPy ufunc stage (kernel evaluation, 2/3 of the time) as well as in dgemm --
    more favourable to the CPU than the reference's own single process, where only BLAS is threaded.  All pairs cost the same:
    each worker times ``n`` pairs one by one; the aggregate rate (sum over workers of pairs / elapsed) is scaled to the
    pair count of the cube (median / min / max pair time = the spread of the sample).  AkA, the triangular solve and mean /
    variance are linear in the column count: timed on one panel with all BLAS threads and scaled by N / panel_cols.  The M x M
    Cholesky is timed in full on an SPD matrix of the true size.  When the whole inversion fits ``target_seconds`` it is run in
    full instead (single process, BLAS on all threads; ``sample`` says "full").
    ``pool`` / ``probe_seconds`` / ``state``: a worker pool, the probe result and a dict for the untimed preparations (operands,
    the panel the small stages are timed on, "workers are warm") of an earlier call -- a caller that times many steps keeps them,
    so that process start-up, the probe and the preparations are paid once and a step is the timed sample only.
    Returns the estimated whole-cube seconds, the per-stage split and the sample description."""
    import multiprocessing as mp
    state = {} if state is None else state
    if "ops" not in state:
        state["ops"] = _timing_operands(cfg, gp_length)
    c, A_list, params, sig, w, amp, pts = state["ops"]
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    N = xN * yN * zN
    Ns = xN * yN
    didx = np.asarray(didx, dtype=int)
    nd = didx.size
    M = 2 * Ns + nd
    workers = int(workers or 1)
    panel_cols = int(min(N, panel_cols))
    cols = np.arange((N - panel_cols) // 2, (N - panel_cols) // 2 + panel_cols)
    all_j = list(range(0, N, jchunk))
    n_panels = -(-N // panel_cols)
    # probe one pair (also warms the BLAS threads up; not part of the sample)
    if probe_seconds is None:
        t0 = time.perf_counter()
        pt_panel(c, params, w, amp, A_list, didx, pts, cols, jchunk=jchunk, jstarts=all_j[:1])
        per = max(time.perf_counter() - t0, 1e-4)
    else:
        per = probe_seconds
    if per * len(all_j) * n_panels * 1.3 <= target_seconds:
        # ---- the whole inversion fits the budget: run it in full
        A_dense = [A[:, np.arange(N)] for A in A_list]
        gl = np.array(params, dtype=float)
        timers = {}
        t0 = time.perf_counter()
        mu, var, logl, _ = predict_lean(c, A_dense, didx, y, gl, sig, w, amp, panel=panel_cols, timers=timers)
        total = time.perf_counter() - t0
        return dict(seconds_estimated=total, seconds_measured=total,
                    stages=dict(kernel_eval=timers.get("kernel_eval", 0.0), dgemm_proj=timers.get("dgemm_proj", 0.0), aka=timers["aka"],
                                chol=timers["chol"], trsm=timers["trsm"], mean_var=timers["mean_var"]),
                    sample="full: the whole inversion in one process (%d panels of %d voxel columns x %d contraction chunks, AkA, Cholesky "
                           "M=%d, triangular solve, mean + variance)" % (n_panels, panel_cols, len(all_j), M),
                    full=True, workers=1, probe_seconds=per, pair_seconds=dict(median=per, min=per, max=per, n=0),
                    checksum=float(np.nansum(mu) + np.nansum(var)))
    # ---- bounded sample: `workers` processes x n pairs each (under contention a pair is slower than the probe: allow for 2x)
    n_each = int(max(1, min(len(all_j), 0.5 * target_seconds / (2.0 * per))))
    jobs = []
    for wk in range(workers):
        sel = [all_j[(wk * n_each + i) % len(all_j)] for i in range(n_each)]
        jobs.append((dict(cfg), None if gp_length is None else list(map(float, gp_length)), didx.tolist(), cols, jchunk, sel,
                     not state.get("workers_warm", False)))
    t0 = time.perf_counter()
    if pool is not None:
        res = pool.map(_pair_worker, jobs, chunksize=1)
    elif workers > 1:
        with mp.get_context("spawn").Pool(workers) as own:       # spawn: the caller may hold a CUDA context, which must not be forked
            res = own.map(_pair_worker, jobs, chunksize=1)
    else:
        res = [_pair_worker(jobs[0])]
    t_wall = time.perf_counter() - t0
    if pool is not None:
        state["workers_warm"] = True
    pair_t = np.concatenate([np.asarray(r[0]) for r in res])
    rate = sum(len(r[0]) / sum(r[0]) for r in res)               # pairs per second, all workers together
    k_ev, k_mm = sum(r[1] for r in res), sum(r[2] for r in res)
    share = k_ev / max(k_ev + k_mm, 1e-12)
    scale_cols = N / panel_cols
    n_pairs = scale_cols * len(all_j)
    t_proj = n_pairs / rate
    # ---- the other stages on one panel, BLAS on all threads
    if "Pt" not in state:
        state["Pt"] = pt_panel(c, params, w, amp, A_list, didx, pts, cols, jchunk=jchunk, jstarts=all_j[:1])      # operand for the timings below
    Pt = state["Pt"]
    t0 = time.perf_counter()
    AkA = np.empty((M, M))
    AkA[:Ns] = A_list[0][:, cols] @ Pt[:, 0, :].T
    AkA[Ns:2 * Ns] = A_list[1][:, cols] @ Pt[:, 1, :].T
    t_aka = time.perf_counter() - t0   # (the nd drill rows of AkA are gathers, not arithmetic)
    # SPD surrogate of the true size for the (size-dependent only) Cholesky timing
    full = np.eye(M)
    full[:2 * Ns, :2 * Ns] = 0.5 * (AkA[:2 * Ns, :2 * Ns] + AkA[:2 * Ns, :2 * Ns].T)
    full[np.diag_indices(M)] += np.abs(full).sum(axis=1) + sig[0] ** 2   # diagonally dominant => SPD (timing only)
    t0 = time.perf_counter()
    L = cholesky(full, lower=True)
    t_chol = time.perf_counter() - t0
    t0 = time.perf_counter()
    u = solve_triangular(L, y, lower=True)
    V = solve_triangular(L, Pt.reshape(M, -1), lower=True, check_finite=False)
    t_trsm = time.perf_counter() - t0
    t0 = time.perf_counter()
    mu = V.T @ u
    var = amp - np.einsum("ij,ij->j", V, V)
    t_mv = time.perf_counter() - t0
    est = t_proj + (t_aka + t_trsm + t_mv) * scale_cols + t_chol
    return dict(seconds_estimated=est, seconds_measured=t_wall + t_aka + t_chol + t_trsm + t_mv,
                stages=dict(kernel_eval=t_proj * share, dgemm_proj=t_proj * (1.0 - share),
                            aka=t_aka * scale_cols, chol=t_chol, trsm=t_trsm * scale_cols, mean_var=t_mv * scale_cols),
                sample="projection Pt=A.K: %d worker processes (1 BLAS thread each) x %d (panel of %d voxel columns, chunk of %d contraction voxels) "
                       "pairs with all %d sensor rows, each pair timed; aggregate %.2f pairs/s scaled to the cube's %.0f pairs; AkA / triangular "
                       "solve / mean+variance timed on one panel with BLAS on all threads, scaled x%.1f; Cholesky M=%d timed in full"
                       % (workers, n_each, panel_cols, jchunk, 2 * Ns, rate, n_pairs, scale_cols, M),
                full=False, workers=workers, probe_seconds=per,
                pair_seconds=dict(median=float(np.median(pair_t)), min=float(pair_t.min()), max=float(pair_t.max()), n=int(pair_t.size)),
                checksum=float(np.nansum(mu) + np.nansum(var)))


# --------------------------------------------------------------------------- synthetic truth (bench inputs)
def cylinders_truth(c, voxelpos):
    """simcube.py:83-92 ('cylinders'): density/magsus cubes as pure functions of voxel coordinates."""
    x3, y3, z3 = (v.reshape(c.yNcube, c.xNcube, c.zNcube) for v in voxelpos)
    rad = c.yLcube / 18.0
    rc1 = (y3 - c.yLcube / 1.3 - rad) ** 2 + (z3 + c.zLcube / 4 - rad) ** 2
    rc2 = (y3 - c.yLcube / 4.0 - rad) ** 2 + (z3 + c.zLcube / 4 - rad) ** 2
    density = x3 * 0.0 + 0.1
    density[rc2 <= rad ** 2] = 1.0
    density[rc1 <= rad ** 2] = 1.0
    density[(x3 < c.xLcube / 5.0) | (x3 > c.xLcube * 4.0 / 5.0)] = 0.1
    return density, c.gp_coeff[1] * density


# ------------------------------------------------------------------------------------------------ acquisition
def spherical2cartes(x0, y0, z0, phi, theta, r):
    """geobo/utils.py:21-37"""
    return x0 + r * np.sin(theta) * np.cos(phi), y0 + r * np.sin(theta) * np.sin(phi), z0 + r * np.cos(theta)


def futility_vertical(params, drill_rec, drill_var, kappa, beta, costs=None):
    """Restates geobo/run_geobo.py:175-200 with the module globals made explicit: minus the upper-confidence utility
    of a vertical hole through voxel column round(params), +inf for non-finite input or a border column."""
    p = np.asarray(params, dtype=float)
    if not np.all(np.isfinite(p)):
        return np.inf
    col = (int(np.round(p[0])), int(np.round(p[1])))
    n0, n1 = drill_rec.shape[:2]
    if not (0 < col[0] < n0 - 1 and 0 < col[1] < n1 - 1):
        return np.inf
    cost_term = 0.0 if costs is None else beta * np.sum(costs[col])
    utility = np.sum(drill_rec[col]) + kappa * np.sqrt(np.sum(drill_var[col])) - cost_term
    return -utility


def futility_drill(params, drill_rec, drill_var, kappa, beta, c, costs=None):
    """Restates geobo/run_geobo.py:203-235 (``c``: config with voxel sizes, zLcube, zmax): a core of length zLcube from
    (x0, y0, zmax) along (azimuth, dip) is sampled at int(2 L / min voxel size) points of np.linspace(0, L); the voxel
    of a sample is the truncated quotient of its coordinates by the voxel sizes on axes (0, 1, 2) (negative indices wrap
    as in NumPy indexing); any index outside the cube makes the reference's try/except return -0.0."""
    length = c.zLcube
    x0, y0, azimuth, dip = params
    nstep = int(2 * length / min(c.xvoxsize, c.yvoxsize, c.zvoxsize))
    r = np.linspace(0, length, nstep)
    phi = (r * 0 + azimuth) * np.pi / 180.
    theta = (180 - (r * 0 + dip)) * np.pi / 180.
    px, py, pz = spherical2cartes(r * 0 + x0, r * 0 + y0, r * 0 + c.zmax, phi, theta, r)
    with np.errstate(all="ignore"):
        idx = ((px / c.xvoxsize).astype(int), (py / c.yvoxsize).astype(int), (-pz / c.zvoxsize).astype(int))
    try:
        total = np.sum(drill_rec[idx]) + kappa * np.sqrt(np.sum(drill_var[idx]))
        total = total - beta * (np.sum(costs[idx]) if costs is not None else np.sum(drill_rec[idx] * 0.))
    except Exception:
        total = 0.
    return -total


# ------------------------------------------------------------------------------------------------ drill voxelisation
def _in_window(c, centre, half):
    return (centre - half <= c) & (c < centre + half)


def align_drill(coord, data, xxx, yyy, zzz, voxelsize):
    """run_geobo.py:132-159 (utils.align_drill2, utils.py:55-83): per voxel the nanmean of the samples inside the half-open
    window centre -/+ ONE VOXEL SIZE per axis; voxels without a finite mean keep 0.  Only voxels near some sample are
    visited (the others cannot select anything); selection and mean are NumPy's own, as in the reference loop."""
    import warnings
    coord = np.asarray(coord, dtype=float).reshape(-1, 3)
    data = np.asarray(data, dtype=float)
    centres = (np.asarray(xxx, dtype=float), np.asarray(yyy, dtype=float), np.asarray(zzz, dtype=float))
    res = np.zeros(centres[0].shape)
    if coord.shape[0] == 0:
        return res
    near = np.zeros(res.shape, dtype=bool)
    for lo in range(0, coord.shape[0], 256):
        close = np.ones(res.shape + (min(256, coord.shape[0] - lo),), dtype=bool)
        for ax in range(3):
            close &= np.abs(centres[ax][..., None] - coord[lo:lo + 256, ax]) <= voxelsize[ax]
        near |= close.any(axis=-1)
    for idx in zip(*np.nonzero(near)):
        sel = np.ones(coord.shape[0], dtype=bool)
        for ax in range(3):
            sel &= _in_window(coord[:, ax], centres[ax][idx], voxelsize[ax])
        picked = data[np.where(sel)]
        if picked.size:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")            # "Mean of empty slice" for all-NaN selections
                m = np.nanmean(picked)
            if np.isfinite(m):
                res[idx] = m
    return res


# ------------------------------------------------------------------------------------------------ simulator
def _binarise_top_decile(layer):
    """simcube.py:58-60: values below the 90th percentile -> 0, the rest -> the maximum."""
    cut = np.percentile(layer, 90)
    layer[layer < cut] = 0.0
    layer[layer >= cut] = layer.max()
    return layer


def syncube(c, modelname, voxelpos):
    """simcube.create_syncube (simcube.py:34-94) without the file output: density and susceptibility cubes (yN, xN, zN) of
    the models 'cylinders', 'layers_2', 'layers_3' as functions of the voxel-centre coordinates."""
    shp = (c.yNcube, c.xNcube, c.zNcube)
    x3, y3, z3 = (np.asarray(v).reshape(shp) for v in voxelpos)
    if modelname == "cylinders":
        density, _ = cylinders_truth(c, voxelpos)
    elif modelname in ("layers_2", "layers_3"):
        # the dip of the layers along y; 'layers_2' centres it on zLcube / 2 (simcube.py:56), 'layers_3' on yLcube / 2 (:68)
        mid = c.zLcube / 2 if modelname == "layers_2" else c.yLcube / 2.0
        with np.errstate(over="ignore"):
            zshift = c.zLcube / 8.0 * 1.0 / (1 + np.exp(2.0 * (-y3 + mid)))
        bands = {}
        for name, amp, top, bottom in (("l3", 6.0, 0.35, 0.375), ("l1", 4.0, 0.3, 0.325), ("l2", 8.0, 0.25, 0.275)):
            if name == "l3" and modelname == "layers_2":
                continue
            # one layer: amp * (S(u_top) - S(u_bottom)), S(u) = 1 / (1 + exp(-2u)), u = -z - zLcube * fraction + zshift
            with np.errstate(over="ignore"):
                bands[name] = _binarise_top_decile(amp * (1.0 / (1 + np.exp(-2 * (-z3 - c.zLcube * top + zshift)))
                                                         - 1.0 / (1 + np.exp(-2 * (-z3 - c.zLcube * bottom + zshift)))))
        density = 0.5 + bands["l1"] + bands["l2"]
        if modelname == "layers_3":
            density = density + bands["l3"]
    else:
        raise ValueError("unknown model %r" % modelname)
    return density, c.gp_coeff[1] * density


def synsurvey(c, density, magsus):
    """simcube.create_synsurvey (simcube.py:119-159) without the file output: sensors over the voxel-column centres at
    height zoff, gravity = A_grav . density, magnetic = A_magn . magsus; returns the two (yN, xN) maps and the locations."""
    Edges, _ = cube_geometry(c)
    xs = np.arange(c.xvoxsize / 2.0, c.xLcube + c.xvoxsize / 2.0, c.xvoxsize)
    ys = np.arange(c.yvoxsize / 2.0, c.yLcube + c.yvoxsize / 2.0, c.yvoxsize)
    xx, yy = np.meshgrid(xs, ys)
    loc = np.asarray([xx.flatten(), yy.flatten(), (xx * 0.0 + c.zoff).flatten()]).T
    grav = a_sens(c, c.magneticField * 0.0, loc, Edges, "grav") @ np.asarray(density).flatten()
    mag = a_sens(c, c.magneticField, loc, Edges, "magn") @ np.asarray(magsus).flatten()
    return grav.reshape(xx.shape), mag.reshape(xx.shape), loc
