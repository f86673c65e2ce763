"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and the reference-shaped
Python surface, against the CPU oracle and the committed golden fixtures.  Run with ``-m gpu``.

Tolerances (float64 path).  The arithmetic is the reference's, so differences come only from
(a) last-ulp differences between CUDA's and NumPy's exp/log/atan/sin/cos, (b) summation order in
the tensor-pipe GEMMs / Cholesky, amplified by cond(AkA) ~ 1e4..1e5:
  * covariance functions / create_cov: 1e-13 relative to max|K| = 1
  * A_sens: 2e-9 relative to max|A| (the 1e6 m edge padding turns 1 ulp of log() into ~1e-9 absolute)
  * posterior mean / variance cubes: 1e-7 norm-wise (max|delta| / max|ref| per cube); the target the
    north star states is 1e-5.
"""
import json

import numpy as np
import pytest

from conftest import CUBES, load_golden, normwise_err
from oracle import numpy_oracle as o

pytestmark = pytest.mark.gpu

TOL_CUBE = 1e-7
TOL_COV = 1e-13
TOL_SENS = 2e-9


def configure(cfg_json, **overrides):
    from geobo_b200 import config_loader
    cfg = json.loads(str(cfg_json)) if not isinstance(cfg_json, dict) else dict(cfg_json)
    cfg.update(overrides)
    config_loader.load_settings(cfg, make_outpath=False)
    return o.make_config(cfg)


@pytest.fixture(scope="module")
def ctx():
    from geobo_b200 import _lib
    return _lib.default_context()


def test_library_is_native_and_on_b200(ctx):
    info = ctx.device_info()
    assert info["cc"][0] == 10, info
    assert info["sm_count"] >= 100


def test_grid_points_and_sqdist(ctx):
    from geobo_b200 import kernels
    g = load_golden("kernels_small.npz")
    pts = kernels.calcGridPoints3D((4, 3, 2), (122.0, 61.0, 50.0))
    assert np.array_equal(pts, g["points"])
    assert np.array_equal(kernels.calcDistanceMatrix(pts), g["D2"])
    # ragged / odd shape against the oracle
    pts = kernels.calcGridPoints3D((7, 1, 3), (3.5, 2.25, 1.125))
    assert np.array_equal(pts, o.grid_points((7, 1, 3), (3.5, 2.25, 1.125)))
    assert np.array_equal(kernels.calcDistanceMatrix(pts), o.sqdist(pts))


@pytest.mark.parametrize("fk", ["exp", "sparse", "matern32"])
def test_create_cov_vs_reference_fixture(ctx, fk):
    from geobo_b200 import kernels
    g = load_golden("kernels_small.npz")
    gl = np.array([244.0, 250.0, 260.0])
    c = kernels.create_cov(g["D2"], gl, [1.0, 0.2, 0.3], fkernel=fk)
    assert c.shape == (72, 72) and c.dtype == np.float64
    assert np.abs(c - g["cov_%s_distinct" % fk]).max() < TOL_COV
    if fk != "matern32":
        gl = np.array([244.0, 244.0, 244.0])
        c = kernels.create_cov(g["D2"], gl, [1.0, 0.2, 0.2], fkernel=fk)
        assert np.abs(c - g["cov_%s_equal" % fk]).max() < TOL_COV
        assert np.array_equal(gl, g["gl_%s_equal_after" % fk])     # Q1: caller's ndarray mutated in place
        lst = [244.0, 244.0, 244.0]
        kernels.create_cov(g["D2"], lst, [1.0, 0.2, 0.2], fkernel=fk)
        assert lst == [244.0, 244.0, 244.0]                         # a list is not mutated


def test_cov_functions_elementwise_incl_branches(ctx):
    from geobo_b200 import kernels
    rng = np.random.default_rng(1)
    d = np.concatenate([[0.0, 1.0, 2.9999, 3.0, 3.0001, 122.0, 247.0, 250.0, 253.0, 500.0], rng.uniform(0, 600, 500)])
    D2 = (d ** 2).reshape(34, 15)
    for got, ref in [
        (kernels.gpkernel(D2, 244.0), o.k_exp(D2, 244.0)),
        (kernels.gpkernel2(D2, (244.0, 250.0)), o.k_exp2(D2, 244.0, 250.0)),
        (kernels.gpkernel_sparse(D2, 244.0), o.k_sparse(D2, 244.0)),
        (kernels.gpkernel_sparse2(D2, (244.0, 250.0)), o.k_sparse2(D2, 244.0, 250.0)),
        (kernels.gpkernel_sparse2(D2, (250.0, 244.0)), o.k_sparse2(D2, 250.0, 244.0)),
        (kernels.gpkernel_sparse2(D2, (244.0, 244.0)), o.k_sparse2(D2, 244.0, 244.0)),   # l1 == l2 offset branch
        (kernels.gpkernel_matern32(D2, 244.0), o.k_matern32(D2, 244.0)),
        (kernels.gpkernel_matern32_2(D2, (244.0, 250.0)), o.k_matern32_2(D2, 244.0, 250.0)),
    ]:
        assert got.shape == D2.shape
        assert np.abs(got - ref).max() < TOL_COV
    with np.errstate(all="ignore"):
        assert np.isnan(kernels.gpkernel_matern32_2(D2, (244.0, 244.0))).all()   # singular like the reference


def test_create_cov_grid_matches_dense(ctx):
    c = configure(load_golden("sens_8x6x5.npz")["cfg"], xNcube=5, yNcube=3, zNcube=4)
    for fk in ("exp", "sparse", "matern32"):
        gl = np.array([244.0, 250.0, 260.0])
        K, ms = ctx.create_cov_grid((5, 3, 4), (c.xvoxsize, c.yvoxsize, c.zvoxsize), gl, [1.0, 0.2, 0.3], 1.0, fk)
        pts = o.grid_points((5, 3, 4), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
        ref = o.create_cov(o.sqdist(pts), gl.copy(), [1.0, 0.2, 0.3], fk)
        assert np.abs(K - ref).max() < TOL_COV, fk
        assert np.array_equal(K, K.T) or np.abs(K - K.T).max() < 1e-15


def test_a_sens_vs_reference_fixture(ctx):
    from geobo_b200 import sensormodel
    s = load_golden("sens_8x6x5.npz")
    c = configure(s["cfg"])
    for B, func, key in [(c.magneticField * 0, "grav", "A_grav"), (c.magneticField, "magn", "A_magn"),
                         (s["B_tilt"], "magn", "A_magn_tilt")]:
        A, ez = sensormodel.A_sens(B, s["locations"], s["Edges"], func)
        assert A.shape == s[key].shape and ez is None
        assert np.abs(A - s[key]).max() / np.abs(s[key]).max() < TOL_SENS, key
    Ad = sensormodel.A_drill(s["voxelpos"][:, s["drill_idx"]], s["voxelpos"])
    assert np.array_equal(Ad, s["A_drill"])
    x, y, z = s["Edges"][0] - 100.0, s["Edges"][1] - 50.0, s["Edges"][2] + 1.0
    assert np.abs(sensormodel.grav_func(x, y, z) - o.grav_corner(x, y, z)).max() < 1e-9
    assert np.abs(sensormodel.magn_func(x, y, z, 3e-4, -2e-4, 9e-4) - o.magn_corner(x, y, z, 3e-4, -2e-4, 9e-4)).max() < 1e-9


def run_cubing(f, gl=None):
    from geobo_b200 import inversion
    inv = inversion.Inversion()
    inv.create_cubegeometry()
    if gl is not None:
        inv.gp_length = np.array(gl, dtype=float)
    out = inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    return inv, out


@pytest.mark.parametrize("name", ["exp_nd7", "sparse_nd7", "matern32_nd7", "exp_nd0"])
def test_cubing_tiny_vs_live_reference_fixture(ctx, name):
    f = load_golden("cubing_%s.npz" % name)
    configure(f["cfg"])
    inv, out = run_cubing(f, gl=f["gl_before"])
    for n, a in zip(CUBES, out):
        assert a.shape == f[n].shape
        assert normwise_err(a, f[n]) < TOL_CUBE, n
    assert abs(inv.logl - float(f["logl"])) < 1e-7 * abs(float(f["logl"]))
    assert np.array_equal(inv.gp_length, f["gl_after"])          # Q1 mutation visible on the instance
    nvox = inv.mu_rec.size // 3
    if f["Fs3"].size == 2 * 48:
        # no drill data: the drill property is NaN in the reference's cubes (std of an empty array, Q9); the library does not
        # compute that block at all, so mu_rec carries NaN there instead of the reference's (unused) finite numbers
        assert normwise_err(inv.mu_rec[:2 * nvox], f["mu_rec"][:2 * nvox]) < TOL_CUBE and np.isnan(inv.mu_rec[2 * nvox:]).all()
    else:
        assert normwise_err(inv.mu_rec, f["mu_rec"]) < TOL_CUBE
    # dense attributes the reference keeps on the instance, materialised lazily
    M = f["Fs3"].size
    assert inv.Asens3.shape == (M, 3 * 240)
    cov = np.asarray(inv.cov_rec)
    assert cov.shape == (720, 720)
    assert np.nanmax(np.abs(np.diag(cov) - inv.cov_rec.diagonal())) < 1e-9


@pytest.mark.parametrize("which", ["1", "2"])
def test_examples_vs_committed_vtk_goldens(ctx, which):
    """The reference's own golden vectors (examples/results/*/cube_*.vtk), N = 6400, M = 1056 / 1668."""
    f = load_golden("example%s.npz" % which)
    configure(f["cfg"])
    inv, out = run_cubing(f)
    for n, a in zip(CUBES, out):
        assert normwise_err(a, f["gold_" + n]) < 2e-7, n         # VTKs come from an older BLAS stack (<= 4e-8)
    assert normwise_err(out[0], f["live_density_rec"]) < TOL_CUBE
    assert normwise_err(out[5], f["live_drill_var"]) < TOL_CUBE
    assert abs(inv.logl - float(f["logl"])) < 1e-6
    assert np.allclose(inv.gp_length, f["gl_after"])


def synthetic_inputs(c, nd, seed=0):
    """SURVEY.md 8(d) recipe on the oracle side (CPU): cylinders truth, A.rho surveys rounded through float32."""
    E, vp = o.cube_geometry(c)
    loc = o.sensor_grid(c)
    dens, mags = o.cylinders_truth(c, vp)
    dens = dens + 0.05 * np.sin(vp[0].reshape(dens.shape) / 400.0)     # avoid degenerate constant data on tiny cubes
    mags = c.gp_coeff[1] * dens
    Ag = o.a_sens(c, c.magneticField * 0, loc, E, "grav")
    Am = o.a_sens(c, c.magneticField, loc, E, "magn")
    grav = (Ag @ dens.ravel()).astype(np.float32).astype(np.float64)
    mag = (Am @ mags.ravel()).astype(np.float32).astype(np.float64)
    N = c.xNcube * c.yNcube * c.zNcube
    d0 = np.zeros(N)
    if nd:
        idx = np.random.default_rng(seed).choice(N, nd, replace=False)
        d0[idx] = dens.ravel()[idx]
    d0 = d0.reshape(c.xNcube, c.yNcube, c.zNcube)
    return dict(grav=grav, mag=mag, drillfield=d0[d0 != 0], sensor_locations=loc, drilldata0=d0)


BASE = None


def base_cfg():
    global BASE
    if BASE is None:
        BASE = json.loads(str(load_golden("sens_8x6x5.npz")["cfg"]))
    return dict(BASE)


@pytest.mark.parametrize("shape,kf,nd", [((7, 5, 3), "exp", 3), ((9, 2, 1), "sparse", 2), ((3, 4, 19), "matern32", 5),
                                         ((16, 16, 16), "exp", 50), ((17, 13, 9), "sparse", 0)])
def test_cubing_odd_shapes_vs_oracle(ctx, shape, kf, nd):
    """Shapes that are not multiples of any tile size (N, Ns, M all ragged), all three kernels, nd = 0 and > 0."""
    c = configure(base_cfg(), xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc=kf)
    f = synthetic_inputs(c, nd)
    gl = c.gp_lengthscale * c.xvoxsize * (np.array([1.0, 1.01, 1.02]) if kf == "matern32" else np.ones(3))
    with np.errstate(all="ignore"):
        ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], gp_length=gl.copy())
    inv, out = run_cubing(f, gl=gl.copy())
    for n, a, r in zip(CUBES, out, ref):
        assert normwise_err(a, r) < TOL_CUBE, n
    assert abs(inv.logl - ex["logl"]) < 1e-7 * abs(ex["logl"])


@pytest.mark.parametrize("outer", ["1", "3", "4"])
def test_two_level_cholesky_flag_vs_oracle(ctx, monkeypatch, outer):
    """GEOBO_B200_CHOL_OUTER (default 4): the panels of a 512-wide block are factored left-looking and the trailing matrix is
    updated once per block (K = 512).  M = 1155 -> 10 panels = two full outer blocks and a ragged one of two panels;
    1 = the plain right-looking factorisation, 3 = three full blocks and a single trailing panel."""
    c = configure(base_cfg(), xNcube=24, yNcube=24, zNcube=4, kernelfunc="exp")
    f = synthetic_inputs(c, 3)
    ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    monkeypatch.setenv("GEOBO_B200_CHOL_OUTER", outer)
    inv, out = run_cubing(f)
    for n, a, r in zip(CUBES, out, ref):
        assert normwise_err(a, r) < TOL_CUBE, n
    assert abs(inv.logl - ex["logl"]) < 1e-7 * abs(ex["logl"])


@pytest.mark.parametrize("outer", ["1", "4"])
def test_cholesky_lookahead_flag_vs_oracle(ctx, monkeypatch, outer):
    """GEOBO_B200_CHOL_LOOKAHEAD=1: the trailing update is split into the next block's columns (main stream) and the rest (side
    stream, concurrent with the next panels).  Every element sees the same updates in the same order, so the cubes must equal
    those of the plain factorisation (to rounding at most) -- and the oracle's within the parity tolerance."""
    c = configure(base_cfg(), xNcube=24, yNcube=24, zNcube=4, kernelfunc="exp")
    f = synthetic_inputs(c, 3)
    ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    monkeypatch.setenv("GEOBO_B200_CHOL_OUTER", outer)
    _, plain = run_cubing(f)
    monkeypatch.setenv("GEOBO_B200_CHOL_LOOKAHEAD", "1")
    inv, out = run_cubing(f)
    for n, a, b, r in zip(CUBES, out, plain, ref):
        assert normwise_err(a, r) < TOL_CUBE, n
        assert normwise_err(a, b) < 1e-12, n
    assert abs(inv.logl - ex["logl"]) < 1e-7 * abs(ex["logl"])


def test_not_positive_definite_exits_like_reference(ctx, capsys):
    """matern32 with equal scales: 0/0 in the cross term -> Cholesky fails -> two prints + sys.exit(1) (inversion.py:99-104)."""
    c = configure(base_cfg(), xNcube=6, yNcube=5, zNcube=4, kernelfunc="matern32")
    f = synthetic_inputs(c, 2)
    with pytest.raises(SystemExit) as e:
        run_cubing(f)
    assert e.value.code == 1
    assert "Cholesky decompostion failed" in capsys.readouterr().out


@pytest.mark.parametrize("prec", ["fp64", "int8x5"])
def test_non_finite_covariance_without_drill_data_exits_like_reference(ctx, capsys, prec):
    """matern32 with equal scales and NO drill rows: the block-structured products never touch the NaN cross block (0, 2), but the
    reference's dense Asens3 . kcov . Asens3^T spreads 0 * NaN over all of AkA, so it always exits (inversion.py:98-104)."""
    c = configure(base_cfg(), xNcube=6, yNcube=5, zNcube=16, kernelfunc="matern32", precision=prec)
    f = synthetic_inputs(c, 0)
    with pytest.raises(SystemExit) as e:
        run_cubing(f)
    assert e.value.code == 1
    assert "Cholesky decompostion failed" in capsys.readouterr().out


def test_repeated_cubing_reuses_the_device_problem_of_the_same_geometry(ctx):
    """A second cubing() on the same Inversion with the same cube / sensors / drilled voxels keeps the device problem (sensitivities,
    workspaces) and only uploads the new data; other drilled voxels rebuild it.  Either way the cubes equal a fresh instance's."""
    c = configure(base_cfg(), xNcube=7, yNcube=5, zNcube=3, kernelfunc="exp")
    f = synthetic_inputs(c, 3)
    inv, out1 = run_cubing(f)
    assert inv._problem_builds == 1
    g = dict(f, grav=f["grav"] * 1.5 + 0.1 * np.arange(f["grav"].size), mag=f["mag"][::-1].copy())
    out2 = inv.cubing(g["grav"], g["mag"], g["drillfield"], g["sensor_locations"], g["drilldata0"])
    assert inv._problem_builds == 1                                     # same device problem
    _, fresh = run_cubing(g)
    for n, a, b in zip(CUBES, out2, fresh):
        assert np.array_equal(a, b), n
    ref, _ = o.cubing_lean(c, g["grav"], g["mag"], g["drillfield"], g["sensor_locations"], g["drilldata0"])
    for n, a, r in zip(CUBES, out2, ref):
        assert normwise_err(a, r) < TOL_CUBE, n
    d0 = f["drilldata0"].copy().ravel()
    idx = np.flatnonzero(d0)
    d0[(idx[0] + 1) % d0.size], d0[idx[0]] = d0[idx[0]], 0.0            # another voxel drilled: another A_drill
    d0 = d0.reshape(f["drilldata0"].shape)
    out3 = inv.cubing(f["grav"], f["mag"], d0[d0 != 0], f["sensor_locations"], d0)
    assert inv._problem_builds == 2
    ref3, _ = o.cubing_lean(c, f["grav"], f["mag"], d0[d0 != 0], f["sensor_locations"], d0)
    for n, a, r in zip(CUBES, out3, ref3):
        assert normwise_err(a, r) < TOL_CUBE, n


def test_calc_logl_vs_oracle(ctx):
    c = configure(base_cfg(), xNcube=6, yNcube=5, zNcube=4, kernelfunc="exp")
    f = synthetic_inputs(c, 4)
    inv, _ = run_cubing(f)
    E, vp = o.cube_geometry(c)
    A = [o.a_sens(c, c.magneticField * 0, f["sensor_locations"], E, "grav"), o.a_sens(c, c.magneticField, f["sensor_locations"], E, "magn")]
    didx = o.drill_indices(f["drilldata0"])
    for params in ([1.0, 2.0, 1.0, 0.2, 0.2], [1.7, 3.1, 0.6, 0.5, 0.9]):
        ref = o.calc_logl(c, A, didx, inv.Fs3, params)
        got = inv.calc_logl(params)
        assert abs(got - ref) < 1e-7 * abs(ref)
    assert inv.calc_logl([-1.0, 2.0, 1.0, 0.2, 0.2]) == np.inf      # not PD -> +inf (inversion.py:150-152)


def test_full_size_properties_32cube(ctx):
    """BASELINE config 2 (32x32x32, exp, fp64, N = 32768, M = 2048) is too large for the CPU oracle inside a
    test, so check size-independent properties: linearity of the mean in the data, data-independence and
    bounds of the variance."""
    from geobo_b200 import _lib
    c = configure(base_cfg(), xNcube=32, yNcube=32, zNcube=32, kernelfunc="exp")
    E, vp = o.cube_geometry(c)
    loc = o.sensor_grid(c)
    prob = _lib.Problem(ctx, (32, 32, 32), (c.xvoxsize, c.yvoxsize, c.zvoxsize), E, loc, c.magneticField,
                        c.c_MILLIGALS_UNITS, c.fcor_grav, 1.0, c.fcor_mag, np.zeros(0, dtype=np.int64))
    gl = c.gp_lengthscale * c.xvoxsize * np.array([1.0, 1.02, 1.0])
    h = prob.hyper(gl, c.gp_err, c.gp_coeff, 1.0, "exp")
    rng = np.random.default_rng(3)
    y1, y2 = rng.standard_normal(prob.M), rng.standard_normal(prob.M)
    res = []
    for y in (y1, y2, 2.0 * y1 - 0.5 * y2):
        prob.set_data(y)
        mu, var, logl, info = prob.predict(h)
        assert info == 0 and np.isfinite(mu[:2]).all() and np.isfinite(var[:2]).all()
        assert np.isnan(mu[2]).all() and np.isnan(var[2]).all()           # no drill rows: the drill property block is not computed (NaN, Q9)
        res.append((mu[:2], var[:2]))
    scale = max(np.abs(res[0][0]).max(), np.abs(res[1][0]).max())
    assert np.abs(res[2][0] - (2.0 * res[0][0] - 0.5 * res[1][0])).max() < 1e-9 * scale
    assert np.array_equal(res[0][1], res[1][1])                       # variance does not depend on the data
    assert res[0][1].max() <= 1.0 + 1e-12 and res[0][1].min() > -1e-9  # 0 <= var <= gp_amp
    prob.set_data(y1)
    prob.predict(h)
    t = prob.timings()
    assert t["project"] > 0 and t["total"] >= t["project"]
    prob.close()
