"""GPU parity tests of the tcgen05 int8 digit-slice projection (settings key ``precision: int8x5 | int8x6``).

The slice path computes Pt = A.K as an exact integer product of fixed-point operands (39 / 47 bits, balanced
digits, one exponent per sensor row and per covariance table); everything after it is the fp64 path.  Stated
tolerance: 1e-5 norm-wise (max|delta| / max|ref| per cube) on posterior mean and variance -- the north star's
figure; measured errors are 1e-8 .. 1e-11 and the assertions below are tighter than the stated tolerance.
"""
import numpy as np
import pytest

from conftest import CUBES, load_golden, normwise_err
from oracle import numpy_oracle as o
from test_gpu_parity import base_cfg, configure, ctx, run_cubing, synthetic_inputs  # noqa: F401

pytestmark = pytest.mark.gpu

TOL_STATED = 1e-5


@pytest.mark.parametrize("which,prec", [("1", "int8x5"), ("1", "int8x6"), ("2", "int8x6")])
def test_int8_examples_vs_committed_vtk_goldens(ctx, which, prec):
    """The reference's own golden cubes (25x16x16, sparse kernel) through the tensor-core slice path."""
    f = load_golden("example%s.npz" % which)
    configure(f["cfg"], precision=prec)
    inv, out = run_cubing(f)
    for n, a in zip(CUBES, out):
        assert normwise_err(a, f["gold_" + n]) < 1e-6 < TOL_STATED, n
    assert abs(inv.logl - float(f["logl"])) < 1e-3


@pytest.mark.parametrize("shape,kf,nd", [((5, 4, 16), "exp", 3), ((9, 8, 16), "sparse", 4), ((6, 5, 32), "matern32", 5),
                                         ((16, 16, 16), "exp", 50), ((11, 4, 48), "exp", 0)])
@pytest.mark.parametrize("prec", ["int8x5", "int8x6"])
def test_int8_cubing_vs_oracle(ctx, shape, kf, nd, prec):
    """Ragged sensor-row / voxel-column tiles (Ns and N not multiples of 128 / 80 / 96), all kernels, nd = 0 and > 0."""
    c = configure(base_cfg(), xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc=kf, precision=prec)
    f = synthetic_inputs(c, nd)
    gl = c.gp_lengthscale * c.xvoxsize * (np.array([1.0, 1.01, 1.02]) if kf == "matern32" else np.ones(3))
    with np.errstate(all="ignore"):
        ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], gp_length=gl.copy())
    inv, out = run_cubing(f, gl=gl.copy())
    tol = 1e-6 if prec == "int8x5" else 1e-7
    for n, a, r in zip(CUBES, out, ref):
        assert normwise_err(a, r) < tol < TOL_STATED, n
    assert abs(inv.logl - ex["logl"]) < 1e-5 * abs(ex["logl"])


def test_int8_needs_z_multiple_of_16(ctx):
    """The Toeplitz digit generation works on 16-voxel z segments: other cubes are refused loudly (no silent fallback)."""
    c = configure(base_cfg(), xNcube=6, yNcube=5, zNcube=12, kernelfunc="exp", precision="int8x6")
    f = synthetic_inputs(c, 2)
    with pytest.raises(Exception) as e:
        run_cubing(f)
    assert "zNcube" in str(e.value)


def test_precision_auto_default_uses_the_tensor_core_path_where_admissible(ctx, monkeypatch):
    """Settings without a ``precision`` key (every reference YAML): 'auto' = int8x5 when zNcube % 16 == 0 -- the reference's
    committed example 1 (25 x 16 x 16) then runs on tcgen05 and still reproduces its VTK goldens --, fp64 otherwise."""
    from geobo_b200 import inversion
    monkeypatch.delenv("GEOBO_B200_DEFAULT_PRECISION", raising=False)
    f = load_golden("example1.npz")
    cfg = configure(f["cfg"])
    from geobo_b200 import config_loader
    assert not hasattr(config_loader, "precision")
    inv, out = run_cubing(f)
    assert inv.precision_used == "int8x5"
    launches_int8 = inv.timings["launches"]
    for n, a in zip(CUBES, out):
        assert normwise_err(a, f["gold_" + n]) < 1e-6 < TOL_STATED, n
    configure(f["cfg"], precision="fp64")
    inv64, out64 = run_cubing(f)
    assert inv64.precision_used == "fp64" and inv64.timings["launches"] != launches_int8     # a different kernel sequence ran
    for n, a, b in zip(CUBES, out, out64):
        assert normwise_err(a, b) < 1e-6, n
    c = configure(base_cfg(), xNcube=7, yNcube=5, zNcube=12, kernelfunc="exp")
    assert inversion.Inversion().precision_used == "fp64"                                    # not admissible -> fp64, no refusal
    f2 = synthetic_inputs(c, 3)
    ref, _ = o.cubing_lean(c, f2["grav"], f2["mag"], f2["drillfield"], f2["sensor_locations"], f2["drilldata0"])
    _, out2 = run_cubing(f2)
    for n, a, r in zip(CUBES, out2, ref):
        assert normwise_err(a, r) < 1e-7, n


@pytest.mark.parametrize("shape,kf,nd", [((5, 6, 16), "exp", 3), ((9, 5, 16), "sparse", 0), ((6, 5, 32), "matern32", 5)])
def test_lean_mode_regenerated_sensitivities_match_resident_ones(ctx, monkeypatch, shape, kf, nd):
    """GEOBO_B200_LEAN_A=1: no resident fp64 sensitivities -- digit extraction, A3^T alpha, A3 z and gb_forward regenerate them in
    column chunks (two voxel rows per chunk here: several chunks, a ragged last one for odd yNcube).  Same digits, same
    exponents; only the summation order of the two refinement GEMVs changes."""
    from geobo_b200 import _lib
    c = configure(base_cfg(), xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc=kf, precision="int8x5")
    f = synthetic_inputs(c, nd)
    gl = c.gp_lengthscale * c.xvoxsize * (np.array([1.0, 1.01, 1.02]) if kf == "matern32" else np.ones(3))
    inv0, out0 = run_cubing(f, gl=gl.copy())
    A0 = inv0._problem.sens("magn")
    x = np.random.default_rng(0).standard_normal(A0.shape[1])
    fw0 = inv0._problem.forward("grav", x)
    monkeypatch.setenv("GEOBO_B200_LEAN_A", "1")
    monkeypatch.setenv("GEOBO_B200_LEAN_CHUNK_ROWS", "2")
    inv1, out1 = run_cubing(f, gl=gl.copy())
    assert np.array_equal(inv1._problem.sens("magn"), A0)
    fw1 = inv1._problem.forward("grav", x)
    assert np.abs(fw1 - fw0).max() <= 1e-12 * np.abs(fw0).max()
    for n, a, b in zip(CUBES, out1, out0):
        assert normwise_err(a, b) < 1e-11, n
    assert abs(inv1.logl - inv0.logl) < 1e-10 * abs(inv0.logl)
    with np.errstate(all="ignore"):
        ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], gp_length=gl.copy())
    for n, a, r in zip(CUBES, out1, ref):
        assert normwise_err(a, r) < 1e-6, n
    # streamed contraction (GEOBO_B200_STREAM_A8=1: the full-width digit blocks are never resident either -- 172 GB at 128x128x64):
    # one projection launch per column chunk on digits built for that chunk, added into Pt; with and without culling
    monkeypatch.setenv("GEOBO_B200_STREAM_A8", "1")
    for cull in ("1", "0"):
        monkeypatch.setenv("GEOBO_B200_CULL", cull)
        inv_s, out_s = run_cubing(f, gl=gl.copy())
        for n, a, b in zip(CUBES, out_s, out1):
            assert normwise_err(a, b) < 1e-9, (n, cull)           # same digits; only the order of the fp64 flush sums changes
    monkeypatch.delenv("GEOBO_B200_CULL")
    monkeypatch.setenv("GEOBO_B200_STREAM_A8", "0")
    # structured projections on a lean problem: the rows are regenerated in sensor-row chunks (7 rows here: ragged), and only the
    # digits of this rank's voxel columns are kept (N side of the AkA products)
    structure = {"exp": "kron", "sparse": "compact", "matern32": "fft"}[kf]
    monkeypatch.setenv("GEOBO_B200_LEAN_ROW_CHUNK", "7")
    configure(base_cfg(), xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc=kf, precision="int8x5", structure=structure)
    inv2, out2 = run_cubing(f, gl=gl.copy())
    for n, a, r in zip(CUBES, out2, ref):
        assert normwise_err(a, r) < 1e-6, n
    monkeypatch.setenv("GEOBO_B200_LEAN_A", "0")
    inv3, out3 = run_cubing(f, gl=gl.copy())
    for n, a, b in zip(CUBES, out2, out3):
        assert normwise_err(a, b) < 1e-11, n
    monkeypatch.setenv("GEOBO_B200_LEAN_A", "1")
    # the fp64 path needs resident rows: refused loudly on a lean problem
    configure(base_cfg(), xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc=kf, precision="fp64")
    with pytest.raises(_lib.GeoboB200Error) as e:
        run_cubing(f, gl=gl.copy())
    assert "lean" in str(e.value)


@pytest.mark.parametrize("shape,kf,nd", [((16, 12, 16), "sparse", 5), ((40, 8, 16), "exp", 3), ((12, 10, 32), "exp", 0), ((10, 6, 32), "matern32", 4)])
def test_zero_digit_culling_and_tile_pacing_are_bitwise_neutral(ctx, monkeypatch, shape, kf, nd):
    """The projection kernel skips K steps whose covariance digits are all zero for the whole tile (compact support of the sparse
    kernel; exp(-d^2 / 2 gamma^2) below the last digit) and paces its tile rounds; both leave every integer accumulator unchanged:
    the cubes must be BITWISE equal to the run that visits every K step with free-running CTAs."""
    c = configure(base_cfg(), xNcube=shape[0], yNcube=shape[1], zNcube=shape[2], kernelfunc=kf, precision="int8x5")
    f = synthetic_inputs(c, nd)
    gl = c.gp_lengthscale * c.xvoxsize * (np.array([1.0, 1.01, 1.02]) if kf == "matern32" else np.ones(3))
    monkeypatch.setenv("GEOBO_B200_CULL", "0")
    monkeypatch.setenv("GEOBO_B200_TILE_SYNC", "0")
    inv0, out0 = run_cubing(f, gl=gl.copy())
    t0 = inv0.timings["project"]
    monkeypatch.setenv("GEOBO_B200_CULL", "1")
    inv1, out1 = run_cubing(f, gl=gl.copy())
    monkeypatch.setenv("GEOBO_B200_TILE_SYNC", "1")
    inv2, out2 = run_cubing(f, gl=gl.copy())
    monkeypatch.setenv("GEOBO_B200_TILE_SYNC", "2")          # one round of slack
    monkeypatch.setenv("GEOBO_B200_TILE_SORT", "0")          # natural tile order instead of the order sorted by K-step count
    inv3, out3 = run_cubing(f, gl=gl.copy())
    for n, a, b, d, e in zip(CUBES, out0, out1, out2, out3):
        assert np.array_equal(a, b, equal_nan=True), n
        assert np.array_equal(a, d, equal_nan=True), n
        assert np.array_equal(a, e, equal_nan=True), n
    assert inv0.logl == inv1.logl == inv2.logl == inv3.logl
    with np.errstate(all="ignore"):
        ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], gp_length=gl.copy())
    for n, a, r in zip(CUBES, out2, ref):
        assert normwise_err(a, r) < 1e-6, n
    print("projection ms: every K step %.3f, culled %.3f" % (t0, inv1.timings["project"]))


def test_variance_product_in_column_chunks_is_bitwise_neutral(ctx, monkeypatch):
    """GEOBO_B200_B8_MB=1: the transposed digit blocks of Pt (N side of colsumsq(Linv . Pt)) are built and multiplied in column
    chunks through a small scratch (at 128x128x64 all of them would take 65 GB per GPU); per column nothing changes."""
    c = configure(base_cfg(), xNcube=12, yNcube=11, zNcube=32, kernelfunc="matern32", precision="int8x5")
    f = synthetic_inputs(c, 6)
    gl = c.gp_lengthscale * c.xvoxsize * np.array([1.0, 1.01, 1.02])
    inv0, out0 = run_cubing(f, gl=gl.copy())
    launches0 = inv0.timings["launches"]
    monkeypatch.setenv("GEOBO_B200_B8_MB", "1")
    inv1, out1 = run_cubing(f, gl=gl.copy())
    assert inv1.timings["launches"] > launches0                       # several chunks ran
    for n, a, b in zip(CUBES, out0, out1):
        assert np.array_equal(a, b), n


def test_int8_full_size_32cube_vs_fp64_path(ctx):
    """BASELINE config 2 size (N = 32768, M = 2048, exp kernel, cond ~ 1e6): slice path against the fp64 DMMA path
    on the same device problem, plus linearity of the mean in the data."""
    from geobo_b200 import _lib
    c = configure(base_cfg(), xNcube=32, yNcube=32, zNcube=32, kernelfunc="exp")
    E, vp = o.cube_geometry(c)
    loc = o.sensor_grid(c)
    prob = _lib.Problem(ctx, (32, 32, 32), (c.xvoxsize, c.yvoxsize, c.zvoxsize), E, loc, c.magneticField,
                        c.c_MILLIGALS_UNITS, c.fcor_grav, 1.0, c.fcor_mag, np.zeros(0, dtype=np.int64))
    gl = c.gp_lengthscale * c.xvoxsize * np.array([1.0, 1.02, 1.0])
    rng = np.random.default_rng(5)
    y1, y2 = rng.standard_normal(prob.M), rng.standard_normal(prob.M)
    prob.set_data(y1)
    mu0, var0, logl0, info0 = prob.predict(prob.hyper(gl, c.gp_err, c.gp_coeff, 1.0, "exp"))
    h6 = prob.hyper(gl, c.gp_err, c.gp_coeff, 1.0, "exp", slices=6)
    mu1, var1, logl1, info1 = prob.predict(h6)
    assert info0 == 0 and info1 == 0
    for a in (mu0, var0, mu1, var1):
        assert np.isnan(a[2]).all()               # no drill rows: the drill property block is not computed (NaN like the reference's cubes)
    mu0, var0, mu1, var1 = mu0[:2], var0[:2], mu1[:2], var1[:2]
    assert np.abs(mu1 - mu0).max() < 1e-6 * np.abs(mu0).max()
    assert np.abs(var1 - var0).max() < 1e-6 * np.abs(var0).max()
    assert abs(logl1 - logl0) < 1e-6 * abs(logl0)
    prob.set_data(y2)
    mu2, var2, _, _ = prob.predict(h6)
    prob.set_data(2.0 * y1 - 0.5 * y2)
    mu3, var3, _, _ = prob.predict(h6)
    mu2, var2, mu3 = mu2[:2], var2[:2], mu3[:2]
    scale = max(np.abs(mu1).max(), np.abs(mu2).max())
    assert np.abs(mu3 - (2.0 * mu1 - 0.5 * mu2)).max() < 1e-8 * scale
    assert np.array_equal(var1, var2)
    prob.close()


def test_int8_multi_flush_paths_match_single_flush(ctx):
    """GEOBO_B200_CHUNK=256 forces several accumulator flushes per tile (projection read-modify-write, AkA store,
    running sums of V before the squares) on a small cube; the result must equal the single-flush run up to fp64
    summation order.  Runs in a subprocess because the interval is read once per process."""
    import os
    import subprocess
    import sys
    from conftest import ROOT, check_subprocess
    code = r'''
import sys, json, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from geobo_b200 import config_loader, inversion, synth
cfg = synth.settings(12, 11, 32, kernelfunc="matern32", precision="int8x5")
config_loader.load_settings(cfg, make_outpath=False)
f = synth.make_inputs(nd=6, seed=2)
inv = inversion.Inversion(); inv.create_cubegeometry()
inv.gp_length = inv.gp_length * np.array([1.0, 1.01, 1.02])
out = inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
np.save(sys.argv[1], np.stack(out))
''' % (ROOT, ROOT)
    import tempfile
    res = []
    for chunk in ("16384", "256"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as tf:
            env = dict(os.environ, GEOBO_B200_CHUNK=chunk)
            r = subprocess.run([sys.executable, "-c", code, tf.name], env=env, capture_output=True, text=True, timeout=600)
            check_subprocess(r)
            res.append(np.load(tf.name))
    for a, b in zip(res[0], res[1]):
        assert np.abs(a - b).max() <= 1e-11 * np.abs(a).max()


def test_int8_calc_logl_vs_oracle(ctx):
    """Inversion.calc_logl (inversion.py:125-152) through the slice path: the marginal likelihood only needs the
    projection, AkA and the Cholesky (no refinement applies to log det), so the stated tolerance is the slice error."""
    c = configure(base_cfg(), xNcube=6, yNcube=5, zNcube=16, kernelfunc="exp", precision="int8x6")
    f = synthetic_inputs(c, 4)
    inv, _ = run_cubing(f)
    E, vp = o.cube_geometry(c)
    A = [o.a_sens(c, c.magneticField * 0, f["sensor_locations"], E, "grav"), o.a_sens(c, c.magneticField, f["sensor_locations"], E, "magn")]
    didx = o.drill_indices(f["drilldata0"])
    for params in ([1.0, 2.0, 1.0, 0.2, 0.2], [1.7, 3.1, 0.6, 0.5, 0.9]):
        ref = o.calc_logl(c, A, didx, inv.Fs3, params)
        got = inv.calc_logl(params)
        assert abs(got - ref) < 1e-6 * abs(ref)
    assert inv.calc_logl([-1.0, 2.0, 1.0, 0.2, 0.2]) == np.inf
