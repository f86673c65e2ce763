"""Pin the CPU oracle (oracle/numpy_oracle.py) against the reference.

* the reference's own committed result cubes (examples/results/{cylinders,sample}/cube_*.vtk,
  stored in tests/golden/example{1,2}.npz together with the cubing inputs its driver produced), and
* outputs of the live, unmodified reference on tiny cubes (tests/golden/make_golden.py).
CPU only; no GPU, no /root/reference at run time.
"""
import numpy as np
import pytest

from conftest import CUBES, load_golden, normwise_err
from oracle import numpy_oracle as o


def test_grid_points_and_sqdist_bit_exact():
    g = load_golden("kernels_small.npz")
    pts = o.grid_points((4, 3, 2), (122.0, 61.0, 50.0))
    assert np.array_equal(pts, g["points"])
    assert np.array_equal(o.sqdist(pts), g["D2"])


@pytest.mark.parametrize("fk", ["exp", "sparse", "matern32"])
def test_create_cov_bit_exact(fk):
    g = load_golden("kernels_small.npz")
    gl = np.array([244.0, 250.0, 260.0])
    assert np.array_equal(o.create_cov(g["D2"], gl, [1.0, 0.2, 0.3], fk), g["cov_%s_distinct" % fk])
    assert np.array_equal(gl, g["gl_%s_distinct_after" % fk])
    if fk != "matern32":
        gl = np.array([244.0, 244.0, 244.0])
        assert np.array_equal(o.create_cov(g["D2"], gl, [1.0, 0.2, 0.2], fk), g["cov_%s_equal" % fk])
        # Q1: the caller's array is mutated in place to [L, 1.02 L, L]
        assert np.array_equal(gl, g["gl_%s_equal_after" % fk])
        assert np.allclose(gl, [244.0, 248.88, 244.0])


def test_create_cov_known_answers_survey_appendix_b():
    # 2-voxel D2, distinct scales
    D2 = np.array([[0.0, 300.0 ** 2], [300.0 ** 2, 0.0]])
    c = o.create_cov(D2, [244.0, 250.0, 260.0], [1.0, 0.2, 0.3], "exp")
    assert np.allclose(c[0], [1, 0.469613528227759, 0.299955747498264, 0.143473113690961, 0.998992696797993,
                              0.492179811628144], rtol=1e-14)
    c = o.create_cov(D2, [244.0, 250.0, 260.0], [1.0, 0.2, 0.3], "matern32")
    assert np.allclose(c[2], [0.299977871301096, 0.11358694216984, 1, 0.385185138004904, 0.199961549552658,
                              0.0791572901670361], rtol=1e-13)
    D2 = np.array([[0.0, 16.0], [16.0, 0.0]])
    c = o.create_cov(D2, [244.0, 250.0, 260.0], [1.0, 0.2, 0.3], "sparse")
    assert np.allclose(c[0], [1, 0.998233281624963, 0.299939840014131, 0.299423054305806, 0.99863303138058,
                              1.02340220923056], rtol=1e-13)   # >1: Q3, follow the code


def test_geometry_and_a_sens_bit_exact():
    s = load_golden("sens_8x6x5.npz")
    c = o.make_config(str(s["cfg"]))
    E, vp = o.cube_geometry(c)
    assert np.array_equal(E, s["Edges"]) and np.array_equal(vp, s["voxelpos"])
    assert np.array_equal(o.sensor_grid(c), s["locations"])
    assert np.array_equal(o.a_sens(c, c.magneticField * 0, s["locations"], E, "grav"), s["A_grav"])
    assert np.array_equal(o.a_sens(c, c.magneticField, s["locations"], E, "magn"), s["A_magn"])
    assert np.array_equal(o.a_sens(c, s["B_tilt"], s["locations"], E, "magn"), s["A_magn_tilt"])
    assert np.array_equal(o.a_drill(vp[:, s["drill_idx"]], vp), s["A_drill"])
    # known answers recorded in SURVEY.md Appendix B
    assert s["A_grav"][0, 0] == pytest.approx(0.145524793285828, rel=1e-13)
    assert s["A_magn"][20, 100] == pytest.approx(3.99005594192128, rel=1e-13)


@pytest.mark.parametrize("name", ["exp_nd7", "sparse_nd7", "matern32_nd7", "exp_nd0"])
def test_cubing_tiny_vs_live_reference(name):
    f = load_golden("cubing_%s.npz" % name)
    c = o.make_config(str(f["cfg"]))
    args = (c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    with np.errstate(all="ignore"):
        lit, ex = o.cubing_literal(*args, gp_length=f["gl_before"])
        lean, exl = o.cubing_lean(*args, gp_length=f["gl_before"])
    for n, a, b in zip(CUBES, lit, lean):
        assert np.array_equal(a, f[n], equal_nan=True), n        # literal restatement: bit-exact
        assert normwise_err(b, f[n]) < 1e-11, n                   # lean restatement: rounding only
    assert ex["logl"] == float(f["logl"])
    assert abs(exl["logl"] - float(f["logl"])) < 1e-9
    assert np.array_equal(ex["gl_after"], f["gl_after"])


@pytest.mark.parametrize("which", ["1", "2"])
def test_examples_vs_committed_vtk_goldens(which):
    """The reference's own golden vectors: the six committed VTK result cubes of each example."""
    f = load_golden("example%s.npz" % which)
    c = o.make_config(str(f["cfg"]))
    with np.errstate(all="ignore"):
        out, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    for n, a in zip(CUBES, out):
        # committed VTKs were produced by an older NumPy/SciPy/BLAS stack: agreement is <=4e-8
        assert normwise_err(a, f["gold_" + n]) < 2e-7, n
    assert normwise_err(out[0], f["live_density_rec"]) < 1e-10
    assert normwise_err(out[5], f["live_drill_var"]) < 1e-10
    assert abs(ex["logl"] - float(f["logl"])) < 1e-6
    assert np.allclose(ex["gl_after"], f["gl_after"])
    assert 2 * c.xNcube * c.yNcube + ex["didx"].size == int(f["M"])
