"""The scalar formulas the CUDA kernels are built from (``csrc/formulas.cuh``), compiled for the host and checked on the
CPU against the reference fixtures: covariance functions with every branch, the 3 x 3 block selection of ``create_cov``,
the stationary-table + gather form the fused kernels use, the prism-corner potentials and the 8-corner assembly of
``A_sens``.  The device runs exactly these source lines (the rounding intrinsics map to plain IEEE operations here, the
harness is built with ``-ffp-contract=off``); what differs is libm's versus CUDA's exp / log / atan / sin / cos in the
last ulp -- the GPU parity tests (``test_gpu_parity.py``) close that gap.  Tolerances as there: 1e-13 on covariance
values (max|K| = 1), 2e-9 of max|A| on sensitivities.
"""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import HARNESS_CXX, ROOT, load_golden
from oracle import numpy_oracle as o

TOL_COV = 1e-13
TOL_SENS = 2e-9
KID = {"sparse": 0, "exp": 1, "matern32": 2}


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = tmp_path_factory.mktemp("formulas_host") / "formulas_host.so"
    src = os.path.join(ROOT, "tests", "host_harness", "formulas_host.cpp")
    subprocess.run(HARNESS_CXX + [src, "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    for name in ("host_cov_function", "host_create_cov", "host_create_cov_grid", "host_corner_func", "host_a_sens"):
        getattr(lib, name).restype = None
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def test_cov_functions_every_branch(host):
    rng = np.random.default_rng(1)
    d = np.concatenate([[0.0, 1.0, 2.9999, 3.0, 3.0001, 122.0, 247.0, 250.0, 253.0, 500.0], rng.uniform(0, 600, 500)])
    D2 = _f(d ** 2)
    out = np.empty_like(D2)
    cases = [("exp", 0, 250.0, 0.0, lambda: o.k_exp(D2, 250.0)), ("exp", 1, 244.0, 250.0, lambda: o.k_exp2(D2, 244.0, 250.0)),
             ("matern32", 0, 250.0, 0.0, lambda: o.k_matern32(D2, 250.0)), ("matern32", 1, 244.0, 250.0, lambda: o.k_matern32_2(D2, 244.0, 250.0)),
             ("sparse", 0, 250.0, 0.0, lambda: o.k_sparse(D2, 250.0)),
             ("sparse", 1, 244.0, 250.0, lambda: o.k_sparse2(D2, 244.0, 250.0)),      # |l2 - l1| / 2 = 3: branch A below, B up to 247
             ("sparse", 1, 250.0, 244.0, lambda: o.k_sparse2(D2, 250.0, 244.0)),
             ("sparse", 1, 250.0, 250.0, lambda: o.k_sparse2(D2, 250.0, 250.0))]       # equal scales: the 1e-3 nudge (kernels.py:125-126)
    for kf, cross, l1, l2, ref in cases:
        host.host_cov_function(KID[kf], cross, _p(D2), D2.size, ctypes.c_double(l1), ctypes.c_double(l2), _p(out))
        with np.errstate(all="ignore"):
            r = ref()
        assert np.abs(out - r).max() < TOL_COV, (kf, cross, l1, l2)
    # the compact kernel is exactly zero outside its support and never negative
    host.host_cov_function(KID["sparse"], 0, _p(D2), D2.size, ctypes.c_double(250.0), ctypes.c_double(0.0), _p(out))
    assert (out[d >= 250.0] == 0).all() and (out >= 0).all()


@pytest.mark.parametrize("fk", ["exp", "sparse", "matern32"])
def test_create_cov_block_selection_vs_reference_fixture(host, fk):
    from geobo_b200.kernels import dedup_lengthscales
    g = load_golden("kernels_small.npz")
    D2 = _f(g["D2"])
    n = D2.shape[0]
    out = np.empty((3 * n, 3 * n))
    variants = [("distinct", np.array([244.0, 250.0, 260.0]), [1.0, 0.2, 0.3])]
    if fk != "matern32":
        variants.append(("equal", np.array([244.0, 244.0, 244.0]), [1.0, 0.2, 0.2]))
    for tag, gl, w in variants:
        dedup_lengthscales(gl)                                               # host side of create_cov (kernels.py:174-180)
        host.host_create_cov(KID[fk], _p(_f(gl)), _p(_f(w)), ctypes.c_double(1.0), _p(D2), n, _p(out))
        assert np.abs(out - g["cov_%s_%s" % (fk, tag)]).max() < TOL_COV, tag
    host.host_create_cov(KID[fk], _p(_f(gl)), _p(_f(w)), ctypes.c_double(1.7), _p(D2), n, _p(out))
    assert np.abs(out - 1.7 * g["cov_%s_%s" % (fk, tag)]).max() < 2 * TOL_COV  # amplitude multiplies the finished block


@pytest.mark.parametrize("fk", ["exp", "sparse", "matern32"])
def test_stationary_tables_and_gather_equal_the_dense_matrix(host, fk):
    """K[(c, j), (r, i)] = table[c][r][L(i) - L(j) + C0] over the extended difference lattice (what the fused projection
    kernels evaluate) against the oracle's create_cov on the grid's distance matrix."""
    c = o.make_config(json.loads(str(load_golden("sens_8x6x5.npz")["cfg"])))
    ncube = np.array([5, 3, 4], dtype=np.int64)
    vox = _f([c.xvoxsize, c.yvoxsize, c.zvoxsize])
    gl, w = np.array([244.0, 250.0, 260.0]), [1.0, 0.2, 0.3]
    N = int(ncube.prod())
    out = np.empty((3 * N, 3 * N))
    host.host_create_cov_grid(KID[fk], _p(_f(gl)), _p(_f(w)), ctypes.c_double(1.0), _p(ncube), _p(vox), _p(out))
    pts = o.grid_points((5, 3, 4), tuple(vox))
    ref = o.create_cov(o.sqdist(pts), gl.copy(), w, fk)
    assert np.abs(out - ref).max() < TOL_COV
    assert np.abs(out - out.T).max() < 1e-15


def test_corner_potentials_and_a_sens_vs_reference_fixture(host):
    s = load_golden("sens_8x6x5.npz")
    c = o.make_config(json.loads(str(s["cfg"])))
    E, loc = _f(s["Edges"]), _f(s["locations"])
    ncube = np.array([c.xNcube, c.yNcube, c.zNcube], dtype=np.int64)
    out = np.empty((loc.shape[0], int(ncube.prod())))
    for kind, B, mul, div, key in [(0, c.magneticField * 0, c.c_MILLIGALS_UNITS, c.fcor_grav, "A_grav"), (1, c.magneticField, 1.0, c.fcor_mag, "A_magn"),
                                   (1, s["B_tilt"], 1.0, c.fcor_mag, "A_magn_tilt")]:
        host.host_a_sens(kind, _p(_f(B)), _p(loc), loc.shape[0], _p(E), _p(ncube), ctypes.c_double(mul), ctypes.c_double(div), _p(out))
        assert np.abs(out - s[key]).max() / np.abs(s[key]).max() < TOL_SENS, key
    x, y, z = _f(s["Edges"][0] - 100.0).ravel(), _f(s["Edges"][1] - 50.0).ravel(), _f(s["Edges"][2] + 1.0).ravel()
    got = np.empty_like(x)
    host.host_corner_func(0, _p(x), _p(y), _p(z), x.size, _p(_f([0, 0, 0])), _p(got))
    assert np.abs(got - o.grav_corner(x, y, z)).max() < 1e-9
    Bt = _f([3e-4, -2e-4, 9e-4])
    host.host_corner_func(1, _p(x), _p(y), _p(z), x.size, _p(Bt), _p(got))
    assert np.abs(got - o.magn_corner(x, y, z, *Bt)).max() < 1e-9
