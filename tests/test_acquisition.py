"""Acquisition sweep: the oracle's restatement of futility_vertical / futility_drill against
the fixture generated from the unmodified reference functions (CPU), and the CUDA sweep against both (GPU)."""
import json

import numpy as np
import pytest

from conftest import load_golden
from oracle import numpy_oracle as o


def _cfg(g):
    return o.make_config(json.loads(str(g["cfg"])))


def test_oracle_acquisition_matches_reference_fixture():
    g = load_golden("acquisition.npz")
    c = _cfg(g)
    rec, var, costs, kappa, beta = g["rec"], g["var"], g["costs"], float(g["kappa"]), float(g["beta"])
    fv = np.array([o.futility_vertical(p, rec, var, kappa, beta) for p in g["pv"]])
    fvc = np.array([o.futility_vertical(p, rec, var, kappa, beta, costs) for p in g["pv"]])
    assert np.array_equal(fv, g["fv"]) and np.array_equal(fvc, g["fvc"])
    with np.errstate(all="ignore"):
        fd = np.array([o.futility_drill(p, rec, var, kappa, beta, c) for p in g["pd"]])
        fdc = np.array([o.futility_drill(p, rec, var, kappa, beta, c, costs) for p in g["pd"]])
    assert np.array_equal(fd, g["fd"]) and np.array_equal(fdc, g["fdc"])
    assert (g["fd"] != 0).sum() > 50          # the fixture exercises rays that stay inside the cube


@pytest.mark.gpu
def test_gpu_acquisition_sweeps_match_reference_fixture():
    from geobo_b200 import acquisition, config_loader
    g = load_golden("acquisition.npz")
    cfg = json.loads(str(g["cfg"]))
    config_loader.load_settings(dict(cfg, kappa=float(g["kappa"]), beta=float(g["beta"])), make_outpath=False)
    rec, var, costs = g["rec"], g["var"], g["costs"]
    acquisition.set_cubes(rec, var)
    for cst, key in ((None, "fv"), (costs, "fvc")):
        util, props = acquisition.sweep_vertical(costs=cst, top=5)
        for p, ref in zip(g["pv"], g[key]):
            if not np.isfinite(p).all():
                assert acquisition.futility_vertical(p, cst) == np.inf
                continue
            a, b = int(np.round(p[0])), int(np.round(p[1]))
            inside = 0 <= a < rec.shape[0] and 0 <= b < rec.shape[1]
            got = -util[a, b] if inside else np.inf
            assert got == ref or abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (p, got, ref)
        best = np.nanmin(g[key][np.isfinite(g[key])])
        assert abs(-props[0, 2] - best) <= 1e-12 * abs(best)                      # the sweep finds the reference's optimum
        assert acquisition.futility_vertical([2.4, 3.6], cst) == pytest.approx(o.futility_vertical([2.4, 3.6], rec, var, float(g["kappa"]), float(g["beta"]), cst), rel=1e-12)
    for cst, key in ((None, "fd"), (costs, "fdc")):
        got = acquisition.futility_drill(g["pd"], cst)
        ref = g[key]
        assert np.array_equal(got == 0, ref == 0)                                  # same rays leave the cube
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    assert acquisition.futility_drill(g["pd"][3]) == pytest.approx(g["fd"][3], rel=1e-12)
