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


class _OracleBackedContext:
    """Stand-in for the device context in the CPU test of the proposal drivers: answers from the oracle."""

    def __init__(self, c):
        self.c = c

    def acquisition_vertical(self, rec, var, kappa, beta, costs=None):
        out = np.full(rec.shape[:2], -np.inf)
        for a in range(rec.shape[0]):
            for b in range(rec.shape[1]):
                v = o.futility_vertical([a, b], rec, var, kappa, beta, costs)
                out[a, b] = -v if np.isfinite(v) else -np.inf
        return out

    def acquisition_drill(self, rec, var, voxsize, zmax, length, params, kappa, beta, costs=None):
        with np.errstate(all="ignore"):
            return np.array([-o.futility_drill(p, rec, var, kappa, beta, self.c, costs) for p in np.asarray(params).reshape(-1, 4)])


def _check_proposals(g, c, tmp_path, n_azimuth=8, n_dip=3):
    """bayesopt_vert / bayesopt_nonvert (run_geobo.py:246-362) as exhaustive sweeps: best proposal = the optimum over all
    candidates of the oracle's objective, reference units and rounding, csv files like the reference's."""
    from geobo_b200 import acquisition
    rec, var, costs, kappa, beta = g["rec"], g["var"], g["costs"], float(g["kappa"]), float(g["beta"])
    acquisition.set_cubes(rec, var)
    for cst in (None, costs):
        rows = acquisition.bayesopt_vert(top=6, costs=cst)
        every = np.array([[o.futility_vertical([a, b], rec, var, kappa, beta, cst) for b in range(rec.shape[1])] for a in range(rec.shape[0])])
        a, b = np.unravel_index(np.argmin(every), every.shape)
        assert rows.shape == (6, 3) and np.all(np.diff(rows[:, 2]) <= 0)                       # sorted by decreasing gain
        assert rows[0, 2] == np.round(-every[a, b], 4)
        assert rows[0, 0] == a * c.yvoxsize + c.ymin + 0.5 * c.yvoxsize and rows[0, 1] == b * c.xvoxsize + c.xmin + 0.5 * c.xvoxsize
        grid = dict(n_start=(5, 6), n_azimuth=n_azimuth, n_dip=n_dip)
        cand, util, props = acquisition.sweep_drill(costs=cst, top=4, **grid)
        assert cand.shape == (5 * 6 * n_azimuth * n_dip, 4) and cand[:, 2].max() < 360 and cand[:, 3].min() == 30 and cand[:, 3].max() == 90
        assert cand[:, 0].min() == c.yvoxsize and cand[:, 1].max() == c.xLcube - c.xvoxsize
        with np.errstate(all="ignore"):
            ref = np.array([-o.futility_drill(p, rec, var, kappa, beta, c, cst) for p in cand])
        assert np.array_equal(util == 0, ref == 0) and np.abs(util - ref).max() <= 1e-12 * np.abs(ref).max()
        assert abs(props[0, 4] - ref.max()) <= 1e-12 * abs(ref.max())
        rows = acquisition.bayesopt_nonvert(top=4, costs=cst, **grid)
        assert rows.shape == (4, 5) and rows[0, 4] == np.round(props[0, 4], 4)
        assert rows[0, 0] == np.round(np.round(props[0, 0], 2) + c.ymin, 1) and rows[0, 2] == props[0, 2]
    lines = open(tmp_path / "newdrill_proposals_vertical.csv").read().splitlines()
    assert lines[0] == "NORTHING,EASTING,BO_GAIN" and len(lines) == 7
    lines = open(tmp_path / "newdrill_proposals_non-vertical.csv").read().splitlines()
    assert lines[0] == "NORTHING,EASTING,AZIMUTH,DIP,BO_GAIN" and len(lines) == 5
    assert [float(v) for v in lines[1].split(",")] == list(rows[0])


def test_proposal_drivers_with_oracle_backed_context(monkeypatch, tmp_path):
    import os
    from geobo_b200 import _lib, config_loader
    g = load_golden("acquisition.npz")
    cfg = dict(json.loads(str(g["cfg"])), kappa=float(g["kappa"]), beta=float(g["beta"]), outpath=str(tmp_path) + os.sep)
    config_loader.load_settings(cfg, make_outpath=False)
    c = o.make_config(cfg)
    monkeypatch.setattr(_lib, "default_context", lambda: _OracleBackedContext(c))
    _check_proposals(g, c, tmp_path)


@pytest.mark.gpu
def test_gpu_proposal_drivers_vs_oracle(tmp_path):
    import os
    from geobo_b200 import config_loader
    g = load_golden("acquisition.npz")
    cfg = dict(json.loads(str(g["cfg"])), kappa=float(g["kappa"]), beta=float(g["beta"]), outpath=str(tmp_path) + os.sep)
    config_loader.load_settings(cfg, make_outpath=False)
    # 7 azimuths x 4 dips: no candidate direction sits on an exact multiple of 30 / 45 degrees except azimuth 0 and dip 90, so no
    # sample of a ray lands on a voxel face where one ulp of sin / cos would decide the voxel
    _check_proposals(g, o.make_config(cfg), tmp_path, n_azimuth=7, n_dip=4)
