"""Synthetic-cube generator and simulated surveys (geobo/simcube.py:34-159).

Fixture ``tests/golden/simdata.npz`` (``make_simdata_golden.py``): ``create_syncube`` of the unmodified reference for the
three models on the example-1 geometry, and the reference's own committed simulator outputs
(``examples/testdata/synthetic/simcube_*.csv``, ``simsurveydata_*.csv``).  CPU: oracle and host generator are bit-exact
against the live reference cubes and the committed DENSITY columns, the oracle's surveys reproduce the committed
GRAVITY / MAGNETIC columns.  GPU: ``create_synsurvey`` (sensitivities and products on the device) against the same
committed columns; tolerance 1e-6 norm-wise (the oracle agrees with them to 3e-15 / 4e-11).
"""
import json
import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import numpy_oracle as o

MODELS = ("cylinders", "layers_2", "layers_3")
COMMITTED = ("cylinders", "layers_3")


def _setup(**extra):
    from geobo_b200 import config_loader
    f = load_golden("simdata.npz")
    cfg = dict(json.loads(str(f["cfg"])), **extra)
    config_loader.load_settings(cfg, make_outpath=False)
    return f, o.make_config(cfg)


@pytest.mark.parametrize("model", MODELS)
def test_oracle_syncube_is_bit_exact_against_the_reference(model):
    f, c = _setup()
    _, vp = o.cube_geometry(c)
    dens, mags = o.syncube(c, model, vp)
    assert np.array_equal(dens, f["density_" + model]) and np.array_equal(mags, f["magsus_" + model])
    if model in COMMITTED:                                     # the reference's committed csv of the same cube
        assert np.array_equal(np.asarray(vp).T, f["csv_xyz_" + model])
        assert np.array_equal(dens.ravel(), f["csv_density_" + model])
        assert np.abs(mags.ravel() - f["csv_magsus_" + model]).max() <= 1e-16      # written by an older NumPy: 1 ulp


@pytest.mark.parametrize("model", COMMITTED)
def test_oracle_synsurvey_reproduces_the_committed_survey_csv(model):
    f, c = _setup()
    grav, mag, loc = o.synsurvey(c, f["csv_density_" + model].reshape(c.yNcube, c.xNcube, c.zNcube),
                                 f["csv_magsus_" + model].reshape(c.yNcube, c.xNcube, c.zNcube))
    assert np.array_equal(loc[:, :2], f["csv_sensor_xy_" + model])
    for got, key in ((grav, "csv_grav_"), (mag, "csv_mag_")):
        ref = f[key + model]
        assert np.abs(got.ravel() - ref).max() / np.abs(ref).max() < 1e-10


@pytest.mark.parametrize("model", MODELS)
def test_create_syncube_matches_reference_and_writes_its_files(model, tmp_path):
    from geobo_b200 import inversion, simcube
    from oracle import vtkio
    f, c = _setup(inpath=str(tmp_path) + os.sep)
    vp = inversion.Inversion().create_cubegeometry()
    dens, mags = simcube.create_syncube(model, vp)
    assert dens.shape == (c.yNcube, c.xNcube, c.zNcube)
    assert np.array_equal(dens, f["density_" + model]) and np.array_equal(mags, f["magsus_" + model])
    rows = open(tmp_path / ("simcube_%s.csv" % model)).read().splitlines()
    assert rows[0] == "x,y,z,DENSITY,MAGSUS" and len(rows) == 1 + dens.size
    if model == "cylinders":                                   # first data row of the reference's committed file
        assert rows[1] == "61.0,61.0,-25.0,0.1," + repr(0.2 * 0.1)
    back = np.loadtxt(tmp_path / ("simcube_%s.csv" % model), delimiter=",", skiprows=1)
    assert np.array_equal(back[:, 3], dens.ravel()) and np.array_equal(back[:, :3], np.asarray(vp).T)
    assert np.array_equal(vtkio.read_cube(str(tmp_path / ("simcube_%s.vtk" % model))), dens)
    drill = open(tmp_path / ("simdrill_%s.csv" % model)).read().splitlines()
    assert drill[0] == "x,y,z,DENSITY,MAGSUS,SiteID" and (len(drill) - 1) % c.zNcube == 0 and 1 <= (len(drill) - 1) // c.zNcube <= 4
    with pytest.raises(ValueError):
        simcube.create_syncube("spheres", vp)


@pytest.mark.gpu
@pytest.mark.parametrize("model", COMMITTED)
def test_gpu_create_synsurvey_vs_committed_survey_csv(model, tmp_path):
    from geobo_b200 import simcube
    f, c = _setup(inpath=str(tmp_path) + os.sep)
    shape = (c.yNcube, c.xNcube, c.zNcube)
    grav2D, magn2D = simcube.create_synsurvey(model, f["csv_density_" + model].reshape(shape), f["csv_magsus_" + model].reshape(shape))
    assert grav2D.shape == magn2D.shape == (c.yNcube, c.xNcube)
    for got, key in ((grav2D, "csv_grav_"), (magn2D, "csv_mag_")):
        ref = f[key + model]
        assert np.abs(got.ravel() - ref).max() / np.abs(ref).max() < 1e-6
    back = np.loadtxt(tmp_path / ("simsurveydata_%s.csv" % model), delimiter=",", skiprows=1)
    assert np.array_equal(back[:, :2], f["csv_sensor_xy_" + model]) and np.array_equal(back[:, 2], grav2D.ravel())
