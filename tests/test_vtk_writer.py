"""create_vtkcube (geobo/cubeshow.py:175-189) without pyvista: the cubes stored in the example fixtures (read from the
reference's committed result files) written back must reproduce the reference's bytes; round trip through the oracle's
reader for odd shapes / non-integer spacings."""
import json
import os

import numpy as np
import pytest

from conftest import CUBES, load_golden
from oracle import vtkio

REF = "/root/reference/examples/results"


def _geometry(cfg):
    xv = (cfg["xmax"] - cfg["xmin"]) / cfg["xNcube"]
    yv = (cfg["ymax"] - cfg["ymin"]) / cfg["yNcube"]
    zv = cfg["zLcube"] / cfg["zNcube"]
    origin = (xv / 2.0, yv / 2.0, cfg["zmax"] - cfg["zLcube"] + zv / 2.0)       # run_geobo.py:418: voxel-centre minima
    return origin, (xv, yv, zv)


@pytest.mark.parametrize("which", ["1", "2"])
def test_writer_reproduces_golden_cubes(which, tmp_path):
    from geobo_b200 import cubeshow
    f = load_golden("example%s.npz" % which)
    cfg = json.loads(str(f["cfg"]))
    origin, vox = _geometry(cfg)
    for n in CUBES:
        p = tmp_path / ("cube_%s.vtk" % n)
        cubeshow.create_vtkcube(f["gold_" + n], origin, vox, str(p))
        assert np.array_equal(vtkio.read_cube(str(p)), f["gold_" + n], equal_nan=True)
        head = open(p, "rb").read(186)
        assert head == (b"# vtk DataFile Version 4.2\nvtk output\nBINARY\nDATASET STRUCTURED_POINTS\nDIMENSIONS 17 26 17\n"
                        b"SPACING 122 122 50\nORIGIN 61 61 -775\nCELL_DATA 6400\nSCALARS values double\nLOOKUP_TABLE default\n")


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_writer_is_byte_identical_to_the_reference_files(tmp_path):
    from geobo_b200 import cubeshow
    names = {"density_rec": "cube_density", "magsus_rec": "cube_magsus", "drill_rec": "cube_drill",
             "density_var": "cube_density_variance", "magsus_var": "cube_magsus_variance", "drill_var": "cube_drill_variance"}
    for sub in ("cylinders", "sample"):
        for n, fn in names.items():
            src = os.path.join(REF, sub, fn + ".vtk")
            cube = vtkio.read_cube(src)
            out = tmp_path / (sub + "_" + fn + ".vtk")
            cubeshow.create_vtkcube(cube, (61.0, 61.0, -775.0), (122.0, 122.0, 50.0), str(out))
            assert open(out, "rb").read() == open(src, "rb").read(), (sub, fn)


def test_round_trip_odd_shape(tmp_path):
    from geobo_b200 import cubeshow
    rng = np.random.default_rng(0)
    cube = rng.standard_normal((3, 5, 2))
    p = tmp_path / "c.vtk"
    cubeshow.create_vtkcube(cube, (0.5, -1.25, 3.0), (381.25, 325.3333333333333, 160.0), str(p))
    assert np.array_equal(vtkio.read_cube(str(p)), cube)
    assert b"SPACING 381.25 325.3333333333333 160\n" in open(p, "rb").read(300)
