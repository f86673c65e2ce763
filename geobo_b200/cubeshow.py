"""``create_vtkcube`` of ``geobo/cubeshow.py:175-189`` without pyvista (the I/O step right after
the hot path, ``run_geobo.py:418-425``).

The reference builds a ``pyvista.UniformGrid`` and saves it; the file that comes out (and that the reference commits
under ``examples/results/``) is legacy VTK 4.2, ``BINARY``, ``STRUCTURED_POINTS`` with one big-endian float64 cell
array ``values`` holding the Fortran-order flattening of the cube.  This writer emits exactly those bytes, so the
cubes returned by ``Inversion.cubing`` can be written for ParaView on a box without pyvista / VTK.  The plotting
functions of the reference module (``skplot2``, ``skplot3``) are out of scope.
"""
import numpy as np


def _fmt(v):
    """VTK's ASCII number formatting of header values (shortest %g-style representation)."""
    return np.format_float_positional(float(v), trim="-") if float(v) == int(float(v)) and abs(float(v)) < 1e15 else repr(float(v))


def create_vtkcube(density, origin, voxelsize, fname):
    """Export a cube as a legacy VTK file (``cubeshow.py:175-189``).

    density: 3-D array (the cell values), origin: coordinates of the grid origin, voxelsize: spacing per axis,
    fname: output path.  Grid dimensions are ``density.shape + 1`` (points), the data are cell data named ``values``."""
    cube = np.asarray(density, dtype=np.float64)
    if cube.ndim != 3:
        raise ValueError("create_vtkcube needs a 3-D cube, got shape %r" % (cube.shape,))
    dims = [n + 1 for n in cube.shape]
    header = ("# vtk DataFile Version 4.2\nvtk output\nBINARY\nDATASET STRUCTURED_POINTS\n"
              "DIMENSIONS %d %d %d\nSPACING %s %s %s\nORIGIN %s %s %s\nCELL_DATA %d\nSCALARS values double\nLOOKUP_TABLE default\n"
              % (dims[0], dims[1], dims[2], _fmt(voxelsize[0]), _fmt(voxelsize[1]), _fmt(voxelsize[2]),
                 _fmt(origin[0]), _fmt(origin[1]), _fmt(origin[2]), cube.size))
    with open(fname, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(cube.flatten(order="F").astype(">f8").tobytes())
        f.write(b"\n")
