"""Reader for the legacy-VTK cubes the reference commits as results.  TEST INFRASTRUCTURE ONLY.

Format (written by the reference through pyvista, ``geobo/cubeshow.py:175-189``):
legacy VTK 4.2, ``BINARY``, ``STRUCTURED_POINTS``, ``DIMENSIONS nx+1 ny+1 nz+1``,
``CELL_DATA n``, one big-endian float64 scalar array after ``LOOKUP_TABLE default``.
The cell values are the Fortran-order flattening of the ``(yN, xN, zN)`` cube the
reference passed in, so ``read_cube`` returns that ``(yN, xN, zN)`` array.
"""
import numpy as np


def read_cube(path):
    with open(path, "rb") as f:
        raw = f.read()
    head_end = raw.index(b"LOOKUP_TABLE")
    head_end = raw.index(b"\n", head_end) + 1
    header = raw[:head_end].decode("ascii", "replace")
    dims = None
    ncell = None
    dtype = ">f8"
    for line in header.splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "DIMENSIONS":
            dims = tuple(int(t) - 1 for t in tok[1:4])
        elif tok[0] == "CELL_DATA":
            ncell = int(tok[1])
        elif tok[0] == "SCALARS":
            dtype = {"double": ">f8", "float": ">f4"}[tok[2]]
    if dims is None or ncell is None:
        raise ValueError("not a STRUCTURED_POINTS cell-data VTK file: %s" % path)
    vals = np.frombuffer(raw, dtype=dtype, count=ncell, offset=head_end).astype(np.float64)
    return vals.reshape(dims, order="F")
