"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GeoBO joint-inversion hot path.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker (or as the timed
CPU baseline), never as the thing measured or shipped.  The product package
``geobo_b200`` never imports this package and has no CPU fallback.

Modules
-------
``numpy_oracle``  NumPy/SciPy restatement of the reference algorithm
                  (``geobo/kernels.py``, ``geobo/sensormodel.py``,
                  ``geobo/inversion.py``): a *literal* path (dense 3N x 3N
                  covariance, dense M x 3N forward matrix; small cubes only) and
                  a *lean* path (same arithmetic, block structure exploited,
                  variance diagonal only) that scales to the benchmark sizes.
``ref_loader``    imports the UNMODIFIED reference from ``/root/reference``
                  (build container only; absent on the GPU box).  Used by
                  ``tests/golden/make_golden.py`` to produce the committed
                  fixtures and by ``tests/test_oracle_vs_reference.py`` (skipped
                  when the reference tree is not there).
``vtkio``         reader for the legacy-VTK result cubes the reference commits
                  (``examples/results/*/cube_*.vtk``).

Parity status: PINNED.  The oracle reproduces the reference's twelve committed
result cubes (examples 1 and 2) and agrees with the live reference on every
fixture under ``tests/golden/`` (see ``tests/test_oracle_golden.py``).
"""
