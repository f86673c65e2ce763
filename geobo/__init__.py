"""``geobo`` -- alias package: the reference's module names on the B200 implementation.

The reference's driver and helpers import the hot path as ``from .config_loader import *``, ``from . import inversion``,
``from .sensormodel import *``, ``from . import cubeshow as cs`` (``geobo/run_geobo.py:385-389``, ``geobo/simcube.py:28-31``).
Each module of this package *is* the corresponding ``geobo_b200`` module (the same module object, registered under both
names), so those imports resolve to the CUDA implementation without editing a line of the callers: put this directory in
place of the reference's ``geobo/{config_loader,kernels,sensormodel,inversion,utils,simcube,cubeshow}.py`` and keep the
reference's own ``main.py`` / ``run_geobo.py`` next to them.
"""
