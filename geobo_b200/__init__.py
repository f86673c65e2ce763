"""geobo_b200 -- B200-native implementation of GeoBO's GP joint-inversion hot path.

Drop-in for the reference's ``geobo.config_loader`` / ``geobo.kernels`` /
``geobo.sensormodel`` / ``geobo.inversion`` modules on that path (see DESIGN.md and
INTEGRATION.md).  All numerics run in hand-written CUDA (``libgeobo_b200.so``, sm_100a)
behind a ctypes C ABI (``include/geobo_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
