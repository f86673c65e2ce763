"""Alias of ``geobo_b200.inversion`` under the reference's module name (see ``geobo/__init__.py``)."""
import sys

from geobo_b200 import inversion as _impl

sys.modules[__name__] = _impl
