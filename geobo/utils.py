"""Alias of ``geobo_b200.utils`` under the reference's module name (see ``geobo/__init__.py``)."""
import sys

from geobo_b200 import utils as _impl

sys.modules[__name__] = _impl
