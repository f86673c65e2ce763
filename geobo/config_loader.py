"""Alias of ``geobo_b200.config_loader`` under the reference's module name (see ``geobo/__init__.py``)."""
import sys

from geobo_b200 import config_loader as _impl

sys.modules[__name__] = _impl
