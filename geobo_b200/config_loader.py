"""Settings loader with the same surface as ``geobo/config_loader.py``.

The reference parses ``sys.argv[1]`` at import time and injects every YAML key plus a
few derived constants as module globals (``geobo/config_loader.py:20-59``); the other
modules star-import them.  This module keeps that surface -- the same names end up as
module globals, ``from geobo_b200.config_loader import *`` works after a load, and the
import-time ``sys.argv[1]`` / ``settings.yaml`` behaviour is preserved when such a file
exists -- and adds ``load_settings(path_or_dict)`` so a process can (re)configure itself
explicitly.  The hot-path modules read the globals at call time, not at import time.

Extra optional keys (defaults keep old YAMLs unchanged): ``device`` (CUDA ordinal).
"""
import os
import sys

import numpy as np
import yaml

_DERIVED = ("xLcube", "yLcube", "zmin", "magneticField", "fname_drilldata", "fname_gravsurvey", "fname_magsurvey",
            "c_MILLIGALS_UNITS", "xvoxsize", "yvoxsize", "zvoxsize", "Nsensor")
_loaded_keys = []
fname_settings = None


def load_settings(source, make_outpath=True):
    """Load a settings YAML (path) or a dict of the same keys; sets the module globals."""
    global fname_settings, _loaded_keys
    if isinstance(source, dict):
        cfg = dict(source)
        fname_settings = None
    else:
        with open(source) as f:
            cfg = yaml.safe_load(f)
        fname_settings = str(source)
    g = globals()
    for k in _loaded_keys:
        g.pop(k, None)
    for key in cfg:                                   # config_loader.py:35-36
        g[str(key)] = cfg[key]
    _loaded_keys = [str(k) for k in cfg] + list(_DERIVED)
    if make_outpath and cfg.get("outpath"):
        os.makedirs(cfg["outpath"], exist_ok=True)   # :39
    g["xLcube"] = cfg["xmax"] - cfg["xmin"]           # :41-42
    g["yLcube"] = cfg["ymax"] - cfg["ymin"]
    g["zmin"] = cfg["zmax"] - cfg["zLcube"]           # :44
    g["magneticField"] = np.asarray([cfg["XMAG"], cfg["YMAG"], cfg["ZMAG"]]) * 1e-3   # :46
    if "outpath" in cfg and "FNAME_drilldata" in cfg:
        g["fname_drilldata"] = cfg["outpath"] + cfg["FNAME_drilldata"]     # :48-50
        g["fname_gravsurvey"] = cfg["inpath"] + cfg["FNAME_gravsurvey"]
        g["fname_magsurvey"] = cfg["inpath"] + cfg["FNAME_magsurvey"]
    g["c_MILLIGALS_UNITS"] = cfg["c_G"] * cfg["c_SI_TO_MILLIGALS"] * cfg["c_GCM3_TO_SI"]   # :53
    g["xvoxsize"] = g["xLcube"] / cfg["xNcube"] * 1.0  # :56-58
    g["yvoxsize"] = g["yLcube"] / cfg["yNcube"] * 1.0
    g["zvoxsize"] = cfg["zLcube"] / cfg["zNcube"] * 1.0
    g["Nsensor"] = cfg["xNcube"] * cfg["yNcube"]      # :59
    g["cfg"] = cfg
    g["__all__"] = [k for k in g if not k.startswith("_") and k not in ("os", "sys", "np", "yaml", "load_settings", "require")]
    return cfg


def require(*names):
    """Raise a clear error if the settings have not been loaded yet."""
    g = globals()
    missing = [n for n in names if n not in g]
    if missing:
        raise RuntimeError("geobo_b200 settings not loaded (missing %s): pass a settings YAML as sys.argv[1] "
                           "or call geobo_b200.config_loader.load_settings(path)" % ", ".join(missing))


def _autoload():
    # config_loader.py:20-31: argv[1] if it is a file, else ./settings.yaml
    cand = None
    if len(sys.argv) == 2 and os.path.isfile(sys.argv[1]) and sys.argv[1].lower().endswith((".yaml", ".yml")):
        cand = sys.argv[1]
    elif os.path.isfile("settings.yaml") and os.environ.get("GEOBO_B200_AUTOLOAD", "1") == "1":
        cand = "settings.yaml"
    if cand is not None:
        try:
            load_settings(cand)
        except Exception as exc:   # a non-GeoBO yaml on the command line must not break `import`
            print("geobo_b200: could not load settings from %s: %s" % (cand, exc))


_autoload()
