"""Import the UNMODIFIED reference (``/root/reference/geobo``) for fixture generation.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works only where the
reference tree exists (the build container); the GPU box has no
``/root/reference`` so nothing run there may call this.

The reference reads its YAML at *import* time from ``sys.argv[1]``
(``geobo/config_loader.py:20-36``) and star-imports the resulting globals into
every module, so one configuration == one Python process.  ``load(yaml_path)``
therefore must be called at most once per process; use ``run_in_subprocess`` to
evaluate several configurations.

Shim: ``geobo/kernels.py:23`` imports ``reshape, sqrt, identity`` from the
top-level ``scipy`` namespace (removed in modern SciPy) without using them; we
alias them to NumPy's before the import.  Nothing under ``/root/reference`` is
modified or copied.
"""
import importlib
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("GEOBO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "geobo", "inversion.py"))


def write_yaml(overrides, base=None, outdir=None):
    """Write a settings YAML = reference example-1 settings + ``overrides``; returns its path."""
    import yaml
    base = base or os.path.join(REFERENCE_ROOT, "examples", "settings_example1.yaml")
    with open(base) as f:
        cfg = yaml.safe_load(f)
    outdir = outdir or tempfile.mkdtemp(prefix="geobo_ref_")
    cfg["outpath"] = os.path.join(outdir, "out") + os.sep
    cfg["gen_simulation"] = False
    cfg.update(overrides or {})
    path = os.path.join(outdir, "settings.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path


def load(yaml_path):
    """Import reference ``kernels``, ``sensormodel``, ``inversion``, ``config_loader`` for one YAML."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "geobo.config_loader" in sys.modules:
        raise RuntimeError("reference already imported in this process (one YAML per process)")
    import numpy as np
    import scipy
    for name in ("reshape", "sqrt", "identity"):
        if not hasattr(scipy, name):
            setattr(scipy, name, getattr(np, name))
    sys.argv = [sys.argv[0] if sys.argv else "main.py", yaml_path]
    sys.path.insert(0, REFERENCE_ROOT)
    mods = {}
    for m in ("config_loader", "kernels", "sensormodel", "inversion"):
        mods[m] = importlib.import_module("geobo." + m)
    return mods


def cubing_inputs_from_truth(mods, nd, seed=0, model="cylinders"):
    """Synthetic ``cubing`` inputs per SURVEY.md 8(d) using the *reference's* forward model."""
    import numpy as np
    cl = mods["config_loader"]
    inv = mods["inversion"].Inversion()
    voxelpos = inv.create_cubegeometry()
    xN, yN, zN = cl.xNcube, cl.yNcube, cl.zNcube
    inv.xxx = voxelpos[0].reshape(xN, yN, zN)
    inv.yyy = voxelpos[1].reshape(xN, yN, zN)
    inv.zzz = voxelpos[2].reshape(xN, yN, zN)
    return inv, voxelpos
