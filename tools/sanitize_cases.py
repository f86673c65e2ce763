#!/usr/bin/env python
"""Tiny inversions covering every kernel family of the library, for `compute-sanitizer` (tools/gpu_sanitize.sh):
both precisions (fp64 DMMA / int8 tcgen05 digit slices) x dense + the three structured projections, calc_logl, align_drill,
the acquisition sweeps and the dense create_cov assembly.  Each case is checked against the CPU oracle, so a sanitizer run that
perturbs scheduling still has to produce the right cubes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from geobo_b200 import _lib, config_loader, inversion, synth  # noqa: E402
from oracle import numpy_oracle as o  # noqa: E402

CASES = {
    "fp64_dense": ((6, 5, 7), "matern32", 3, "fp64", "dense"),
    "int8_dense": ((5, 4, 16), "exp", 3, "int8x5", "dense"),
    "int8x6_dense": ((5, 4, 16), "sparse", 0, "int8x6", "dense"),
    "fp64_kron": ((6, 5, 7), "exp", 3, "fp64", "kron"),
    "int8_kron": ((5, 4, 16), "exp", 2, "int8x5", "kron"),
    "fp64_compact": ((6, 5, 7), "sparse", 3, "fp64", "compact"),
    "int8_compact": ((5, 4, 16), "sparse", 2, "int8x5", "compact"),
    "fp64_fft": ((6, 5, 7), "matern32", 3, "fp64", "fft"),
    "int8_fft": ((5, 4, 16), "matern32", 2, "int8x5", "fft"),
}


def run(name):
    shape, kf, nd, prec, structure = CASES[name]
    cfg = synth.settings(*shape, kernelfunc=kf, precision=prec, structure=structure)
    config_loader.load_settings(cfg, make_outpath=False)
    f = synth.make_inputs(nd=nd, seed=3)
    inv = inversion.Inversion()
    inv.create_cubegeometry()
    if kf == "matern32":
        inv.gp_length = inv.gp_length * np.array([1.0, 1.01, 1.02])
    gl = inv.gp_length.copy()
    out = inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
    nll = inv.calc_logl([1.0, config_loader.gp_lengthscale, 1.0, 0.2, 0.2])
    c = o.make_config(cfg)
    with np.errstate(all="ignore"):
        ref, _ = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], gp_length=gl)
    worst = 0.0
    for a, r in zip(out, ref):
        if np.isnan(r).all():
            assert np.isnan(a).all()
            continue
        worst = max(worst, float(np.abs(a - r).max() / np.abs(r).max()))
    # (matern32 with one common length-scale factor is singular: +inf like the reference)
    assert worst < 1e-6 and (np.isfinite(nll) or kf == "matern32"), (name, worst, nll)
    print("SANITIZE_CASE_OK %s worst=%.2e nll=%.6f" % (name, worst, nll), flush=True)


def extras():
    from geobo_b200 import acquisition, kernels, utils
    cfg = synth.settings(6, 5, 4, kernelfunc="sparse")
    config_loader.load_settings(dict(cfg, kappa=1.0, beta=0.1), make_outpath=False)
    pts = kernels.calcGridPoints3D((3, 2, 2), (1.0, 2.0, 3.0))
    D2 = kernels.calcDistanceMatrix(pts)
    K = kernels.create_cov(D2, np.array([2.0, 2.5, 3.0]), [1.0, 0.2, 0.3], fkernel="exp")
    assert K.shape == (36, 36) and np.isfinite(K).all()
    ctx = _lib.default_context()
    Kg, _ = ctx.create_cov_grid((3, 2, 4), (1.0, 2.0, 3.0), [2.0, 2.5, 3.0], [1.0, 0.2, 0.3], 1.0, "matern32")
    assert np.isfinite(Kg).all()
    rng = np.random.default_rng(0)
    rec, var = rng.standard_normal((5, 6, 4)), rng.random((5, 6, 4))
    u = ctx.acquisition_vertical(rec, var, 1.0, 0.1)
    assert u.shape == (5, 6)
    prm = np.column_stack([rng.uniform(0, 3000, 7), rng.uniform(0, 1900, 7), rng.uniform(0, 360, 7), rng.uniform(30, 90, 7)])
    ctx.acquisition_drill(rec, var, (config_loader.xvoxsize, config_loader.yvoxsize, config_loader.zvoxsize), 0.0, 400.0, prm, 1.0, 0.1)
    coord = np.column_stack([rng.uniform(0, 3050, 50), rng.uniform(0, 1952, 50), -rng.uniform(0, 800, 50)])
    utils.align_drill(coord, rng.standard_normal(50))
    print("SANITIZE_CASE_OK extras", flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES) + ["extras"]
    for n in names:
        extras() if n == "extras" else run(n)
    # give everything back before exit so that memcheck's leak check only reports real leaks (the default context and its buffer
    # cache otherwise live until the process dies)
    import gc
    gc.collect()
    ctx = _lib.default_context()
    ctx.release_cache()
    ctx.close()
