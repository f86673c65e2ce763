"""Compare the int8 digit-slice tcgen05 projection (slices = 4, 5, 6) with the fp64 DMMA path on the device."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from geobo_b200 import _lib, config_loader, inversion, synth  # noqa: E402


def main():
    cases = [((8, 6, 16), "exp", 5), ((12, 10, 16), "matern32", 7), ((16, 16, 16), "sparse", 50)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        cases = [((32, 32, 32), "exp", 0), ((32, 32, 32), "matern32", 50)]
    if len(sys.argv) > 1 and sys.argv[1] == "cfg3":
        cases = [((64, 64, 32), "matern32", 50)]
    if len(sys.argv) > 1 and sys.argv[1] == "cfg3e":
        cases = [((64, 64, 32), "exp", 0)]
    ctx = _lib.default_context()
    for shape, kf, nd in cases:
        cfg = synth.settings(*shape, kernelfunc=kf)
        config_loader.load_settings(cfg, make_outpath=False)
        f = synth.make_inputs(nd=nd, seed=0, ctx=ctx)
        inv = inversion.Inversion()
        inv.create_cubegeometry()
        if kf == "matern32":
            inv.gp_length = inv.gp_length * np.array([1.0, 1.01, 1.02])
        inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
        prob = inv._problem
        h0 = inv._hyper()
        mu0, var0, logl0, info0 = prob.predict(h0)
        t0 = prob.timings()
        for S in (4, 5, 6):
            for nref in (0, 1, 2):
                h = prob.hyper(inv.gp_length, inv.gp_sigma, inv.coeffm, inv.gp_amp, kf, slices=S, refine=nref)
                prob.predict(h)
                mu, var, logl, info = prob.predict(h)
                t = prob.timings()
                emu = np.abs(mu - mu0).max() / np.abs(mu0).max()
                evar = np.abs(var - var0).max() / np.abs(var0).max()
                print("%s %-8s nd=%-3d S=%d refine=%d  mean err %.2e  var err %.2e  dlogl %.2e  info=%d | project %.2f ms (fp64 %.2f)  aka %.2f chol %.2f trsm %.2f (fp64 %.2f)  total %.2f (fp64 %.2f)"
                      % (shape, kf, nd, S, nref, emu, evar, abs(logl - logl0), info, t["project"], t0["project"], t["aka"], t["chol"], t["trsm"], t0["trsm"],
                         t["total"], t0["total"]), flush=True)


if __name__ == "__main__":
    main()
