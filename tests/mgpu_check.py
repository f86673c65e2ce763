"""Multi-GPU parity check; run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py
Every rank runs Inversion.cubing on its voxel-column shard; rank 0 compares the gathered cubes with the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from geobo_b200 import _lib, config_loader, dist, inversion, synth  # noqa: E402


def main():
    ctx = _lib.default_context()
    rank, world = dist.init_from_env(ctx)
    worst = 0.0
    cases = [((16, 16, 8), "exp", 5, "fp64"), ((12, 10, 9), "sparse", 0, "fp64"), ((16, 12, 6), "matern32", 7, "fp64"),
             ((16, 16, 16), "exp", 9, "int8x5"), ((12, 11, 32), "matern32", 6, "int8x6"), ((20, 16, 16), "sparse", 0, "int8x5")]
    if os.environ.get("GEOBO_B200_LEAN_A") == "1":
        cases = [case for case in cases if case[3] != "fp64"]          # lean problems only run the int8 tensor-core path
    cases = [case + ("dense",) for case in cases]
    if os.environ.get("GEOBO_B200_MGPU_STRUCTURED") == "1":
        # the opt-in structured projections; 12 x 11 x 16: the two voxel-column shards meet in the middle of an x-z plane
        cases = [((12, 11, 16), "exp", 4, "fp64", "kron"), ((16, 16, 8), "exp", 0, "fp64", "kron"), ((12, 11, 16), "exp", 5, "int8x5", "kron"),
                 ((12, 11, 16), "sparse", 4, "fp64", "compact"), ((12, 11, 16), "sparse", 0, "int8x5", "compact"),
                 ((12, 11, 16), "matern32", 4, "fp64", "fft"), ((16, 16, 8), "sparse", 0, "fp64", "fft"), ((12, 11, 16), "exp", 5, "int8x5", "fft")]
    for shape, kf, nd, prec, structure in cases:
        cfg = synth.settings(*shape, kernelfunc=kf, precision=prec, structure=structure)
        config_loader.load_settings(cfg, make_outpath=False)
        f = synth.make_inputs(nd=nd, seed=1, ctx=ctx)
        inv = inversion.Inversion()
        inv.create_cubegeometry()
        if kf == "matern32":
            inv.gp_length = inv.gp_length * np.array([1.0, 1.01, 1.02])
        gl = inv.gp_length.copy()
        out = inv.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
        if rank == 0:
            from oracle import numpy_oracle as o
            c = o.make_config(cfg)
            with np.errstate(all="ignore"):
                ref, ex = o.cubing_lean(c, f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"], gp_length=gl)
            for a, r in zip(out, ref):
                if np.isnan(r).all():
                    assert np.isnan(a).all()
                    continue
                worst = max(worst, float(np.abs(a - r).max() / np.abs(r).max()))
            assert abs(inv.logl - ex["logl"]) < (1e-7 if prec == "fp64" else 1e-5) * abs(ex["logl"]), (inv.logl, ex["logl"])
        dist.barrier()
    if rank == 0:
        assert worst < 1e-7, worst
        print("MGPU_OK world=%d worst_normwise_err=%.3e" % (world, worst))


if __name__ == "__main__":
    main()
