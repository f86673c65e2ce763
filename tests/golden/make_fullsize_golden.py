#!/usr/bin/env python
"""Full-size CPU results of the oracle for the BASELINE configurations the bench is quoted on.

    python tests/golden/make_fullsize_golden.py cfg2 [--workers 6]     # 32x32x32, exp, nd = 0        (minutes)
    python tests/golden/make_fullsize_golden.py cfg3 [--workers 6]     # 64x64x32, matern32, nd = 50  (about 1.5 h on 6 cores)
    python tests/golden/make_fullsize_golden.py cfg3e [--workers 7]    # 64x64x32, exp, nd = 0        (about 1 h on 7 cores)

Build-container job (CPU only, nothing here needs the reference tree or a GPU).  It runs the oracle's lean
restatement of ``Inversion.cubing`` (``oracle/numpy_oracle.py``: ``a_sens``, ``pt_panel``, the arithmetic of
``predict_lean``) on the same synthetic inputs ``bench.py`` uses for that workload, except that the surveys are
forward-simulated with the oracle's own ``a_sens`` -- and are stored in the fixture, so the GPU test feeds the device
path exactly these numbers.  ``predict_lean`` itself would need about 60 GB at 64x64x32 (Pt plus LAPACK's Fortran
copy of it); this script evaluates the same panels in worker processes (one BLAS thread each), keeps Pt in
``/dev/shm`` and applies the triangular solve panel by panel.  ``--check`` compares that arrangement with
``cubing_lean`` on a small cube.

Output ``tests/golden/fullsize_<cfg>.npz``: the settings, the five ``cubing`` inputs in compact form, every
``stride``-th voxel of the six result cubes (stride coprime to the cube edges), max|cube| and sum(cube) of the six full
cubes, log-likelihood, and the measured CPU times (wall, and core-seconds per stage).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

WORKLOADS = {   # bench.py WORKLOADS
    "check": dict(shape=(8, 6, 16), kernel="matern32", nd=5, gl_mult=(1.0, 1.01, 1.02), stride=1, panel=128),
    "cfg2": dict(shape=(32, 32, 32), kernel="exp", nd=0, gl_mult=(1.0, 1.0, 1.0), stride=3, panel=2048),
    "cfg3": dict(shape=(64, 64, 32), kernel="matern32", nd=50, gl_mult=(1.0, 1.01, 1.02), stride=5, panel=2048),
    # the north star's target cube: 64x64x32, two properties (no drill data), squared-exponential kernel (bench workload cfg3e)
    "cfg3e": dict(shape=(64, 64, 32), kernel="exp", nd=0, gl_mult=(1.0, 1.0, 1.0), stride=5, panel=2048),
}

G = {}   # inherited by the forked workers


def _panel_job(ip):
    from threadpoolctl import threadpool_limits
    from oracle import numpy_oracle as o
    c, A, didx, pts, params, w, amp, Ns, nd, M, N, panel = (G[k] for k in ("c", "A", "didx", "pts", "params", "w", "amp", "Ns", "nd", "M", "N", "panel"))
    cols = np.arange(ip * panel, min(N, (ip + 1) * panel))
    timers = {}
    with threadpool_limits(1):
        P = o.pt_panel(c, params, w, amp, A, didx, pts, cols, timers=timers)        # (M, 3, nc)
        t0 = time.perf_counter()
        pt = np.memmap(G["pt_path"], dtype=np.float64, mode="r+", shape=(G["npanel"], M, 3, panel))
        pt[ip, :, :, :len(cols)] = P
        pt.flush()
        del pt
        part = np.zeros((2 * Ns, M))
        part[:Ns] = A[0][:, cols] @ P[:, 0, :].T            # rows of Asens3 in block 0 only touch property 0
        part[Ns:] = A[1][:, cols] @ P[:, 1, :].T
        timers["aka"] = time.perf_counter() - t0
        acc = np.memmap(G["aka_path"] + ".%d" % ip, dtype=np.float64, mode="w+", shape=(2 * Ns, M))
        acc[:] = part
        acc.flush()
        del acc
        drill_rows = {}
        if nd:
            for k, d in enumerate(didx):
                if cols[0] <= d <= cols[-1]:
                    drill_rows[k] = P[:, 2, d - cols[0]].copy()
    return ip, timers, drill_rows


def _solve_job(ip):
    from threadpoolctl import threadpool_limits
    from scipy.linalg import solve_triangular
    M, N, panel = G["M"], G["N"], G["panel"]
    nc = min(N, (ip + 1) * panel) - ip * panel
    with threadpool_limits(1):
        t0 = time.perf_counter()
        pt = np.memmap(G["pt_path"], dtype=np.float64, mode="r", shape=(G["npanel"], M, 3, panel))
        B = np.asfortranarray(np.asarray(pt[ip]).reshape(M, 3 * panel))
        V = solve_triangular(G["L"], B, lower=True, overwrite_b=True, check_finite=False)
        t1 = time.perf_counter()
        mu = (V.T @ G["u"]).reshape(3, panel)[:, :nc]
        var = (G["amp"] - np.einsum("ij,ij->j", V, V)).reshape(3, panel)[:, :nc]
        t2 = time.perf_counter()
    return ip, mu, var, t1 - t0, t2 - t1


def run(name, workers, out_path=None):
    from scipy.linalg import cholesky, solve_triangular
    from geobo_b200 import config_loader, synth
    from oracle import numpy_oracle as o
    wl = WORKLOADS[name]
    xN, yN, zN = wl["shape"]
    cfg = synth.settings(xN, yN, zN, kernelfunc=wl["kernel"])
    config_loader.load_settings(cfg, make_outpath=False)
    c = o.make_config(cfg)
    N, Ns, nd = xN * yN * zN, xN * yN, wl["nd"]
    M = 2 * Ns + nd
    t_start = time.perf_counter()
    Edges, voxelpos = o.cube_geometry(c)
    loc = o.sensor_grid(c)
    assert np.array_equal(loc, synth.sensor_grid())
    density, magsus = synth.cylinders(voxelpos)                 # the truth cubes bench.py uses (host NumPy, no device work)
    t0 = time.perf_counter()
    Ag = o.a_sens(c, c.magneticField * 0.0, loc, Edges, "grav")
    Am = o.a_sens(c, c.magneticField, loc, Edges, "magn")
    t_sens = time.perf_counter() - t0
    grav = (Ag @ density.ravel()).astype(np.float32).astype(np.float64)        # simcube.py:147-150, :196-199
    mag = (Am @ magsus.ravel()).astype(np.float32).astype(np.float64)
    drilldata0 = np.zeros(N)
    if nd:
        idx = np.random.default_rng(0).choice(N, nd, replace=False)
        drilldata0[idx] = density.ravel()[idx]
    drilldata0 = drilldata0.reshape(xN, yN, zN)
    drillfield = drilldata0[drilldata0 != 0]
    gl0 = c.gp_lengthscale * c.xvoxsize * np.asarray(wl["gl_mult"])
    gl, sig, w, amp = o._gp_setup(c, gl0.copy())
    y, stds = o._normalise(c, grav, mag, drillfield)
    didx = o.drill_indices(drilldata0)
    params = o.dedup_lengths(gl)
    pts = o.grid_points((xN, yN, zN), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    panel = wl["panel"]
    npanel = -(-N // panel)
    tag = "%s_%d" % (name, os.getpid())
    pt_path, aka_path = "/dev/shm/geobo_pt_" + tag, "/dev/shm/geobo_aka_" + tag
    np.memmap(pt_path, dtype=np.float64, mode="w+", shape=(npanel, M, 3, panel)).flush()
    G.update(c=c, A=[Ag, Am], didx=didx, pts=pts, params=params, w=w, amp=amp, Ns=Ns, nd=nd, M=M, N=N, panel=panel, npanel=npanel,
             pt_path=pt_path, aka_path=aka_path)
    stages = {}
    AkA = np.zeros((M, M))
    try:
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(workers) as pool:
            for k, (ip, timers, drill_rows) in enumerate(pool.imap_unordered(_panel_job, range(npanel))):
                for key, v in timers.items():
                    stages[key] = stages.get(key, 0.0) + v
                part = np.memmap(aka_path + ".%d" % ip, dtype=np.float64, mode="r", shape=(2 * Ns, M))
                AkA[:2 * Ns] += part
                del part
                os.unlink(aka_path + ".%d" % ip)
                for kd, row in drill_rows.items():
                    AkA[2 * Ns + kd] = row
                if k % 8 == 0:
                    print("  panel %d / %d  (%.0f s)" % (k + 1, npanel, time.perf_counter() - t0), flush=True)
        t_proj = time.perf_counter() - t0
        yerr2 = np.hstack((np.full(Ns, sig[0] ** 2), np.full(Ns, sig[1] ** 2), np.full(nd, sig[2] ** 2)))
        AkA[np.diag_indices(M)] += yerr2
        t0 = time.perf_counter()
        L = cholesky(AkA, lower=True)
        stages["chol"] = time.perf_counter() - t0
        u = solve_triangular(L, y, lower=True)
        logl = -0.5 * (u @ u + np.log(np.diag(L) ** 2).sum() + N * np.log(2 * np.pi))
        G.update(L=L, u=u)
        mu = np.empty((3, N))
        var = np.empty((3, N))
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(workers) as pool:
            for ip, m_, v_, ts, tm in pool.imap_unordered(_solve_job, range(npanel)):
                c0 = ip * panel
                mu[:, c0:c0 + m_.shape[1]] = m_
                var[:, c0:c0 + v_.shape[1]] = v_
                stages["trsm"] = stages.get("trsm", 0.0) + ts
                stages["mean_var"] = stages.get("mean_var", 0.0) + tm
        t_solve = time.perf_counter() - t0
    finally:
        for p in [pt_path] + [aka_path + ".%d" % i for i in range(npanel)]:
            if os.path.exists(p):
                os.unlink(p)
    cubes = o._finish(c, mu.reshape(-1), var.reshape(-1), stds)
    wall = time.perf_counter() - t_start
    result = dict(cubes=cubes, logl=logl, gl_after=gl, cfg=cfg, grav=grav, mag=mag, didx=didx, drillvals=drilldata0.ravel()[didx],
                  gl0=gl0, stages=stages, wall=wall, t_sens=t_sens, t_proj=t_proj, t_solve=t_solve, workers=workers,
                  inputs=(grav, mag, drillfield, loc, drilldata0), c=c)
    if out_path:
        stride = wl["stride"]
        names = ["density_rec", "magsus_rec", "drill_rec", "density_var", "magsus_var", "drill_var"]
        save = dict(cfg=json.dumps(cfg), workload=name, gl0=gl0, gl_after=np.asarray(gl), grav=grav, mag=mag, didx=didx,
                    drillvals=result["drillvals"], logl=logl, stride=stride,
                    cpu=json.dumps(dict(host_cores=os.cpu_count(), workers=workers, blas_threads_per_worker=1, wall_s=wall,
                                        a_sens_s=t_sens, projection_aka_wall_s=t_proj, solve_wall_s=t_solve,
                                        core_seconds_per_stage=stages, voxels_per_s_wall=N / (wall - t_sens),
                                        voxels_per_s_one_core=N / sum(stages.values()))))
        for n, cube in zip(names, cubes):
            flat = np.asarray(cube).ravel()
            save["sub_" + n] = flat[::stride].copy()
            with np.errstate(all="ignore"):
                save["max_" + n] = np.nanmax(np.abs(flat)) if not np.isnan(flat).all() else np.nan
                save["sum_" + n] = np.nansum(flat) if not np.isnan(flat).all() else np.nan
        np.savez_compressed(out_path, **save)
        print("wrote %s (%.1f kB); wall %.1f s with %d workers; core-seconds %s"
              % (out_path, os.path.getsize(out_path) / 1e3, wall, workers, {k: round(v, 1) for k, v in stages.items()}))
    return result


def check(workers):
    """The panel / worker arrangement of this script against ``cubing_lean`` on a small cube."""
    from oracle import numpy_oracle as o
    r = run("check", workers)
    grav, mag, drillfield, loc, drilldata0 = r["inputs"]
    ref, ex = o.cubing_lean(r["c"], grav, mag, drillfield, loc, drilldata0, gp_length=r["gl0"].copy())
    worst = max(float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(r["cubes"], ref))
    print("check: worst norm-wise difference to cubing_lean %.2e, logl %.12g vs %.12g" % (worst, r["logl"], ex["logl"]))
    assert worst < 1e-11 and abs(r["logl"] - ex["logl"]) < 1e-9 * abs(ex["logl"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=sorted(WORKLOADS))
    ap.add_argument("--workers", type=int, default=max(1, (os.cpu_count() or 2) - 2))
    args = ap.parse_args()
    if args.workload == "check":
        check(args.workers)
    else:
        run(args.workload, args.workers, os.path.join(HERE, "fullsize_%s.npz" % args.workload))


if __name__ == "__main__":
    main()
