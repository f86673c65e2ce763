#!/usr/bin/env python
"""Full-size CPU cross-check of the structured restatements against the committed full-size fixtures.

    python tests/golden/check_fullsize_structured.py cfg2 kron  [--workers 7]     # 32x32x32, exp       (about a minute)
    python tests/golden/check_fullsize_structured.py cfg2 fft
    python tests/golden/check_fullsize_structured.py cfg3 fft                     # 64x64x32, matern32  (tens of minutes)
    python tests/golden/check_fullsize_structured.py cfg2 kron-host               # Pt from csrc/kron.cuh compiled for the host

Build-container job (CPU only).  ``fullsize_<cfg>.npz`` holds the result of the oracle's DENSE lean path
(``make_fullsize_golden.py``: every covariance entry evaluated, dgemm projection).  This script recomputes the same inversion
with ``Pt = Asens3 . kcov`` from the oracle's STRUCTURED restatements (``oracle/kron.py``: Kronecker mode products;
``oracle/fftconv.py``: circulant embedding) -- a different algorithm for the dominant stage -- and everything downstream
(AkA, Cholesky, solves, mean, variance) as in ``predict_lean``, then compares with the fixture: two independent CPU routes to
the same cubes at the size the bench is quoted on.  The measured differences go to ``profiles/r1_fullsize_structured_oracle.json``.
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
sys.path.insert(0, HERE)

import make_fullsize_golden as mk  # noqa: E402

G = mk.G
CUBES = ["density_rec", "magsus_rec", "drill_rec", "density_var", "magsus_var", "drill_var"]


def _rows_job(job):
    """Rows [s0, s1) of data block cb of Pt for the three property blocks, through the structured restatement."""
    from threadpoolctl import threadpool_limits
    cb, s0, s1 = job
    c, A, params, w, amp, Ns, M, N, panel, npanel = (G[k] for k in ("c", "A", "params", "w", "amp", "Ns", "M", "N", "panel", "npanel"))
    t0 = time.perf_counter()
    with threadpool_limits(1):
        pt = np.memmap(G["pt_path"], dtype=np.float64, mode="r+", shape=(npanel, M, 3, panel))
        X = A[cb][s0:s1]
        host_out = _host_rows(cb, X) if G["structure"].endswith("-host") else None
        for r in range(3):
            if host_out is not None:
                out = host_out[:, r, :]
            elif G["structure"] == "kron":
                from oracle import kron as kr
                out = kr.apply_block(c, params, w, amp, cb, r, X)
            else:
                out = _fft_rows(cb, r, X)
            for ip in range(npanel):
                nc = min(N, (ip + 1) * panel) - ip * panel
                pt[ip, cb * Ns + s0:cb * Ns + s1, r, :nc] = out[:, ip * panel:ip * panel + nc]
        pt.flush()
        del pt
    return time.perf_counter() - t0


def _host_rows(cb, X):
    """Rows of Pt from the DEVICE SOURCE compiled for the host (tests/host_harness/kron_host.cpp / fftconv_host.cpp: the per-thread
    code of csrc/kron.cuh / csrc/fftconv.cuh driven with the launch geometry of the kernels) -> (rows, 3, N)."""
    import ctypes
    c, params, w, amp, N = (G[k] for k in ("c", "params", "w", "amp", "N"))
    lib = ctypes.CDLL(G["host_so"])
    P, L, D, I = ctypes.c_void_p, ctypes.c_long, ctypes.c_double, ctypes.c_int
    p = lambda a: a.ctypes.data_as(P)                                            # noqa: E731
    l = np.ascontiguousarray(params, dtype=float)
    ww = np.ascontiguousarray(w, dtype=float)
    ncube = np.array([c.xNcube, c.yNcube, c.zNcube], dtype=np.int64)
    vox = np.array([c.xvoxsize, c.yvoxsize, c.zvoxsize], dtype=float)
    X = np.ascontiguousarray(X, dtype=float)
    out = np.full((X.shape[0], 3 * N), np.nan)
    kid = {"sparse": 0, "exp": 1, "matern32": 2}[c.kernelfunc]
    if G["structure"] == "kron-host":
        lib.kron_host_apply.argtypes = [I, P, P, D, P, P, I, P, L, L, L, L, L, P, L, L, I]
        lib.kron_host_apply.restype = None
        lib.kron_host_apply(kid, p(l), p(ww), amp, p(ncube), p(vox), cb * 3, p(X), N, X.shape[0], 0, N, X.shape[0], p(out), 3 * N, N, 0)
    else:
        pad = np.zeros(3, dtype=np.int32)
        lib.fftconv_host_apply.argtypes = [I, P, P, D, P, P, I, P, L, L, L, L, L, P, L, L, I, P]
        lib.fftconv_host_apply.restype = None
        lib.fftconv_host_apply(kid, p(l), p(ww), amp, p(ncube), p(vox), cb * 3, p(X), N, X.shape[0], 0, N, 2, p(out), 3 * N, N, 0, p(pad))
    assert np.isfinite(out).all()
    return out.reshape(X.shape[0], 3, N)


def _fft_rows(cb, r, X):
    """oracle/fftconv.apply_block with the spectrum of the block computed once in the parent."""
    c = G["c"]
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    S = G["spectra"][(cb, r)]
    P = S.shape
    src = np.asarray(X, dtype=float).reshape(-1, yN, xN, zN)
    pad = np.zeros((src.shape[0],) + P)
    pad[:, :yN, :xN, :zN] = src
    out = np.fft.ifftn(np.fft.fftn(pad, axes=(1, 2, 3)) * S, axes=(1, 2, 3)).real[:, :yN, :xN, :zN]
    return out.reshape(X.shape)


def _aka_job(ip):
    """Drill rows of Pt for column panel ip (dense gathers, numpy_oracle.pt_panel's drill part) and the panel's part of AkA."""
    from threadpoolctl import threadpool_limits
    from oracle import numpy_oracle as o
    c, A, didx, pts, params, w, amp, Ns, nd, M, N, panel = (G[k] for k in ("c", "A", "didx", "pts", "params", "w", "amp", "Ns", "nd", "M", "N", "panel"))
    cols = np.arange(ip * panel, min(N, (ip + 1) * panel))
    with threadpool_limits(1):
        pt = np.memmap(G["pt_path"], dtype=np.float64, mode="r+", shape=(G["npanel"], M, 3, panel))
        if nd:
            acc = 0
            for d in range(3):
                delta = pts[cols, d][None, :] - pts[didx, d][:, None]
                acc = acc + delta ** 2
            for r in range(3):
                pt[ip, 2 * Ns:, r, :len(cols)] = amp * o.cov_block(acc, params, w, c.kernelfunc, 2, r)
            pt.flush()
        P = np.asarray(pt[ip])[:, :, :len(cols)]
        part = np.zeros((2 * Ns, M))
        part[:Ns] = A[0][:, cols] @ P[:, 0, :].T
        part[Ns:] = A[1][:, cols] @ P[:, 1, :].T
        drill_rows = {}
        if nd:
            for k, d in enumerate(didx):
                if cols[0] <= d <= cols[-1]:
                    drill_rows[k] = P[:, 2, d - cols[0]].copy()
        del pt
    return ip, part, drill_rows


def run(name, structure, workers, rows_per_job):
    from scipy.linalg import cholesky, solve_triangular
    from geobo_b200 import config_loader
    from oracle import fftconv as fc
    from oracle import numpy_oracle as o
    g = np.load(os.path.join(HERE, "fullsize_%s.npz" % name))
    wl = mk.WORKLOADS[name]
    cfg = json.loads(str(g["cfg"]))
    config_loader.load_settings(cfg, make_outpath=False)
    c = o.make_config(cfg)
    xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
    N, Ns = xN * yN * zN, xN * yN
    didx = np.asarray(g["didx"])
    nd = didx.size
    M = 2 * Ns + nd
    t_start = time.perf_counter()
    Edges, _ = o.cube_geometry(c)
    loc = o.sensor_grid(c)
    Ag = o.a_sens(c, c.magneticField * 0.0, loc, Edges, "grav")
    Am = o.a_sens(c, c.magneticField, loc, Edges, "magn")
    print("  a_sens %.0f s" % (time.perf_counter() - t_start), flush=True)
    d0 = np.zeros(N)
    d0[didx] = g["drillvals"]
    drillfield = d0.reshape(xN, yN, zN)[d0.reshape(xN, yN, zN) != 0]
    gl, sig, w, amp = o._gp_setup(c, np.array(g["gl0"], dtype=float))
    y, stds = o._normalise(c, g["grav"], g["mag"], drillfield)
    params = o.dedup_lengths(gl)
    pts = o.grid_points((xN, yN, zN), (c.xvoxsize, c.yvoxsize, c.zvoxsize))
    panel = wl["panel"]
    npanel = -(-N // panel)
    pt_path = "/dev/shm/geobo_pts_%s_%d" % (name, os.getpid())
    np.memmap(pt_path, dtype=np.float64, mode="w+", shape=(npanel, M, 3, panel)).flush()
    spectra = {}
    host_so = None
    if structure.endswith("-host"):
        import subprocess
        import tempfile
        src = os.path.join(ROOT, "tests", "host_harness", "kron_host.cpp" if structure == "kron-host" else "fftconv_host.cpp")
        host_so = os.path.join(tempfile.mkdtemp(prefix="geobo_host_"), "harness.so")
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", host_so], check=True)
    if structure == "fft":
        spectra = {(cb, r): fc.spectrum(c, params, w, amp, c.kernelfunc, cb, r) for cb in range(2) for r in range(3)}
    G.update(c=c, A=[Ag, Am], didx=didx, pts=pts, params=params, w=w, amp=amp, Ns=Ns, nd=nd, M=M, N=N, panel=panel, npanel=npanel,
             pt_path=pt_path, structure=structure, spectra=spectra, host_so=host_so)
    AkA = np.zeros((M, M))
    try:
        t0 = time.perf_counter()
        jobs = [(cb, s0, min(Ns, s0 + rows_per_job)) for cb in range(2) for s0 in range(0, Ns, rows_per_job)]
        with mp.get_context("fork").Pool(workers) as pool:
            core_s = sum(pool.imap_unordered(_rows_job, jobs))
        t_proj = time.perf_counter() - t0
        print("  structured projection (%s): %.0f s wall, %.0f core-seconds" % (structure, t_proj, core_s), flush=True)
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(workers) as pool:
            for ip, part, drill_rows in pool.imap_unordered(_aka_job, range(npanel)):
                AkA[:2 * Ns] += part
                for kd, row in drill_rows.items():
                    AkA[2 * Ns + kd] = row
        t_aka = time.perf_counter() - t0
        yerr2 = np.hstack((np.full(Ns, sig[0] ** 2), np.full(Ns, sig[1] ** 2), np.full(nd, sig[2] ** 2)))
        AkA[np.diag_indices(M)] += yerr2
        L = cholesky(AkA, lower=True)
        u = solve_triangular(L, y, lower=True)
        logl = -0.5 * (u @ u + np.log(np.diag(L) ** 2).sum() + N * np.log(2 * np.pi))
        G.update(L=L, u=u)
        mu = np.empty((3, N))
        var = np.empty((3, N))
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(workers) as pool:
            for ip, m_, v_, ts, tm in pool.imap_unordered(mk._solve_job, range(npanel)):
                c0 = ip * panel
                mu[:, c0:c0 + m_.shape[1]] = m_
                var[:, c0:c0 + v_.shape[1]] = v_
        t_solve = time.perf_counter() - t0
    finally:
        if os.path.exists(pt_path):
            os.unlink(pt_path)
    cubes = o._finish(c, mu.reshape(-1), var.reshape(-1), stds)
    stride = int(g["stride"])
    errs = {}
    for n, cube in zip(CUBES, cubes):
        sub, ref_max = g["sub_" + n], float(g["max_" + n])
        got = np.asarray(cube).ravel()
        if np.isnan(sub).all():
            assert np.isnan(got).all(), n
            continue
        errs[n] = float(np.abs(got[::stride] - sub).max() / ref_max)
        errs[n + "_sum"] = abs(float(got.sum()) - float(g["sum_" + n])) / (N * ref_max)
    errs["logl_rel"] = abs(logl - float(g["logl"])) / abs(float(g["logl"]))
    res = dict(workload=name, structure=structure, voxels=N, M=M, errors_vs_dense_fixture=errs, worst=max(v for k, v in errs.items() if k != "logl_rel"),
               wall_s=time.perf_counter() - t_start, projection_wall_s=t_proj, projection_core_s=core_s, aka_wall_s=t_aka, solve_wall_s=t_solve,
               workers=workers, dense_fixture_projection_core_s=json.loads(str(g["cpu"]))["core_seconds_per_stage"])
    print(json.dumps(res, indent=1))
    out = os.path.join(ROOT, "profiles", "r1_fullsize_structured_oracle.json")
    allr = json.load(open(out)) if os.path.exists(out) else {}
    allr["%s_%s" % (name, structure)] = res
    json.dump(allr, open(out, "w"), indent=1)
    assert res["worst"] < 1e-5, errs
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["cfg2", "cfg3", "cfg3e"])
    ap.add_argument("structure", choices=["kron", "fft", "kron-host", "fft-host"],
                    help="kron / fft: the NumPy restatements; kron-host / fft-host: the device sources compiled for the host")
    ap.add_argument("--workers", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--rows-per-job", type=int, default=16)
    args = ap.parse_args()
    run(args.workload, args.structure, args.workers, args.rows_per_job)


if __name__ == "__main__":
    main()
