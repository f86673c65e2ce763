#!/usr/bin/env python
"""Full-size CPU result of the oracle for BASELINE config 4 (96x96x48 cube, 442 368 voxels, two properties, exp kernel, M = 18 432).

    python tests/golden/make_fullsize_cfg4.py --check            # the arrangement of this script against cubing_lean on a small cube
    python tests/golden/make_fullsize_cfg4.py [--workers 7] [--scratch /tmp/geobo_cfg4]      # about 2 h on 8 cores, 66 GB of disk

Build-container job (CPU only).  The dense lean path of ``make_fullsize_golden.py`` would need ~60 h here (2.2e16 flops) and
196 GB for Pt, so this generator goes through the oracle's Kronecker restatement of the exp blocks (``oracle/kron.py``, equal to the
dense oracle to 1.3e-11 at 64x64x32, ``profiles/r1_fullsize_structured_oracle.json``) and never stores Pt or V:

  0. sensitivities by the oracle's ``a_sens`` in worker processes -> two float64 files on disk (2 x 32.6 GB); surveys = A . truth cube
     rounded through float32 (the bench's synthetic recipe, ``geobo_b200/synth.py``);
  1. AkA row chunk by row chunk:  AkA[(c, S), (c', :)] = (A_c[S] . K_cc') . A_c'^T   (mode products, then one streamed dgemm);
  2. Cholesky, u = L^-1 y, logl, alpha = L^-T u, Linv = L^-1;
  3. mean  mu_r = sum_c K_rc (A_c^T alpha_c);
  4. variance  var_r = amp - colsumsq(V_r),  V_r[m-chunk] = sum_c (Linv[m-chunk, c-block] . A_c) . K_cr  -- rows of V are produced
     chunk by chunk and squared on the fly (V = L^-1 Pt = (L^-1 A3) K: the mode products commute with the row operation).

Output ``tests/golden/fullsize_cfg4.npz`` in the format of the other full-size fixtures (inputs, every ``stride``-th voxel of the six
cubes, max / sum over the full cubes, logl, CPU timings).
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

WORKLOADS = {
    "check": dict(shape=(8, 6, 16), kernel="exp", nd=0, stride=1, rows=24, mrows=40),
    "cfg4": dict(shape=(96, 96, 48), kernel="exp", nd=0, stride=97, rows=2304, mrows=2304),
}
G = {}


def _sens_job(job):
    from threadpoolctl import threadpool_limits
    from oracle import numpy_oracle as o
    kind, s0, s1 = job
    c, loc, E, N, Ns = (G[k] for k in ("c", "loc", "E", "N", "Ns"))
    with threadpool_limits(1):
        B = c.magneticField * 0.0 if kind == 0 else c.magneticField
        rows = o.a_sens(c, B, loc, E, "grav" if kind == 0 else "magn", sensors=list(range(s0, s1)))
        mm = np.memmap(G["a_path"][kind], dtype=np.float64, mode="r+", shape=(Ns, N))
        mm[s0:s1] = rows
        mm.flush()
        del mm
        truth = G["truth"][kind]
        return kind, s0, s1, rows @ truth


def toeplitz(line, n):
    a = np.arange(n)
    return np.ascontiguousarray(line[a[:, None] - a[None, :] + n - 1])


class KronBlocks:
    """``oracle.kron.apply_block`` with the three mode products as BLAS matmuls (same factor lines, same order y, z, x)."""

    def __init__(self, c, params, w, amp):
        from oracle import kron as kr
        self.c = c
        self.f = {}
        for cb in range(3):
            for r in range(3):
                t0, fy, fx, fz = kr.factor_lines(c, params, w, amp, cb, r)
                if t0 == 0.0:
                    self.f[(cb, r)] = None
                else:
                    self.f[(cb, r)] = (toeplitz(fy / t0 / t0, c.yNcube), toeplitz(fx, c.xNcube), toeplitz(fz, c.zNcube))

    def apply(self, cb, r, X):
        c = self.c
        xN, yN, zN = c.xNcube, c.yNcube, c.zNcube
        if self.f[(cb, r)] is None:
            return np.zeros_like(X)
        Ty, Tx, Tz = self.f[(cb, r)]
        s = X.shape[0]
        T = np.matmul(Ty, X.reshape(s, yN, xN * zN))                         # y mode: out[s, a, q] = sum_b Ty[a, b] T[s, b, q]
        T = np.matmul(T.reshape(s * yN * xN, zN), Tz.T)                      # z mode: out[.., a] = sum_b Tz[a, b] T[.., b]
        T = np.matmul(Tx, T.reshape(s * yN, xN, zN))                         # x mode: out[sy, a, z] = sum_b Tx[a, b] T[sy, b, z]
        return T.reshape(s, -1)


def streamed_abt(P, A_mm, blk=768):
    """P (rows, N) times A^T for a (Ns, N) memmap, A streamed in row blocks."""
    Ns = A_mm.shape[0]
    out = np.empty((P.shape[0], Ns))
    for b0 in range(0, Ns, blk):
        out[:, b0:b0 + blk] = P @ np.asarray(A_mm[b0:b0 + blk]).T
    return out


def streamed_la(Lblk, A_mm, blk=768):
    """Lblk (rows, Ns) times A for a (Ns, N) memmap, A streamed in row blocks."""
    Ns, N = A_mm.shape
    out = np.zeros((Lblk.shape[0], N))
    for b0 in range(0, Ns, blk):
        out += Lblk[:, b0:b0 + blk] @ np.asarray(A_mm[b0:b0 + blk])
    return out


def run(name, workers, scratch, out_path=None):
    from scipy.linalg import cholesky, solve_triangular
    from geobo_b200 import config_loader, synth
    from oracle import numpy_oracle as o
    wl = WORKLOADS[name]
    xN, yN, zN = wl["shape"]
    cfg = synth.settings(xN, yN, zN, kernelfunc=wl["kernel"])
    config_loader.load_settings(cfg, make_outpath=False)
    c = o.make_config(cfg)
    N, Ns, nd = xN * yN * zN, xN * yN, wl["nd"]
    assert nd == 0, "two-property workload (no drill rows)"
    M = 2 * Ns
    os.makedirs(scratch, exist_ok=True)
    t_start = time.perf_counter()
    E, voxelpos = o.cube_geometry(c)
    loc = o.sensor_grid(c)
    density, magsus = synth.cylinders(voxelpos)
    a_path = [os.path.join(scratch, "A_%s_%d.f64" % (name, k)) for k in range(2)]
    for p in a_path:
        np.memmap(p, dtype=np.float64, mode="w+", shape=(Ns, N)).flush()
    G.update(c=c, loc=loc, E=E, N=N, Ns=Ns, a_path=a_path, truth=[density.ravel(), magsus.ravel()])
    stages = {}
    # ---- 0. sensitivities + surveys
    t0 = time.perf_counter()
    surveys = [np.empty(Ns), np.empty(Ns)]
    step = max(1, min(64, Ns // (4 * workers)))
    jobs = [(k, s0, min(Ns, s0 + step)) for k in range(2) for s0 in range(0, Ns, step)]
    with mp.get_context("fork").Pool(workers) as pool:
        for k, s0, s1, vals in pool.imap_unordered(_sens_job, jobs):
            surveys[k][s0:s1] = vals
    t_sens = time.perf_counter() - t0
    print("sensitivities: %.0f s" % t_sens, flush=True)
    grav = surveys[0].astype(np.float32).astype(np.float64)        # simcube.py:147-150, :196-199
    mag = surveys[1].astype(np.float32).astype(np.float64)
    drilldata0 = np.zeros((xN, yN, zN))
    drillfield = drilldata0[drilldata0 != 0]
    gl0 = c.gp_lengthscale * c.xvoxsize * np.ones(3)
    gl, sig, w, amp = o._gp_setup(c, gl0.copy())
    with np.errstate(all="ignore"):
        y, stds = o._normalise(c, grav, mag, drillfield)
    params = o.dedup_lengths(gl)
    KB = KronBlocks(c, params, w, amp)
    A = [np.memmap(p, dtype=np.float64, mode="r", shape=(Ns, N)) for p in a_path]
    # ---- 1. AkA (lower block triangle; block (0, 1) by symmetry)
    t0 = time.perf_counter()
    AkA = np.zeros((M, M))
    R = wl["rows"]
    for cb, cp in ((0, 0), (1, 0), (1, 1)):
        for s0 in range(0, Ns, R):
            X = np.asarray(A[cb][s0:s0 + R])
            P = KB.apply(cb, cp, X)                                 # Pt[(cb, S), (cp, :)]
            AkA[cb * Ns + s0:cb * Ns + s0 + X.shape[0], cp * Ns:(cp + 1) * Ns] = streamed_abt(P, A[cp])
            print("  AkA block (%d,%d) rows %d / %d  (%.0f s)" % (cb, cp, s0 + X.shape[0], Ns, time.perf_counter() - t0), flush=True)
    AkA[:Ns, Ns:] = AkA[Ns:, :Ns].T
    AkA = 0.5 * (AkA + AkA.T)          # the two routes to a symmetric pair differ by rounding only
    AkA[np.diag_indices(M)] += np.hstack((np.full(Ns, sig[0] ** 2), np.full(Ns, sig[1] ** 2)))
    stages["aka_and_projection"] = time.perf_counter() - t0
    # ---- 2. Cholesky, u, logl, alpha, Linv
    t0 = time.perf_counter()
    L = cholesky(AkA, lower=True)
    del AkA
    u = solve_triangular(L, y, lower=True)
    logl = -0.5 * (u @ u + np.log(np.diag(L) ** 2).sum() + N * np.log(2 * np.pi))
    alpha = solve_triangular(L, u, lower=True, trans="T")
    Linv = solve_triangular(L, np.eye(M), lower=True, overwrite_b=True, check_finite=False)
    del L
    stages["chol_and_inverse"] = time.perf_counter() - t0
    print("Cholesky + inverse: %.0f s, logl %.9f" % (stages["chol_and_inverse"], logl), flush=True)
    # ---- 3. mean
    t0 = time.perf_counter()
    wv = []
    for cb in range(2):
        acc = np.zeros(N)
        for b0 in range(0, Ns, 768):
            b1 = min(Ns, b0 + 768)
            acc += alpha[cb * Ns + b0:cb * Ns + b1] @ np.asarray(A[cb][b0:b1])
        wv.append(acc)
    mu = np.full((3, N), np.nan)
    for r in range(2):
        mu[r] = sum(KB.apply(cb, r, wv[cb][None, :])[0] for cb in range(2))
    mu[2] = 0.0                                                             # scaled by NaN below (no drill data, Q9)
    stages["mean"] = time.perf_counter() - t0
    # ---- 4. variance
    t0 = time.perf_counter()
    ss = np.zeros((3, N))
    Rm = wl["mrows"]
    for m0 in range(0, M, Rm):
        Gc = [streamed_la(Linv[m0:m0 + Rm, cb * Ns:(cb + 1) * Ns], A[cb]) for cb in range(2)]
        for r in range(2):                                                  # the drill property is NaN without drill data (Q9)
            V = KB.apply(0, r, Gc[0])
            V += KB.apply(1, r, Gc[1])
            ss[r] += np.einsum("ij,ij->j", V, V)
            del V
        del Gc
        print("  variance rows %d / %d  (%.0f s)" % (min(M, m0 + Rm), M, time.perf_counter() - t0), flush=True)
    var = amp - ss
    stages["variance"] = time.perf_counter() - t0
    cubes = o._finish(c, mu.reshape(-1), var.reshape(-1), stds)
    wall = time.perf_counter() - t_start
    result = dict(cubes=cubes, logl=logl, c=c, gl0=gl0, inputs=(grav, mag, drillfield, loc, drilldata0))
    for p in a_path:
        os.unlink(p)
    if out_path:
        stride = wl["stride"]
        names = ["density_rec", "magsus_rec", "drill_rec", "density_var", "magsus_var", "drill_var"]
        didx = np.zeros(0, dtype=np.int64)
        save = dict(cfg=json.dumps(cfg), workload=name, gl0=gl0, gl_after=np.asarray(gl), grav=grav, mag=mag, didx=didx,
                    drillvals=np.zeros(0), logl=logl, stride=stride,
                    cpu=json.dumps(dict(host_cores=os.cpu_count(), workers=workers, route="oracle/kron.py mode products, two passes, nothing of "
                                        "size M x 3N stored (tests/golden/make_fullsize_cfg4.py)", wall_s=wall, a_sens_s=t_sens,
                                        core_seconds_per_stage={}, stage_wall_s=stages)))
        for n, cube in zip(names, cubes):
            flat = np.asarray(cube).ravel()
            save["sub_" + n] = flat[::stride].copy()
            with np.errstate(all="ignore"):
                save["max_" + n] = np.nanmax(np.abs(flat)) if not np.isnan(flat).all() else np.nan
                save["sum_" + n] = np.nansum(flat) if not np.isnan(flat).all() else np.nan
        np.savez_compressed(out_path, **save)
        print("wrote %s (%.1f kB); wall %.0f s; stages %s" % (out_path, os.path.getsize(out_path) / 1e3, wall, {k: round(v) for k, v in stages.items()}))
    return result


def check(workers, scratch):
    from oracle import numpy_oracle as o
    r = run("check", workers, scratch)
    grav, mag, drillfield, loc, drilldata0 = r["inputs"]
    with np.errstate(all="ignore"):
        ref, ex = o.cubing_lean(r["c"], grav, mag, drillfield, loc, drilldata0, gp_length=r["gl0"].copy())
    worst = 0.0
    for a, b in zip(r["cubes"], ref):
        if np.isnan(b).all():
            assert np.isnan(a).all()
            continue
        worst = max(worst, float(np.abs(a - b).max() / np.abs(b).max()))
    print("check: worst norm-wise difference to cubing_lean %.2e, logl %.12g vs %.12g" % (worst, r["logl"], ex["logl"]))
    assert worst < 1e-9 and abs(r["logl"] - ex["logl"]) < 1e-9 * abs(ex["logl"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--workers", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--scratch", default="/tmp/geobo_cfg4")
    args = ap.parse_args()
    if args.check:
        check(args.workers, args.scratch)
    else:
        run("cfg4", args.workers, args.scratch, os.path.join(HERE, "fullsize_cfg4.npz"))


if __name__ == "__main__":
    main()
