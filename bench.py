#!/usr/bin/env python
"""Benchmark of the GeoBO joint-inversion hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

Metric (BASELINE.json): voxels/sec of the joint inversion = N_voxels / time(covariance assembly +
projection A.K.A^T + Cholesky + triangular solves + posterior mean + variance diagonal) -- the
``Inversion.predict3`` stage.  A "step" is one such inversion of the named synthetic cube.

  value     device-resident inputs (sensitivities, data vector in HBM), timed per step with CUDA events on the
            library's stream and cross-checked by host wall clock between synchronisation points; max over ranks.
  e2e       the same metric through the public Python API, ``Inversion.cubing(...)`` with HOST NumPy arrays:
            includes building the device problem (sensitivity matrices on the GPU), every H2D copy and the D2H
            of the six result cubes.
  roofline  the dominant kernel: fused covariance generation + projection Pt = A.K, by default exact int8 digit
            products on the tcgen05 tensor cores (--precision int8x5), or the fp64 DMMA tensor pipe (--precision fp64).
  cpu_baseline  the oracle's lean NumPy/SciPy restatement of the reference on this box's host cores (bounded sample).

For N > 1 (torchrun) the voxel columns of Pt = A.K are sharded over ranks; the only data-path collective
is the NCCL all-reduce of AkA inside the library (strong scaling: the cube is fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# The CPU legs use every host core for BLAS / LAPACK whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1, which
# would silently halve the reference arm): set the thread count BEFORE NumPy loads its BLAS, and again at run time through
# threadpoolctl (cpu_threads()).
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(host_cores())

import numpy as np  # noqa: E402

# BASELINE.json configs.  cfg3 is the default: the 64x64x32 cube the north star quotes its target on
# ("64x64x32 cube, grav+mag + 50 drillcore constraints, Matern-3/2 cross-cov", configs[2]) -- the largest configuration that
# fits one GPU (59 GB); cfg3e is its two-property sqexp variant, cfg2 = configs[1] (32x32x32), cfg4 = configs[3].
WORKLOADS = {
    "cfg1": dict(shape=(25, 16, 16), kernel="sparse", nd=256, name="25x16x16 example-1 cube (sparse, nd=256)"),
    "cfg1b": dict(shape=(16, 16, 16), kernel="exp", nd=50, name="16x16x16 cube (exp, nd=50)"),
    "cfg2": dict(shape=(32, 32, 32), kernel="exp", nd=0, name="32x32x32 cube, grav+mag joint inversion, sqexp kernel, fp64"),
    "cfg3": dict(shape=(64, 64, 32), kernel="matern32", nd=50, gl_mult=(1.0, 1.01, 1.02),
                 name="64x64x32 cube, grav+mag + 50 drill constraints, Matern-3/2 cross-cov"),
    "cfg3e": dict(shape=(64, 64, 32), kernel="exp", nd=0, name="64x64x32 two-property cube, sqexp"),
    "cfg4": dict(shape=(96, 96, 48), kernel="exp", nd=0, name="96x96x48 cube, 2-property joint inversion"),
    "cfg5": dict(shape=(128, 128, 64), kernel="exp", nd=50, name="128x128x64 cube, 3 cross-correlated properties + BO acquisition sweep"),
}
METRIC = "voxels/sec joint-inversion (cov+chol+solve)"


def read_peaks():
    peaks = {}
    for name in ("MEASURED_PEAKS.json", os.path.join("profiles", "fp64_peaks_r1.json"), os.path.join("profiles", "int8_peaks_r1.json")):
        p = os.path.join(ROOT, name)
        if os.path.exists(p):
            try:
                peaks.update(json.load(open(p)))
            except Exception:
                pass
    return peaks


class ClockSampler:
    """SM clocks / throttle reasons / power sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is
    polled in-process every 50 ms (no process spawn inside or next to the timed region: starting `nvidia-smi` takes
    NVML/driver locks for hundreds of ms on a multi-GPU box and stalls the CUDA calls of every rank); `nvidia-smi
    -lms` is the fallback when the NVML bindings are missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    _nvml = None

    @classmethod
    def init(cls):
        """Call once before any timed region."""
        if cls._nvml is None:
            try:
                import pynvml
                pynvml.nvmlInit()
                cls._nvml = pynvml
            except Exception:
                cls._nvml = False
        return cls._nvml

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.device])
            except Exception:
                pass
        return self.device

    def start(self):
        nv = self.init()
        if nv:
            try:
                self.handle = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
                self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
                self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
                self.thread.start()
                return
            except Exception:
                pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self._nvml
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.rows.append([str(self.device), sm, self.max_sm, pw, ""] + ["Active" if mask & bit else "Not Active" for _, bit in names])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        elif self.thread:
            self.thread.join(timeout=1)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if str(v).lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_threads():
    """All host cores for BLAS / LAPACK, independent of the launcher's environment; returns the thread count in effect."""
    n = host_cores()
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        return int(max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] + [1]))
    except Exception:
        return n


def workload_config(wl, workload_id):
    """The `config` dict of BOTH arms (identical by construction: the driver compares them)."""
    xN, yN, zN = wl["shape"]
    N, Ns, nd = xN * yN * zN, xN * yN, wl["nd"]
    return {"workload": wl["name"], "workload_id": workload_id, "voxels": N, "sensors_per_survey": Ns, "drill_rows": nd,
            "data_rows_M": 2 * Ns + nd, "kernel": wl["kernel"],
            "l2": "no L2 flush between steps: every step streams operands far larger than the 126 MB L2 (sensitivities + Pt: "
                  "%.1f GB fp64)" % ((2 * Ns * N + (2 * Ns + nd) * 3 * N) * 8 / 1e9)}


def load_fixture(workload_id):
    """tests/golden/fullsize_<id>.npz: the bench's own synthetic inputs for this workload and the result of ONE complete CPU-oracle
    inversion of them (sub-sampled cubes, full-cube aggregates, logl) -- the reference the in-run `parity` is computed against."""
    p = os.path.join(ROOT, "tests", "golden", "fullsize_%s.npz" % workload_id)
    return np.load(p) if os.path.exists(p) else None


def fixture_inputs(g, shape):
    xN, yN, zN = shape
    d0 = np.zeros(xN * yN * zN)
    d0[g["didx"]] = g["drillvals"]
    d0 = d0.reshape(xN, yN, zN)
    return dict(grav=np.array(g["grav"], dtype=float), mag=np.array(g["mag"], dtype=float), drillfield=d0[d0 != 0], drilldata0=d0)


CUBE_NAMES = ["density_rec", "magsus_rec", "drill_rec", "density_var", "magsus_var", "drill_var"]


def parity_vs_fixture(g, cubes, logl, N):
    """Norm-wise errors (max|delta| / max|ref| per cube, the norm of the north star's 1e-5) of the six cubes of THIS run against
    the full CPU-oracle inversion stored in the fixture: every stride-th voxel, plus max and sum over the full cubes."""
    stride = int(g["stride"])
    errs = {}
    for n, cube in zip(CUBE_NAMES, cubes):
        sub, ref_max = g["sub_" + n], float(g["max_" + n])
        got = np.asarray(cube).ravel()
        if np.isnan(sub).all():
            errs[n] = 0.0 if np.isnan(got).all() else float("inf")
            continue
        errs[n] = max(float(np.abs(got[::stride] - sub).max() / ref_max), abs(float(np.abs(got).max()) - ref_max) / ref_max,
                      abs(float(got.sum()) - float(g["sum_" + n])) / (N * ref_max))
    mean_err = max(errs[n] for n in CUBE_NAMES[:3])
    var_err = max(errs[n] for n in CUBE_NAMES[3:])
    return {"max_err": max(mean_err, var_err), "mean_err": mean_err, "var_err": var_err,
            "logl_rel_err": abs(logl - float(g["logl"])) / abs(float(g["logl"])), "per_cube": errs, "tolerance": 1e-5,
            "norm": "max|delta| / max|ref| per cube (every %d-th voxel + full-cube max and sum)" % stride,
            "reference": "tests/golden/fullsize_*.npz: one complete CPU-oracle inversion of the same inputs (build container)"}


def traffic_from_profiles(key, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the `ncu --set full` capture of this
    workload committed under profiles/ (profiles/traffic_r2.json names the .ncu-rep it was read from).  A capture describes the
    unsharded one-GPU launch, so it is only quoted for N = 1; None when no capture of this workload exists."""
    if world != 1:
        return None
    p = os.path.join(ROOT, "profiles", "traffic_r2.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


def effective_lengths(cfg_mod, wl):
    from geobo_b200 import kernels
    gl = cfg_mod.gp_lengthscale * np.asarray([cfg_mod.xvoxsize] * 3) * np.asarray(wl.get("gl_mult", (1.0, 1.0, 1.0)))
    return kernels.dedup_lengthscales(np.array(gl, dtype=float))


def algorithmic_flops(N, Ns, nd, ncol):
    """block-structured dense algorithm, counted once (per rank for its ncol voxel columns).  nrp = property
    blocks actually computed: 3 with drill data; 2 without (the reference's drill cubes are NaN then, inversion.py:213-214, so the
    library does not compute that block -- only the work that is really done is credited)."""
    M = 2 * Ns + nd
    nrp = 3 if nd else 2
    return dict(project=4.0 * nrp * ncol * N * Ns, aka=2.0 * ncol * (3 * Ns * Ns), chol=M ** 3 / 3.0, trsm=float(nrp) * ncol * M * M,
                mean_var=4.0 * nrp * ncol * M, nrp=nrp)


def run_ours(args):
    from geobo_b200 import _lib, config_loader, dist, inversion, synth
    wl = WORKLOADS[args.workload]
    ctx = _lib.default_context()
    rank, world = dist.init_from_env(ctx)
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    xN, yN, zN = wl["shape"]
    cfg = synth.settings(xN, yN, zN, kernelfunc=wl["kernel"], precision=args.precision, refine=args.refine, structure=args.structure)
    config_loader.load_settings(cfg, make_outpath=False)
    kron = args.structure in ("kron", "compact", "fft")        # the two opt-in structure-exploiting projections share the reporting below
    slices = inversion.Inversion._slices()
    N, Ns, nd = xN * yN * zN, xN * yN, wl["nd"]
    fixture = load_fixture(args.workload)
    if fixture is not None:
        # the synthetic inputs of this workload as stored with the full CPU-oracle run (same recipe as synth.make_inputs: cylinders
        # truth cube, forward-simulated surveys rounded through float32, rng(0) drill picks) -- so the cubes can be checked in-run
        f = fixture_inputs(fixture, wl["shape"])
        f["sensor_locations"] = synth.sensor_grid()
    else:
        f = synth.make_inputs(nd=nd, seed=0, ctx=ctx)
        ctx.release_cache()          # the small forward-model problem's buffers: nothing of their size is needed again
    info = ctx.device_info()

    # ---------------- device-resident problem (value) ----------------
    inv = inversion.Inversion()
    inv.create_cubegeometry()
    inv.gp_length = inv.gp_length * np.asarray(wl.get("gl_mult", (1.0, 1.0, 1.0)))
    gl_eff = effective_lengths(config_loader, wl)
    c0, c1 = inv.shard_bounds()[rank] if world > 1 else (0, N)      # balanced by estimated work per voxel column (culling)
    drill_idx = np.flatnonzero(f["drilldata0"].ravel() != 0)
    prob = _lib.Problem(ctx, (xN, yN, zN), (config_loader.xvoxsize, config_loader.yvoxsize, config_loader.zvoxsize),
                        inv.Edges, f["sensor_locations"], config_loader.magneticField, config_loader.c_MILLIGALS_UNITS,
                        config_loader.fcor_grav, 1.0, config_loader.fcor_mag, drill_idx, c0, c1)
    y = np.hstack([(f["grav"] - f["grav"].mean()) / f["grav"].std(), (f["mag"] - f["mag"].mean()) / f["mag"].std(),
                   (f["drillfield"] - f["drillfield"].mean()) / f["drillfield"].std() if nd else np.zeros(0)])
    prob.set_data(y)
    h = prob.hyper(gl_eff, config_loader.gp_err, config_loader.gp_coeff, 1.0, wl["kernel"], slices=slices, refine=args.refine, structure=args.structure)
    t_sens_ms = prob.timings()["a_sens"]
    ClockSampler.init()
    for _ in range(args.warmup):
        prob.predict(h, want_host=False)
    sampler = ClockSampler(ctx_device())
    dist.barrier()
    sampler.start()
    stage_ms = {}
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, logl, info_pd = prob.predict(h, want_host=False)     # synchronises the stream on return
        tm = prob.timings()
        dev_ms += tm["total"]
        for k, v in tm.items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
    wall = time.perf_counter() - t0
    dist.barrier()
    clocks = sampler.stop()
    launches = int(round(tm["launches"]))
    dev_s = dist.max_over_ranks(dev_ms / 1e3)
    wall_s = dist.max_over_ranks(wall)
    ms_per_step = dev_s * 1e3 / args.steps
    value = N * args.steps / dev_s
    stage_ms = {k: v / args.steps for k, v in stage_ms.items() if k != "launches"}
    ks_frac = stage_ms.pop("ksteps_frac", 1.0) or 1.0       # int8 projection: share of the K steps visited (the rest multiply zero digits)

    # ---------------- the fp64 DMMA path on the same device problem (one step; reported as `extra`, not the headline) ----------------
    fp64_extra = None
    if world == 1 and slices and not kron and not args.no_fp64_extra:
        h64 = prob.hyper(gl_eff, config_loader.gp_err, config_loader.gp_coeff, 1.0, wl["kernel"], slices=0, refine=0, structure="dense")
        prob.predict(h64, want_host=False)
        t64 = prob.timings()
        fl64 = algorithmic_flops(N, Ns, nd, c1 - c0)
        pk = read_peaks().get("cublas_dgemm_8192_tflops")
        ach64 = fl64["project"] / (t64["project"] / 1e3) / 1e12
        fp64_extra = {"precision": "fp64 (DMMA tensor pipe, gemm_f64_kernel<B_GEN>)", "value": N / (t64["total"] / 1e3), "unit": "voxels/s",
                      "ms_per_step": t64["total"], "steps": 1, "project_ms": t64["project"],
                      "roofline": {"bound": "tensor", "achieved": ach64, "peak": pk, "unit": "TFLOP/s", "frac": (ach64 / pk) if pk else None}}
    device_bytes = prob.device_bytes()
    prob.close()      # its buffers go to the context's cache and are reused by the end-to-end problem below (one cube resident at a time)

    # ---------------- end to end through Inversion.cubing (host arrays in, six host cubes out) ----------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    dist.barrier()
    inv2 = inversion.Inversion()
    inv2.create_cubegeometry()
    inv2.gp_length = inv2.gp_length * np.asarray(wl.get("gl_mult", (1.0, 1.0, 1.0)))
    gl0 = inv2.gp_length.copy()
    t0 = time.perf_counter()
    inv2.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])   # warm-up = the cold call of this geometry
    first_call_s = dist.max_over_ranks(time.perf_counter() - t0)
    dist.barrier()
    # no NVML polling inside this region: it is host-latency sensitive (problem build = many synchronous driver calls) and
    # NVML queries share driver locks with them; the clocks of the device-timed region above are the ones reported
    t0 = time.perf_counter()
    e2e_dev_ms = 0.0
    for _ in range(e2e_steps):
        inv2.gp_length = gl0.copy()
        cubes = inv2.cubing(f["grav"], f["mag"], f["drillfield"], f["sensor_locations"], f["drilldata0"])
        e2e_dev_ms += inv2.timings["total"]
    e2e_s = dist.max_over_ranks(time.perf_counter() - t0)
    dist.barrier()
    M = 2 * Ns + nd
    # a repeated cubing() on the same cube / sensors / drilled voxels keeps the device problem (geometry only: sensitivities, their
    # digit blocks, workspaces) and uploads just the new data vector; the cold call that also builds the problem is reported beside it
    h2d = 8 * M
    h2d_first = 8 * (f["grav"].size + f["mag"].size + f["sensor_locations"].size + inv2.Edges.size + M) + 8 * nd
    d2h = 8 * (6 * (c1 - c0) + 2) + 4
    e2e = {"value": N * e2e_steps / e2e_s, "unit": "voxels/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "ms_per_step": e2e_s * 1e3 / e2e_steps, "device_ms_per_step": e2e_dev_ms / e2e_steps,
           "includes": "Inversion.cubing(host arrays) -> six host cubes: normalisation, geometry hash, H2D of the data vector, predict3 stage, "
                       "D2H of the result shards (+ NCCL all-gather for N > 1); the device problem of this geometry is reused",
           "first_call": {"ms": first_call_s * 1e3, "voxels_per_s": N / first_call_s, "h2d_bytes": int(h2d_first),
                          "includes": "the same plus the device problem build (A_sens x2 on the GPU, digit blocks, every allocation)"}}
    finite = bool(all(np.isfinite(cb).all() for cb in cubes[:2]))
    parity = parity_vs_fixture(fixture, cubes, inv2.logl, N) if fixture is not None else None
    # size-independent checks on the full cubes (every workload, also where no CPU fixture of that size exists): the posterior
    # variance lies in [0, prior variance] voxel by voxel (prior = gp_amp x data std^2), the cubes are finite (NaN drill cubes without
    # drill data, like the reference)
    checks = {}
    stds = [float(np.std(f["grav"])), float(np.std(f["mag"])), float(np.std(f["drillfield"])) if nd else float("nan")]
    for i, nm in enumerate(("density", "magsus", "drill")):
        if i == 2 and not nd:
            checks["drill_cubes_nan"] = bool(np.isnan(cubes[2]).all() and np.isnan(cubes[5]).all())
            continue
        v = np.asarray(cubes[3 + i])
        checks[nm + "_var_in_prior_bounds"] = bool(np.isfinite(v).all() and v.min() >= -1e-9 * stds[i] ** 2 and v.max() <= (1.0 + 1e-9) * stds[i] ** 2)
        checks[nm + "_var_reduction_min"] = float(1.0 - v.max() / stds[i] ** 2)
        checks[nm + "_rec_finite"] = bool(np.isfinite(cubes[i]).all())
    # BASELINE config 5: "BO acquisition sweep" on the result cubes -- the utility of a vertical drill hole at EVERY voxel column
    # (run_geobo.py:175-200 for all (x, y) at once, csrc/acq.cu) on rank 0, timed end to end (upload of two cubes, sweep, download)
    acq = None
    if args.acq_sweep and rank == 0:
        from geobo_b200 import acquisition
        rec = np.nan_to_num(cubes[2] if nd else cubes[0])
        var = np.nan_to_num(cubes[5] if nd else cubes[3])
        acquisition.sweep_vertical(rec, var, kappa=1.0, beta=0.1)  # warm-up
        t0 = time.perf_counter()
        util, props = acquisition.sweep_vertical(rec, var, kappa=1.0, beta=0.1, top=5)
        t_acq = time.perf_counter() - t0
        acq = {"ms": t_acq * 1e3, "columns": int(util.size), "columns_per_s": util.size / t_acq, "best": [[int(a), int(b), float(u)] for a, b, u in props[:3]],
               "cube": "drill" if nd else "density", "what": "futility_vertical for every voxel column (exhaustive sweep, one launch)"}

    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel ----------------
    peaks = read_peaks()
    fl = algorithmic_flops(N, Ns, nd, c1 - c0)
    proj_s = stage_ms["project"] / 1e3
    achieved = fl["project"] * ks_frac / proj_s / 1e12       # work the tensor pipe really did (skipped all-zero K steps are not credited)
    if kron:
        # opt-in structure-exploiting path, reported separately from the dense contraction: the projection
        # is three Toeplitz mode products per block -- 12 Ns ncol (xN + yN + zN) flops instead of 12 Ns ncol N -- whose
        # algorithmic traffic is one read of A and one write of Pt; bound: HBM.
        kron_bytes = 8.0 * (2.0 * Ns * N + 2.0 * Ns * fl["nrp"] * (c1 - c0))
        hbm = peaks.get("hbm_gbs")
        ach = kron_bytes / proj_s / 1e9
        roofline = {"kernel": ("kron_y_kernel + kron_zx_kernel (Pt = A3.K for the separable exp blocks: y mode into an L2-resident scratch, "
                               "z and x modes in shared memory; fp64 FMA)") if args.structure == "kron" else
                              "stencil_kernel (Pt = A3.K for the compact-support blocks: tap sum over the support window; fp64 FMA)"
                              if args.structure == "compact" else
                              "fft_pass_kernel (Pt = A3.K as zero-padded 3-D FFT convolutions with the stationary tables; fp64 radix-2 passes in shared memory)",
                    "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": (ach / hbm) if hbm else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth measured on this pool)",
                    "algorithmic_bytes_per_step": kron_bytes, "ms_per_step": stage_ms["project"],
                    "mode_product_flops": 4.0 * fl["nrp"] * Ns * (c1 - c0) * (xN + yN + zN) if args.structure == "kron" else None,
                    "dense_flops_replaced": fl["project"],
                    "share_of_step": stage_ms["project"] / ms_per_step, "traffic": None}
        dtype = ("f64 structured projection (%s) + " % args.structure + ("s8 digit slices x%d for A.Pt^T and L^-1.Pt + f64 Cholesky / refinement" % slices if slices else "f64 DMMA"))
    elif slices:
        # int8 digit-slice kernel: every algorithmic fp64 multiply-add is S(S+1)/2 exact int8 digit products on the
        # tensor cores; the ceiling is the measured dense int8 tcgen05 rate (tools/peaks_i8.cu) divided by that count.
        nprod = slices * (slices + 1) // 2
        runs = [r for r in peaks.get("runs", []) if r.get("unrolled") and r.get("n", 0) >= 128]
        i8_peak = max([r["int8_tops"] for r in runs] + [0.0]) or None
        peak = (i8_peak / nprod) if i8_peak else None
        roofline = {"kernel": "ozaki_project_kernel<%d> (fused covariance-digit generation + projection Pt = A.K; tcgen05.mma kind::i8, "
                              "%d balanced 8-bit digits per operand, exact int32 accumulation in TMEM)" % (slices, slices),
                    "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s (fp64-equivalent, algorithmic flops counted once)",
                    "frac": (achieved / peak) if peak else None,
                    "int8_tops_achieved": achieved * nprod, "int8_tops_peak_measured": i8_peak, "digit_products_per_mac": nprod,
                    "peak_source": "dense int8 tcgen05 rate measured on this pool with tools/peaks_i8.cu (profiles/int8_peaks_r1.json: 8192 MAC/clk/SM, "
                                   "nominal 4.5 POP/s = 2x the bf16 figure of MEASURED_PEAKS.json, %s TFLOP/s burst), divided by the %d digit "
                                   "products per multiply-add" % (peaks.get("bf16_tflops"), nprod),
                    "algorithmic_flops_per_launch": fl["project"], "ms_per_launch": stage_ms["project"],
                    "k_steps_visited_frac": ks_frac,
                    "dense_equivalent_rate": fl["project"] / proj_s / 1e12,
                    "culling": "K steps whose covariance digits are all zero for the whole tile are skipped (bitwise-neutral); `achieved` "
                               "counts only the %.1f %% of the dense contraction that was executed, `dense_equivalent_rate` the whole of it" % (100 * ks_frac),
                    "share_of_step": stage_ms["project"] / ms_per_step,
                    "traffic": traffic_from_profiles("project_%s_int8x%d" % (args.workload, slices), world)}
        dtype = "s8 digit slices x%d (exact s32 accumulate) for the three dense products + f64 Cholesky / refinement" % slices
    else:
        fp64_peak = peaks.get("cublas_dgemm_8192_tflops")
        roofline = {"kernel": "gemm_f64_kernel<B_GEN> (fused covariance assembly + projection Pt = A.K, fp64 DMMA)",
                    "bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (achieved / fp64_peak) if fp64_peak else None,
                    "peak_source": "fp64 tensor pipe measured on this pool with tools/peaks.cu (cuBLAS DGEMM 8192^3; DMMA issue peak %s); "
                                   "MEASURED_PEAKS.json carries no fp64 figure (its bf16 figure, %s TFLOP/s, does not bound an fp64 kernel)"
                                   % (peaks.get("dmma_tflops_w16"), peaks.get("bf16_tflops")),
                    "algorithmic_flops_per_launch": fl["project"], "ms_per_launch": stage_ms["project"],
                    "share_of_step": stage_ms["project"] / ms_per_step,
                    "traffic": traffic_from_profiles("project_%s_fp64" % args.workload, world)}
        dtype = "f64"
    out = {"metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype,
           "data": "synthetic (cylinders truth cube, forward-simulated grav/mag surveys, seed 0)",
           "config": workload_config(wl, args.workload),
           "impl_config": {"precision": args.precision + (" + %d refinement step(s)" % args.refine if slices else ""),
                           "property_blocks_computed": fl["nrp"],
                           "structure": ("%s: %s (opt-in fast path, SURVEY 8(f) row 3; not the dense contraction the headline is quoted on)"
                                         % (args.structure, {"kron": "separable exp blocks as Toeplitz mode products",
                                                             "compact": "compact-support blocks as a tap sum over the support window",
                                                             "fft": "block-Toeplitz blocks as zero-padded 3-D FFT convolutions"}[args.structure])) if kron else "dense",
                           "parallelism": "voxel-column shards of Pt x%d" % world, "device_bytes": device_bytes,
                           "inputs": "tests/golden/fullsize_%s.npz (stored synthetic inputs)" % args.workload if fixture is not None else "synth.make_inputs(seed 0)"},
           "parity": parity, "checks": checks, "acquisition_sweep": acq,
           "wall_ms_per_step": wall_s * 1e3 / args.steps, "stage_ms": stage_ms, "a_sens_ms": t_sens_ms,
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches * args.steps, "roofline": roofline,
           "logl": logl, "info": info_pd, "finite": finite, "gpu": info["name"]}
    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, cfg, wl, f, y)
    if fp64_extra is not None:
        out["extra"] = {"fp64_path": fp64_extra}
    print(json.dumps(out))


def ctx_device():
    return int(os.environ.get("GEOBO_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def cpu_baseline(args, cfg, wl, f, y, target_seconds=20.0, pool=None, probe_seconds=None, state=None):
    """Lean NumPy/SciPy restatement of the reference (oracle/numpy_oracle.py) on this box's host cores: the projection in one
    worker process per core (the arrangement of the full-size fixture runs), the other stages with BLAS on all threads."""
    from oracle import numpy_oracle as o
    threads = cpu_threads()
    c = o.make_config(cfg)
    N = c.xNcube * c.yNcube * c.zNcube
    didx = o.drill_indices(f["drilldata0"])
    gl = c.gp_lengthscale * c.xvoxsize * np.asarray(wl.get("gl_mult", (1.0, 1.0, 1.0)))
    # the sensitivities only enter the timed stages as dgemm operands (timing is value-independent): the oracle's A_sens for a
    # few sensors, tiled (oracle.TiledSens) -- the bounded sample does not spend minutes in the (untimed) A_sens loop
    res = o.cpu_baseline_sample(cfg, didx, np.nan_to_num(y), gp_length=gl.copy(), target_seconds=target_seconds, workers=threads, pool=pool,
                                probe_seconds=probe_seconds, state=state)
    out = {"value": N / res["seconds_estimated"], "unit": "voxels/s", "cores": int(threads), "kind": "port",
           "host_cpus": os.cpu_count(), "sample": res["sample"], "full_inversion": bool(res["full"]),
           "seconds_estimated_full": res["seconds_estimated"], "seconds_measured": res["seconds_measured"], "stages_s": res["stages"],
           "pair_seconds": res["pair_seconds"], "probe_seconds": res["probe_seconds"],
           "what": "oracle/numpy_oracle.py lean restatement of geobo predict3; every host core busy: %d worker processes for the "
                   "projection (NumPy ufuncs are single-threaded, so the reference's own single process would leave all but one core "
                   "idle for 2/3 of the time), OpenBLAS / LAPACK on %d threads for the rest" % (res["workers"], threads)}
    g = load_fixture(args.workload)
    if g is not None and "cpu" in g and "dgemm_proj" in json.loads(str(g["cpu"])).get("core_seconds_per_stage", {}):
        # (only for fixtures computed by the DENSE lean path -- the 96x96x48 one went through the Kronecker restatement)
        cf = json.loads(str(g["cpu"]))
        stage_s = cf["wall_s"] - cf["a_sens_s"]           # the metric's stage: everything after the sensitivities
        out["calibration_full_run"] = {"seconds_predict3_stage": stage_s, "worker_processes": cf["workers"], "host_cores": cf["host_cores"],
                                       "voxels_per_s": N / stage_s, "core_seconds_per_stage": cf["core_seconds_per_stage"],
                                       "note": "one COMPLETE oracle inversion of this workload in the same arrangement, run once in the build "
                                               "container (tests/golden/make_fullsize_golden.py; stored in tests/golden/fullsize_%s.npz)" % args.workload}
        out["calibration_full_run"]["sample_vs_full_per_core"] = (out["value"] / threads) / (N / stage_s / cf["workers"])
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is pure Python and
    cannot travel to the GPU box) on the host cores, bounded sample per step, same metric / config / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from geobo_b200 import synth
    from oracle import numpy_oracle as o
    wl = WORKLOADS[args.workload]
    xN, yN, zN = wl["shape"]
    cfg = synth.settings(xN, yN, zN, kernelfunc=wl["kernel"])
    c = o.make_config(cfg)
    N, Ns, nd = xN * yN * zN, xN * yN, wl["nd"]
    rng = np.random.default_rng(0)
    fixture = load_fixture(args.workload)
    if fixture is not None:
        f = fixture_inputs(fixture, wl["shape"])
        f["sensor_locations"] = o.sensor_grid(c)
        with np.errstate(all="ignore"):
            y = np.hstack([(f["grav"] - f["grav"].mean()) / f["grav"].std(), (f["mag"] - f["mag"].mean()) / f["mag"].std(),
                           (f["drillfield"] - f["drillfield"].mean()) / f["drillfield"].std() if nd else np.zeros(0)])
    else:
        d0 = np.zeros(N)
        if nd:
            d0[rng.choice(N, nd, replace=False)] = 1.0
        f = dict(sensor_locations=o.sensor_grid(c), drilldata0=d0.reshape(xN, yN, zN))
        y = rng.standard_normal(2 * Ns + nd)
    per_step_budget = max(5.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, last = [], None
    import multiprocessing as mp
    # one worker pool and one probe for all steps (process start-up is not part of a step)
    state = {}
    with mp.get_context("spawn").Pool(cpu_threads()) as pool:
        for i in range(args.warmup + args.steps):
            last = cpu_baseline(args, cfg, wl, f, y, target_seconds=per_step_budget, pool=pool,
                                probe_seconds=None if last is None else last["probe_seconds"], state=state)
            if i >= args.warmup:
                vals.append(last["value"])
    value = float(np.mean(vals))
    cb = {k: last[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cpus", "full_inversion", "pair_seconds") if k in last}
    cb["value"] = value
    cb["spread_over_steps"] = {"min": float(np.min(vals)), "max": float(np.max(vals)), "rel": float((np.max(vals) - np.min(vals)) / value)}
    if "calibration_full_run" in last:
        cb["calibration_full_run"] = last["calibration_full_run"]
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": N / value * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic (cylinders truth cube, forward-simulated grav/mag surveys, seed 0)",
           "config": workload_config(wl, args.workload),
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "ms_per_step is the estimated whole-cube time of one CPU inversion; each step times a bounded sample (see cpu_baseline.sample)"
                   if not last.get("full_inversion") else "every step is one complete CPU inversion"}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GEOBO_B200_WORKLOAD", "cfg3"), choices=sorted(WORKLOADS))
    ap.add_argument("--refine", type=int, default=1, help="refinement steps of the int8 paths (fp64 matrix-free residual)")
    ap.add_argument("--precision", default=os.environ.get("GEOBO_B200_PRECISION", "int8x5"),
                    choices=["fp64", "int8x4", "int8x5", "int8x6"],
                    help="projection arithmetic: fp64 DMMA, or error-free int8 digit products on tcgen05 (31/39/47 bits)")
    ap.add_argument("--structure", default=os.environ.get("GEOBO_B200_STRUCTURE", "dense"), choices=["dense", "kron", "compact", "fft"],
                    help="dense: the contraction Pt = A3.K the metric is quoted on (default); kron: separable exp blocks as mode products "
                         "(exp workloads only); compact: tap sum over the support of the compact kernel (sparse workloads only); fft: 3-D FFT "
                         "convolutions (any kernel); all reported separately")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-warmup", type=int, default=3, help="lower bound of --warmup (3 by the measurement contract; 1 only for the "
                    "multi-GPU lines of workloads whose steps take tens of seconds, and said so in the line's `warmup`)")
    ap.add_argument("--acq-sweep", action="store_true", help="also time the exhaustive vertical-drillhole acquisition sweep on the result cubes (BASELINE config 5)")
    ap.add_argument("--no-fp64-extra", action="store_true", help="skip the one extra step on the fp64 DMMA path (reported under `extra`)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, args.min_warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
