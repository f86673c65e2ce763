"""ctypes binding of ``libgeobo_b200.so`` (the C ABI declared in ``include/geobo_b200.h``).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
present, every compute entry point raises ``GeoboB200Error``.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgeobo_b200.so")

GB_OK = 0
KERNEL_IDS = {"sparse": 0, "exp": 1, "matern32": 2}
STRUCTURE_IDS = {"dense": 0, "kron": 1, "compact": 2, "fft": 3}     # gb_hyper.structure (GB_STRUCTURE_*)
SENS_KINDS = {"grav": 0, "magn": 1}
FLAG_MEAN, FLAG_VAR, FLAG_LOGL, FLAG_ALL = 1, 2, 4, 7
TIMER_NAMES = ["a_sens", "tables", "project", "drill_rows", "aka", "allreduce", "chol", "trsm", "mean_var", "total", "d2h", "launches",
               "ksteps_frac"]


class GeoboB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("geobo_b200 error %d: %s" % (code, msg))
        self.code = code


class ProblemDesc(C.Structure):
    _fields_ = [("ncube", C.c_int64 * 3), ("voxsize", C.c_double * 3), ("edges", C.c_void_p), ("locations", C.c_void_p),
                ("nsens", C.c_int64), ("magnetic_field", C.c_double * 3), ("grav_mul", C.c_double), ("grav_div", C.c_double),
                ("magn_mul", C.c_double), ("magn_div", C.c_double), ("drill_idx", C.c_void_p), ("ndrill", C.c_int64),
                ("col_begin", C.c_int64), ("col_end", C.c_int64)]


class Hyper(C.Structure):
    _fields_ = [("gp_length", C.c_double * 3), ("gp_sigma", C.c_double * 3), ("coeffm", C.c_double * 3),
                ("gp_amp", C.c_double), ("kernel_id", C.c_int), ("slices", C.c_int), ("refine", C.c_int), ("structure", C.c_int)]


# name -> (restype, argtypes); every symbol here is declared in include/geobo_b200.h
_P = C.c_void_p
_SIGNATURES = {
    "gb_version": (C.c_int, []),
    "gb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "gb_ctx_destroy": (C.c_int, [_P]),
    "gb_ctx_release_cache": (C.c_int, [_P]),
    "gb_last_error": (C.c_char_p, [_P]),
    "gb_device_info": (C.c_int, [_P, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gb_comm_unique_id": (C.c_int, [_P, _P]),
    "gb_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "gb_comm_allgather": (C.c_int, [_P, _P, C.c_int64, _P]),
    "gb_grid_points": (C.c_int, [_P, _P, _P, _P]),
    "gb_sqdist": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P]),
    "gb_create_cov": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int, _P]),
    "gb_cov_function": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int64, C.c_double, C.c_double, _P]),
    "gb_create_cov_grid": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, C.c_int, _P, C.POINTER(C.c_float)]),
    "gb_a_sens": (C.c_int, [_P, C.c_int, _P, _P, C.c_int64, _P, _P, C.c_double, C.c_double, _P]),
    "gb_corner_func": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int64, _P, _P]),
    "gb_a_drill": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, _P]),
    "gb_problem_create": (C.c_int, [_P, C.POINTER(ProblemDesc), C.POINTER(_P)]),
    "gb_problem_destroy": (C.c_int, [_P]),
    "gb_problem_set_data": (C.c_int, [_P, _P]),
    "gb_predict": (C.c_int, [_P, C.POINTER(Hyper), C.c_int, _P, _P, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "gb_neg_logl": (C.c_int, [_P, C.POINTER(Hyper), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "gb_problem_get_sens": (C.c_int, [_P, C.c_int, _P]),
    "gb_forward": (C.c_int, [_P, C.c_int, _P, _P]),
    "gb_acquisition_vertical": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, C.c_double, _P]),
    "gb_acquisition_drill": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_double, C.c_double, _P, C.c_int64, C.c_double, C.c_double, _P]),
    "gb_align_drill": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int64, _P, _P]),
    "gb_posterior_cov": (C.c_int, [_P, C.POINTER(Hyper), _P]),
    "gb_get_timings": (C.c_int, [_P, _P, C.c_int]),
    "gb_problem_device_bytes": (C.c_uint64, [_P]),
}

_lib = None
_lock = threading.Lock()


def load_library():
    """dlopen the in-tree shared library and attach the signatures.  Raises if it was not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise GeoboB200Error(-2, "%s not found: build it with `python geobo_b200/csrc/build.py` "
                                         "(there is no CPU fallback)" % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One CUDA context/stream (and optionally one NCCL rank) -- ``gb_ctx``."""

    def __init__(self, device=-1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.gb_ctx_create(int(device), C.byref(h))
        if rc != GB_OK:
            raise GeoboB200Error(rc, self.lib.gb_last_error(None).decode())
        self.h = h
        self.rank, self.nranks = 0, 1

    def check(self, rc):
        if rc < 0:
            raise GeoboB200Error(rc, self.lib.gb_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.gb_ctx_destroy(self.h)
            self.h = None

    def release_cache(self):
        """Return the cached device / pinned-host buffers of closed problems to the driver."""
        self.check(self.lib.gb_ctx_release_cache(self.h))

    def device_info(self):
        name = C.create_string_buffer(256)
        sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
        fr, tot = C.c_uint64(), C.c_uint64()
        self.check(self.lib.gb_device_info(self.h, name, 256, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(fr), C.byref(tot)))
        return dict(name=name.value.decode(), sm_count=sm.value, cc=(maj.value, mnr.value), free_bytes=fr.value,
                    total_bytes=tot.value)

    # ---- multi-GPU
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self.check(self.lib.gb_comm_unique_id(self.h, buf))
        return bytes(buf.raw)

    def comm_init(self, uid, rank, nranks):
        buf = C.create_string_buffer(bytes(uid), 128)
        self.check(self.lib.gb_comm_init(self.h, buf, int(rank), int(nranks)))
        self.rank, self.nranks = int(rank), int(nranks)

    # ---- geobo/run_geobo.py acquisition functions
    def acquisition_vertical(self, rec, var, kappa, beta, costs=None):
        rec, var = _f64(rec), _f64(var)
        shp = np.ascontiguousarray(rec.shape, dtype=np.int64)
        cst = _f64(costs) if costs is not None else None
        out = np.empty(rec.shape[:2])
        self.check(self.lib.gb_acquisition_vertical(self.h, _ptr(rec), _ptr(var), _ptr(cst) if cst is not None else None, _ptr(shp),
                                                    float(kappa), float(beta), _ptr(out)))
        return out

    def acquisition_drill(self, rec, var, voxsize, zmax, length, params, kappa, beta, costs=None):
        rec, var = _f64(rec), _f64(var)
        shp = np.ascontiguousarray(rec.shape, dtype=np.int64)
        cst = _f64(costs) if costs is not None else None
        prm = _f64(params).reshape(-1, 4)
        vs = _f64(voxsize)
        out = np.empty(prm.shape[0])
        self.check(self.lib.gb_acquisition_drill(self.h, _ptr(rec), _ptr(var), _ptr(cst) if cst is not None else None, _ptr(shp), _ptr(vs),
                                                 float(zmax), float(length), _ptr(prm), int(prm.shape[0]), float(kappa), float(beta), _ptr(out)))
        return out

    # ---- geobo/run_geobo.py drill data
    def align_drill(self, voxelpos, coord, data, voxsize):
        """Flat voxel cube (voxel order of ``voxelpos``, shape (3, N)) of the windowed sample means (run_geobo.py:132-159)."""
        vp = _f64(voxelpos).reshape(3, -1)
        cd = _f64(coord).reshape(-1, 3)
        dv = _f64(data).ravel()
        if dv.size != cd.shape[0]:
            raise ValueError("align_drill: %d coordinates but %d data values" % (cd.shape[0], dv.size))
        vs = _f64(voxsize)
        out = np.empty(vp.shape[1])
        ns = int(cd.shape[0])
        self.check(self.lib.gb_align_drill(self.h, _ptr(vp), int(vp.shape[1]), _ptr(cd) if ns else None, _ptr(dv) if ns else None, ns,
                                           _ptr(vs), _ptr(out)))
        return out

    def allgather(self, local):
        """NCCL all-gather of equally sized float64 host arrays: returns (nranks, local.size)."""
        loc = _f64(local).ravel()
        out = np.empty((max(1, getattr(self, "nranks", 1)), loc.size))
        self.check(self.lib.gb_comm_allgather(self.h, _ptr(loc), int(loc.size), _ptr(out)))
        return out

    # ---- geobo/kernels.py
    def grid_points(self, lpix, pixscale):
        lp = np.ascontiguousarray(lpix, dtype=np.int64)
        sc = _f64(pixscale)
        out = np.empty((int(lp.prod()), 3))
        self.check(self.lib.gb_grid_points(self.h, _ptr(lp), _ptr(sc), _ptr(out)))
        return out

    def sqdist(self, points):
        p = _f64(points)
        n, dim = p.shape
        out = np.empty((n, n))
        self.check(self.lib.gb_sqdist(self.h, _ptr(p), n, dim, _ptr(out)))
        return out

    def create_cov(self, D2, gpl, w, kernel):
        d2 = _f64(D2)
        n = d2.shape[0]
        g, ww = _f64(gpl), _f64(w)
        out = np.empty((3 * n, 3 * n))
        self.check(self.lib.gb_create_cov(self.h, _ptr(d2), n, _ptr(g), _ptr(ww), KERNEL_IDS[kernel], _ptr(out)))
        return out

    def cov_function(self, kernel, cross, D2, l1, l2=0.0):
        d2 = _f64(D2)
        out = np.empty_like(d2)
        self.check(self.lib.gb_cov_function(self.h, KERNEL_IDS[kernel], int(cross), _ptr(d2), d2.size, float(l1), float(l2), _ptr(out)))
        return out

    def create_cov_grid(self, ncube, voxsize, gpl, w, amp, kernel, want_output=True):
        nc = np.ascontiguousarray(ncube, dtype=np.int64)
        vs, g, ww = _f64(voxsize), _f64(gpl), _f64(w)
        n = int(nc.prod())
        out = np.empty((3 * n, 3 * n)) if want_output else None
        ms = C.c_float()
        self.check(self.lib.gb_create_cov_grid(self.h, _ptr(nc), _ptr(vs), _ptr(g), _ptr(ww), float(amp), KERNEL_IDS[kernel],
                                               _ptr(out) if out is not None else None, C.byref(ms)))
        return out, ms.value

    # ---- geobo/sensormodel.py
    def a_sens(self, kind, B, locations, edges, ncube, mul, div):
        b, loc, ed = _f64(B), _f64(locations), _f64(edges)
        nc = np.ascontiguousarray(ncube, dtype=np.int64)
        out = np.empty((loc.shape[0], int(nc.prod())))
        self.check(self.lib.gb_a_sens(self.h, SENS_KINDS[kind], _ptr(b), _ptr(loc), loc.shape[0], _ptr(ed), _ptr(nc),
                                      float(mul), float(div), _ptr(out)))
        return out

    def corner_func(self, kind, x, y, z, B=None):
        x, y, z = np.broadcast_arrays(_f64(x), _f64(y), _f64(z))
        x, y, z = _f64(x), _f64(y), _f64(z)
        out = np.empty_like(x)
        b = _f64(B if B is not None else [0.0, 0.0, 0.0])
        self.check(self.lib.gb_corner_func(self.h, SENS_KINDS[kind], _ptr(x), _ptr(y), _ptr(z), x.size, _ptr(b), _ptr(out)))
        return out

    def a_drill(self, loc, voxelpos):
        loc, vp = _f64(loc), _f64(voxelpos)
        nd, n = loc.shape[1], vp.shape[1]
        out = np.zeros((nd, n))
        self.check(self.lib.gb_a_drill(self.h, _ptr(loc), nd, _ptr(vp), n, _ptr(out)))
        return out


class Problem:
    """Device-resident joint-inversion problem -- ``gb_problem``."""

    def __init__(self, ctx, ncube, voxsize, edges, locations, magnetic_field, grav_mul, grav_div, magn_mul, magn_div,
                 drill_idx, col_begin=0, col_end=0):
        self.ctx = ctx
        self.lib = ctx.lib
        self._edges, self._loc = _f64(edges), _f64(locations)
        self._drill = np.ascontiguousarray(drill_idx, dtype=np.int64)
        d = ProblemDesc()
        d.ncube[:] = [int(v) for v in ncube]
        d.voxsize[:] = [float(v) for v in voxsize]
        d.edges, d.locations = self._edges.ctypes.data, self._loc.ctypes.data
        d.nsens = self._loc.shape[0]
        d.magnetic_field[:] = [float(v) for v in magnetic_field]
        d.grav_mul, d.grav_div, d.magn_mul, d.magn_div = float(grav_mul), float(grav_div), float(magn_mul), float(magn_div)
        d.drill_idx = self._drill.ctypes.data if self._drill.size else None
        d.ndrill = self._drill.size
        d.col_begin, d.col_end = int(col_begin), int(col_end)
        self.N = int(np.prod([int(v) for v in ncube]))
        self.Ns, self.nd = int(d.nsens), int(d.ndrill)
        self.M = 2 * self.Ns + self.nd
        self.col_begin = int(col_begin)
        self.col_end = int(col_end) if col_end else self.N
        self.ncol = self.col_end - self.col_begin
        h = C.c_void_p()
        ctx.check(self.lib.gb_problem_create(ctx.h, C.byref(d), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.gb_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_data(self, fs3):
        y = _f64(fs3)
        if y.size != self.M:
            raise ValueError("data vector has %d entries, expected M=%d" % (y.size, self.M))
        self.ctx.check(self.lib.gb_problem_set_data(self.h, _ptr(y)))

    @staticmethod
    def hyper(gp_length, gp_sigma, coeffm, gp_amp, kernel, slices=0, refine=1, structure="dense"):
        h = Hyper()
        h.gp_length[:] = [float(v) for v in gp_length]
        h.gp_sigma[:] = [float(v) for v in gp_sigma]
        h.coeffm[:] = [float(v) for v in coeffm]
        h.gp_amp = float(gp_amp)
        if kernel not in KERNEL_IDS:
            raise ValueError("kernelfunc must be one of %s, got %r" % (sorted(KERNEL_IDS), kernel))
        h.kernel_id = KERNEL_IDS[kernel]
        h.slices = int(slices)
        h.refine = int(refine)
        if structure not in STRUCTURE_IDS:
            raise ValueError("structure must be one of %s, got %r" % (sorted(STRUCTURE_IDS), structure))
        h.structure = STRUCTURE_IDS[structure]
        return h

    def predict(self, hyper, want_host=True, flags=FLAG_ALL):
        """Returns (mu[3, ncol], var[3, ncol], logl, info); mu/var None when ``want_host`` is False."""
        mu = np.empty((3, self.ncol)) if want_host else None
        var = np.empty((3, self.ncol)) if want_host else None
        logl, info = C.c_double(), C.c_int()
        rc = self.lib.gb_predict(self.h, C.byref(hyper), int(flags), _ptr(mu) if want_host else None,
                                 _ptr(var) if want_host else None, C.byref(logl), C.byref(info))
        self.ctx.check(rc)
        return mu, var, logl.value, info.value

    def neg_logl(self, hyper):
        out, info = C.c_double(), C.c_int()
        self.ctx.check(self.lib.gb_neg_logl(self.h, C.byref(hyper), C.byref(out), C.byref(info)))
        return out.value, info.value

    def sens(self, kind):
        out = np.empty((self.Ns, self.N))
        self.ctx.check(self.lib.gb_problem_get_sens(self.h, SENS_KINDS[kind], _ptr(out)))
        return out

    def forward(self, kind, x):
        x = _f64(x).ravel()
        out = np.empty(self.Ns)
        self.ctx.check(self.lib.gb_forward(self.h, SENS_KINDS[kind], _ptr(x), _ptr(out)))
        return out

    def posterior_cov(self, hyper):
        out = np.empty((3 * self.N, 3 * self.N))
        self.ctx.check(self.lib.gb_posterior_cov(self.h, C.byref(hyper), _ptr(out)))
        return out

    def timings(self):
        ms = np.zeros(16)
        self.ctx.check(self.lib.gb_get_timings(self.h, _ptr(ms), 16))
        return {n: float(ms[i]) for i, n in enumerate(TIMER_NAMES)}

    def device_bytes(self):
        return int(self.lib.gb_problem_device_bytes(self.h))


_default_ctx = None


def default_context():
    """Process-wide context on the current device (LOCAL_RANK if set) -- created on first use."""
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("GEOBO_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _default_ctx = Context(dev)
    return _default_ctx
