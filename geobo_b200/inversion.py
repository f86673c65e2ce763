"""``geobo.inversion`` surface on the B200 (reference: ``geobo/inversion.py``).

``Inversion().cubing(gravfield, magfield, drillfield, sensor_locations, drilldata0)`` returns the
same six ``(yN, xN, zN)`` cubes as the reference.  The numerical path -- sensitivity matrices,
multi-output covariance, projection ``A K A^T``, Cholesky, triangular solves, posterior mean and
variance diagonal -- runs in ``libgeobo_b200.so``; nothing of size N x N or 3N x 3N is ever formed
unless a caller explicitly asks for one of the reference's dense attributes (``D2``, ``kcov``,
``Asens3``, ``cov_rec``), which are materialised lazily on the device for small cubes only.

Behaviour kept from the reference: Q1 in-place de-duplication of
``gp_length``; Q5 ``N_vox log 2 pi`` in ``logl``; Q6 ``xvoxsize`` for every length scale; Q7
``xNcube*yNcube`` sensors; Q8 voxel order / output shape; Q9 no drill data -> NaN drill cubes;
a non-positive-definite ``AkA`` prints the reference's two lines and ``sys.exit(1)``.
"""
import sys

import numpy as np

from . import _lib
from . import config_loader as _cfg
from . import dist as _dist
from . import kernels as kernel


class PosteriorCovariance:
    """Stand-in for the 3N x 3N posterior covariance ``kcov - V^T V`` (``inversion.py:117``).

    Only its diagonal is used by the reference (``np.diag(self.cov_rec)``, ``inversion.py:238``); it is
    held as a vector.  ``np.asarray(obj)`` materialises the dense matrix on the device (small cubes)."""

    def __init__(self, diag, owner):
        self._diag = diag
        self._owner = owner
        self.shape = (diag.size, diag.size)

    def diagonal(self):
        return self._diag

    def __array__(self, dtype=None, copy=None):
        dense = self._owner._dense_posterior_cov()
        return dense if dtype is None else dense.astype(dtype)


class Inversion:
    """Joint GP inversion of gravity, magnetic and drill-core data into 3-D property cubes."""

    def __init__(self):
        _cfg.require("gp_lengthscale", "gp_err", "gp_coeff", "xNcube")
        # inversion.py:46-51
        self.gp_length = _cfg.gp_lengthscale * np.asarray([_cfg.xvoxsize, _cfg.xvoxsize, _cfg.xvoxsize])
        self.gp_sigma = np.asarray(_cfg.gp_err)
        self.coeffm = np.asarray(_cfg.gp_coeff)
        self.gp_amp = 1.
        self._problem = None
        self._lazy = {}

    # ------------------------------------------------------------------ geometry (inversion.py:54-74)
    def create_cubegeometry(self):
        """Voxel edge lattice ``Edges`` (3, yN+1, xN+1, zN+1; z stored depth-positive) and voxel centres."""
        xN, yN, zN = _cfg.xNcube, _cfg.yNcube, _cfg.zNcube
        xedge = np.linspace(0, xN, xN + 1) * _cfg.xvoxsize
        yedge = np.linspace(0, yN, yN + 1) * _cfg.yvoxsize
        zedge = np.linspace(0, -zN, zN + 1) * _cfg.zvoxsize + _cfg.zmax
        xE, yE, zE = np.meshgrid(xedge, yedge, zedge)
        self.Edges = np.asarray([xE, yE, -zE])
        xc = np.arange(_cfg.xvoxsize / 2., _cfg.xLcube + _cfg.xvoxsize / 2., _cfg.xvoxsize)
        yc = np.arange(_cfg.yvoxsize / 2., _cfg.yLcube + _cfg.yvoxsize / 2., _cfg.yvoxsize)
        zc = _cfg.zmax - np.arange(_cfg.zvoxsize / 2., _cfg.zLcube + _cfg.zvoxsize / 2., _cfg.zvoxsize)
        self.xxx, self.yyy, self.zzz = np.meshgrid(xc, yc, zc)
        self.voxelpos = np.vstack([self.xxx.flatten(), self.yyy.flatten(), self.zzz.flatten()])
        return self.voxelpos

    # ------------------------------------------------------------------ device problem
    @staticmethod
    def _slices():
        """Optional settings key ``precision`` (absent in reference YAMLs -> 'auto'):
        'auto'   : 'int8x5' whenever the cube admits the tensor-core path (zNcube % 16 == 0, e.g. both committed examples),
                   else 'fp64' -- so an unmodified reference YAML runs on tcgen05 where it can.  int8x5 is the floor: 39-bit
                   operands keep the (unrefined) variance at <= 3e-7 of its maximum on the full-size fixtures, 31 bits (int8x4)
                   only reach 8.5e-6, a hair inside the 1e-5 target (profiles/r1_int8_vs_fp64_32cube.log);
        'fp64'   : projection A.K on the fp64 tensor pipe (DMMA), any cube shape;
        'int8x4' / 'int8x5' / 'int8x6' : exact int8 digit products on tcgen05/TMEM with 31 / 39 / 47 bits per
        operand for the three dense products (A.K, A.Pt^T, L^-1.Pt); the Cholesky stays fp64 and the mean comes from
        ``refine`` (optional key, default 1) steps of iterative refinement against the fp64 matrix-free operator.
        Needs zNcube % 16 == 0 (refused loudly otherwise).
        The environment variable GEOBO_B200_DEFAULT_PRECISION replaces the default for settings without the key."""
        import os
        prec = str(getattr(_cfg, "precision", os.environ.get("GEOBO_B200_DEFAULT_PRECISION", "auto"))).lower()
        if prec == "auto":
            return 5 if int(_cfg.zNcube) % 16 == 0 else 0
        if prec in ("fp64", "f64", "double"):
            return 0
        if prec in ("int8x4", "int8x5", "int8x6"):
            return int(prec[-1])
        raise ValueError("settings key 'precision' must be 'auto', 'fp64', 'int8x4', 'int8x5' or 'int8x6', got %r" % prec)

    @property
    def precision_used(self):
        """'fp64' or 'int8xS': the arithmetic the settings resolve to for this cube."""
        s = self._slices()
        return "int8x%d" % s if s else "fp64"

    @staticmethod
    def _structure():
        """Optional settings key ``structure`` (absent in reference YAMLs -> 'dense'):
        'dense' : the dense contraction ``Asens3 . kcov`` (``inversion.py:96,114``), any kernel;
        'kron'  : for ``kernelfunc: 'exp'`` only -- on the voxel grid every block of ``create_cov`` (``kernels.py:81-99``)
        is the Kronecker product of three small Toeplitz matrices, so the products with ``kcov`` run as three mode
        products per block.  Same result to rounding; other kernels are refused by the library;
        'compact' : for ``kernelfunc: 'sparse'`` only -- the compact-support kernels (``kernels.py:101-138``) vanish beyond
        their length scale, so the products with ``kcov`` run as a tap sum over the offsets inside the support;
        'fft'   : any kernel -- every block of ``create_cov`` is block-Toeplitz on the voxel grid, so the products with
        ``kcov`` run as zero-padded 3-D FFT convolutions with the block's values on the offset lattice (fp64);
        'auto'  : 'kron' for 'exp', 'compact' for 'sparse', 'fft' otherwise."""
        st = str(getattr(_cfg, "structure", "dense")).lower()
        if st == "auto":        # the cheapest structured form the kernel family admits
            return {"exp": "kron", "sparse": "compact"}.get(_cfg.kernelfunc, "fft")
        if st not in _lib.STRUCTURE_IDS:
            raise ValueError("settings key 'structure' must be 'dense', 'kron', 'compact', 'fft' or 'auto', got %r" % st)
        return st

    def _hyper(self, gp_length=None, coeffm=None, gp_amp=None):
        return _lib.Problem.hyper(self.gp_length if gp_length is None else gp_length, self.gp_sigma,
                                  self.coeffm if coeffm is None else coeffm,
                                  self.gp_amp if gp_amp is None else gp_amp, _cfg.kernelfunc, self._slices(),
                                  int(getattr(_cfg, "refine", 1)), self._structure())

    def _culling_reach_rows(self, slices):
        """Voxel rows beyond which every digit of the covariance tables is zero for the int8 path with ``slices`` digits (an
        estimate for load balancing only -- the kernel takes the exact extents from the digit tables): the blocks' values along
        the y axis, evaluated on the device, against the last-digit threshold 2^-(8 slices) of the block's maximum."""
        gl = np.array(self.gp_length, dtype=float).copy()
        kernel.dedup_lengthscales(gl)
        yN = int(_cfg.yNcube)
        d2 = (np.arange(yN, dtype=float) * _cfg.yvoxsize) ** 2
        ctx = _lib.default_context()
        reach = 0
        with np.errstate(all="ignore"):
            for c in range(2):
                for r in range(3):
                    v = np.abs(ctx.cov_function(_cfg.kernelfunc, int(c != r), d2, gl[c], gl[r] if c != r else 0.0))
                    v = np.nan_to_num(v, nan=0.0, posinf=0.0)
                    if v.max() > 0:
                        nz = np.flatnonzero(v >= v.max() * 2.0 ** (-8 * slices))
                        reach = max(reach, int(nz.max()) if nz.size else 0)
        return reach

    def shard_bounds(self):
        """Voxel-column shards of all ranks.  On the tensor-core path they balance the estimated work per voxel column -- the
        projection visits only the voxel rows inside the covariance's reach (zero-digit culling), fewer for the cube's edge rows;
        the triangular-solve / variance stages cost the same for every column -- instead of the column count."""
        xN, yN, zN = _cfg.xNcube, _cfg.yNcube, _cfg.zNcube
        N, world = xN * yN * zN, _dist.world_size()
        slices = self._slices()
        if world <= 1 or not slices or self._structure() != "dense":
            return _dist.shard_bounds(N, world)
        # (memoised: the estimate costs a few small device calls with allocations, ~0.1 s next to tens of GB of live buffers)
        nd_now = int(np.count_nonzero(np.asarray(self.drilldata0))) if hasattr(self, "drilldata0") else 0
        key = (xN, yN, zN, world, slices, _cfg.kernelfunc, tuple(np.asarray(self.gp_length, dtype=float).tolist()), float(_cfg.yvoxsize), nd_now)
        cache = self.__dict__.setdefault("_bounds_cache", {})
        if key in cache:
            return cache[key]
        cache[key] = self._shard_bounds_weighted(xN, yN, zN, N, world, slices)
        return cache[key]

    def _shard_bounds_weighted(self, xN, yN, zN, N, world, slices):
        reach = self._culling_reach_rows(slices)
        iy = np.arange(yN)
        nrows = np.minimum(yN - 1, iy + reach) - np.maximum(0, iy - reach) + 1
        share = float(getattr(_cfg, "shard_projection_share", 0.75))     # share of the projection in the per-column cost
        # the ranks that own edge rows get more columns, i.e. a larger slice of Pt: keep the largest shard inside the device memory
        # (rough per-column / fixed byte counts of the tensor-core path), falling back towards the uniform split
        Ns = xN * yN
        nd = int(np.count_nonzero(np.asarray(self.drilldata0))) if hasattr(self, "drilldata0") else 0
        Mp = -(-(2 * Ns + nd) // 128) * 128
        per_col = Mp * (3 if nd else 2) * 8.0 + 3.0 * slices * Ns         # Pt, shard digit blocks, row digits of one AkA block
        fixed = 2.0 * Mp * Mp * 8.0 + 20e9                                # AkA / L and Linv; digit scratch, chunk buffers, tables, NCCL
        # The limit must be the SAME number on every rank (all ranks have to pick the same split), so it comes from the device's
        # total memory, not from what happens to be free on this rank; 0.8 of it leaves room for context, NCCL and fragmentation
        # (at 128x128x64 on 8 GPUs an early, laxer rule left 0.24 GB free on the edge ranks, profiles/r2_memlog_cfg5_dense_n8_rank0.txt).
        total = float(_lib.default_context().device_info()["total_bytes"])
        bounds = _dist.shard_bounds(N, world)
        for sh in (share, 0.5 * share, 0.25 * share):
            w = sh * nrows / nrows.max() + (1.0 - sh)
            cand = _dist.shard_bounds(N, world, weights=w)
            if max(c1 - c0 for c0, c1 in cand) * per_col + fixed < 0.8 * total:
                bounds = cand
                break
        return bounds

    def _build_problem(self):
        if not hasattr(self, "Edges"):
            self.create_cubegeometry()
        xN, yN, zN = _cfg.xNcube, _cfg.yNcube, _cfg.zNcube
        N = xN * yN * zN
        # A_drill (sensormodel.py:136-153) selects, for every non-zero entry of drilldata0 in flat order,
        # exactly that voxel (inversion.py:219 builds the coordinates from the same arrays).
        self._drill_idx = np.flatnonzero(np.asarray(self.drilldata0).ravel() != 0)
        nsens = xN * yN
        loc = np.asarray(self.sensor_locations, dtype=float)
        if loc.shape[0] < nsens:
            raise IndexError("sensor_locations has %d rows; the forward model needs xNcube*yNcube=%d" % (loc.shape[0], nsens))
        ctx = _lib.default_context()
        world, rk = _dist.world_size(), _dist.rank()
        self._bounds = self.shard_bounds() if world > 1 else [(0, N)]
        c0, c1 = self._bounds[rk]
        # The device problem (sensitivity matrices, their digit blocks, every workspace) depends only on the geometry: a repeated
        # cubing() on the same cube, sensors and drilled voxels -- new survey data, new hyper-parameters -- keeps it and only
        # uploads the new data vector instead of recomputing A_sens and re-slicing its digits.
        import hashlib
        hsh = hashlib.blake2b(digest_size=16)
        for a in (np.ascontiguousarray(self.Edges, dtype=float), np.ascontiguousarray(loc[:nsens]), np.ascontiguousarray(self._drill_idx, dtype=np.int64),
                  np.ascontiguousarray(_cfg.magneticField, dtype=float)):
            hsh.update(a.tobytes())
        key = (xN, yN, zN, float(_cfg.xvoxsize), float(_cfg.yvoxsize), float(_cfg.zvoxsize), float(_cfg.c_MILLIGALS_UNITS), float(_cfg.fcor_grav),
               float(_cfg.fcor_mag), c0, c1, id(ctx), hsh.hexdigest())
        if self._problem is not None and getattr(self._problem, "h", None) and getattr(self, "_problem_key", None) == key:
            self._problem.set_data(self.Fs3)
            self._lazy = {}
            return
        if self._problem is not None:
            self._problem.close()
        self._problem = _lib.Problem(ctx, (xN, yN, zN), (_cfg.xvoxsize, _cfg.yvoxsize, _cfg.zvoxsize), self.Edges,
                                     loc[:nsens], _cfg.magneticField, _cfg.c_MILLIGALS_UNITS, _cfg.fcor_grav, 1.0,
                                     _cfg.fcor_mag, self._drill_idx, c0, c1)
        self._problem_key = key
        self._problem_builds = getattr(self, "_problem_builds", 0) + 1
        self._problem.set_data(self.Fs3)
        self._lazy = {}

    # ------------------------------------------------------------------ predict3 (inversion.py:77-122)
    def predict3(self, calclogl=False):
        """Posterior mean, covariance (lazy; diagonal held) and log marginal likelihood."""
        if self._problem is None:
            raise RuntimeError("predict3 needs the data set up by cubing()")
        with np.errstate(all="ignore"):
            self.datastd = np.mean([np.nanstd(self.gravfield), np.nanstd(self.magfield), np.nanstd(self.drillfield)])
        # kernels.create_cov de-duplicates the length scales in place on the caller's array (Q1)
        kernel.dedup_lengthscales(np.asarray(self.gp_length))
        mu, var, logl, info = self._problem.predict(self._hyper())
        if info > 0:
            print("Cholesky decompostion failed, AkA matrix i likely not positive semitive.")
            print("Change GP parameter settings")
            sys.exit(1)
        N = self._problem.N
        if _dist.world_size() > 1:
            both = _dist.allgather_columns(np.vstack([mu, var]), N, ctx=self._problem.ctx, bounds=self._bounds)
            mu, var = both[:3], both[3:]
        self._lazy = {}
        mu = np.ascontiguousarray(mu).reshape(3 * N)
        covar = PosteriorCovariance(np.ascontiguousarray(var).reshape(3 * N), self)
        return mu, covar, (logl if calclogl else 0.)

    def calc_logl(self, params):
        """Negative log marginal likelihood for ``[amp, lengthscale factor, w1, w2, w3]`` (``inversion.py:125-152``)."""
        gp_length = params[1] * np.asarray([_cfg.xvoxsize, _cfg.xvoxsize, _cfg.xvoxsize])
        kernel.dedup_lengthscales(gp_length)
        # numerical failures (non-PD AkA, non-finite result) come back as +inf from the library (inversion.py:150-152); anything else
        # -- CUDA / NCCL errors, an unsupported configuration -- is a real error and propagates instead of poisoning the optimiser
        h = self._hyper(gp_length=gp_length, coeffm=np.asarray(params[2:5], dtype=float), gp_amp=params[0])
        nll, _ = self._problem.neg_logl(h)
        return nll

    def optimize_gp(self):
        """Maximise the marginal likelihood over amplitude, length scale and correlation coefficients with
        ``scipy.optimize.shgo`` (host driver, ``inversion.py:155-178``); every evaluation runs on the GPU.

        The reference stores a *scalar* length scale afterwards, which makes its next ``create_cov`` raise
        (SURVEY.md Q2); here the optimised factor is expanded to the three per-property scales."""
        from scipy.optimize import shgo
        print("Optimizing GP hyperparameters and correlation coefficients, this may take a while...")
        gl, gc = _cfg.gp_lengthscale, _cfg.gp_coeff
        res = shgo(self.calc_logl, bounds=((0.5, 2), (0.5 * gl, 10 * gl), (0.5 * gc[0], 1), (0.5 * gc[1], 1), (0.5 * gc[2], 1)),
                   n=10, iters=10, sampling_method='sobol')
        if not res.success:
            print('WARNING: ' + res.message)
        else:
            print("Initial parameter [amplitude, lengthscale, corr1, corr2, corr3]:")
            print(self.gp_amp, self.gp_length, self.coeffm)
            self.gp_amp = res.x[0]
            self.gp_length = res.x[1] * np.asarray([_cfg.xvoxsize, _cfg.xvoxsize, _cfg.xvoxsize])
            self.coeffm = np.asarray([res.x[2:]]).flatten()
            print("Optimized parameter [amplitude, lengthscale, corr1, corr2, corr3]:")
            print(self.gp_amp, self.gp_length, self.coeffm)

    # ------------------------------------------------------------------ cubing (inversion.py:182-248)
    def cubing(self, gravfield, magfield, drillfield, sensor_locations, drilldata0):
        """Joint inversion; returns density_rec, magsus_rec, drill_rec, density_var, magsus_var, drill_var."""
        self.gravfield = gravfield
        self.magfield = magfield
        self.drillfield = drillfield
        self.sensor_locations = sensor_locations
        self.drilldata0 = drilldata0
        xN, yN, zN = _cfg.xNcube, _cfg.yNcube, _cfg.zNcube
        with np.errstate(all="ignore"):          # empty drill data -> NaN statistics, like the reference (Q9)
            grav_mean, grav_std = self.gravfield.mean(), self.gravfield.std()
            gravfield_norm = (self.gravfield - grav_mean) / grav_std
            magn_mean, magn_std = self.magfield.mean(), self.magfield.std()
            magfield_norm = (self.magfield - magn_mean) / magn_std
            drillfield = np.asarray(self.drillfield, dtype=float)
            drill_mean, drill_std = (drillfield.mean(), drillfield.std()) if drillfield.size else (np.nan, np.nan)
            drillfield_norm = (drillfield - drill_mean) / drill_std
        self.Fs3 = np.hstack((gravfield_norm, magfield_norm, drillfield_norm))
        self._build_problem()
        if self.Fs3.size != self._problem.M:
            raise ValueError("drillfield has %d values but drilldata0 has %d non-zero voxels"
                             % (drillfield.size, self._problem.nd))
        if _cfg.optimize_gp:
            self.optimize_gp()
        self.mu_rec, self.cov_rec, self.logl = self.predict3(calclogl=True)
        results_rec = self.mu_rec.reshape(3, yN, xN, zN)
        results_var = self.cov_rec.diagonal().reshape(3, yN, xN, zN)
        density_rec = results_rec[0] * grav_std          # model represents the deviation from the mean
        density_var = results_var[0] * grav_std**2
        magsus_rec = results_rec[1] * magn_std
        magsus_var = results_var[1] * magn_std**2
        drill_rec = results_rec[2] * drill_std
        drill_var = results_var[2] * drill_std**2
        return density_rec, magsus_rec, drill_rec, density_var, magsus_var, drill_var

    # ------------------------------------------------------------------ lazily materialised reference attributes
    def _dense_posterior_cov(self):
        if _dist.world_size() > 1:
            raise NotImplementedError("dense posterior covariance is only available on one GPU")
        return self._problem.posterior_cov(self._hyper())

    def __getattr__(self, name):
        # only called when normal lookup fails: dense arrays the reference stores on the instance
        if name.startswith("_") or name not in ("points3D", "D2", "Asens3", "kcov", "timings"):
            raise AttributeError(name)
        lazy = self.__dict__.setdefault("_lazy", {})
        if name in lazy:
            return lazy[name]
        if name == "timings":
            return self._problem.timings()
        xN, yN, zN = _cfg.xNcube, _cfg.yNcube, _cfg.zNcube
        if name == "points3D":      # inversion.py:216
            val = kernel.calcGridPoints3D((xN, yN, zN), (_cfg.xvoxsize, _cfg.yvoxsize, _cfg.zvoxsize))
        elif name == "D2":          # inversion.py:217
            val = kernel.calcDistanceMatrix(self.points3D)
        elif name == "kcov":        # inversion.py:92 (gp_length already de-duplicated by predict3)
            val = self.gp_amp * kernel.create_cov(self.D2, np.array(self.gp_length, dtype=float), self.coeffm, fkernel=_cfg.kernelfunc)
        else:                       # Asens3, inversion.py:227-230: block-diagonal [grav | magn | drill]
            p = self._problem
            if p is None:
                raise AttributeError(name)
            N, Ns, nd = p.N, p.Ns, p.nd
            val = np.zeros((p.M, 3 * N))
            val[:Ns, :N] = p.sens("grav")
            val[Ns:2 * Ns, N:2 * N] = p.sens("magn")
            val[2 * Ns + np.arange(nd), 2 * N + self._drill_idx] = 1.0
        lazy[name] = val
        return val
