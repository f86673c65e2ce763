"""``geobo.kernels`` surface on the B200 (reference: ``geobo/kernels.py``).

Same function names, positional signatures, defaults and return conventions
(float64, C-contiguous NumPy arrays); every value is computed by the CUDA library.
Quirks kept on purpose: the length-scale de-duplication of
``create_cov`` is coded as in ``kernels.py:174-180`` and mutates an ndarray argument
in place (Q1); ``gpkernel_sparse2`` follows the code, not the paper (Q3).
"""
import numpy as np

from . import _lib


def calcGridPoints3D(Lpix, pixscale):
    """Grid points for the distance matrix (``kernels.py:27-42``): rows ``(iy*xN+ix)*zN+iz``, columns x, y, z."""
    Lpix = np.asarray(Lpix)
    pixscale = np.asarray(pixscale)
    return _lib.default_context().grid_points(Lpix[:3], pixscale[:3])


_DEFAULT_DIST = object()


def calcDistanceMatrix(nDimPoints, distFunc=_DEFAULT_DIST):
    """Matrix of squared distances (``kernels.py:45-61``).  Only the reference's default ``distFunc``
    (sum of squared coordinate differences) is implemented on the device."""
    if distFunc is not _DEFAULT_DIST:
        raise NotImplementedError("geobo_b200.kernels.calcDistanceMatrix only implements the default distFunc")
    return _lib.default_context().sqdist(np.array(nDimPoints, dtype=float))


def _elementwise(kernel, cross, D2, l1, l2=0.0):
    D2 = np.asarray(D2, dtype=float)
    return _lib.default_context().cov_function(kernel, cross, D2, l1, l2).reshape(D2.shape)


def gpkernel(D2, gamma):
    """Squared-exponential kernel (``kernels.py:81-88``)."""
    return _elementwise("exp", 0, D2, gamma)


def gpkernel2(D2, gammas):
    """Squared-exponential cross kernel (``kernels.py:90-99``)."""
    return _elementwise("exp", 1, D2, gammas[0], gammas[1])


def gpkernel_sparse(D2, gamma):
    """Compact-support kernel of Melkumyan & Ramos (``kernels.py:101-114``)."""
    return _elementwise("sparse", 0, D2, gamma)


def gpkernel_sparse2(D2, gammas):
    """Compact-support cross kernel, as coded (``kernels.py:116-138``)."""
    return _elementwise("sparse", 1, D2, gammas[0], gammas[1])


def gpkernel_matern32(D2, gamma):
    """Matern-3/2 kernel (``kernels.py:140-146``)."""
    return _elementwise("matern32", 0, D2, gamma)


def gpkernel_matern32_2(D2, gammas):
    """Matern-3/2 cross kernel (``kernels.py:148-156``); singular for equal scales like the reference."""
    return _elementwise("matern32", 1, D2, gammas[0], gammas[1])


def dedup_lengthscales(params):
    """``kernels.py:174-180`` exactly as coded: in place; the second test rewrites element 1 (Q1)."""
    if params[1] == params[0]:
        params[1] = 1.01 * params[0]
    if params[2] == params[0]:
        params[1] = 1.02 * params[0]
    if params[2] == params[1]:
        params[2] = 1.01 * params[1]
    return params


def create_cov(D2, gplength, crossweights=[1, 1, 1], fkernel='sparse'):
    """Multi-output covariance matrix, 3N x 3N (``kernels.py:158-195``).

    w1: density-drill, w2: magnetic-drill, w3: density-magnetic correlation.  An unknown ``fkernel``
    raises ``UnboundLocalError`` in the reference (no branch assigns the strips); here ``ValueError``.
    """
    params = dedup_lengthscales(np.asarray(gplength))     # aliases an ndarray argument, like the reference
    w = np.asarray(crossweights, dtype=float)
    if fkernel not in _lib.KERNEL_IDS:
        raise ValueError("fkernel must be 'sparse', 'exp' or 'matern32'")
    return _lib.default_context().create_cov(np.asarray(D2, dtype=float), np.asarray(params, dtype=float), w, fkernel)
