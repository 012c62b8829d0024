"""Host-side setup of the reduced-rank GP eigenbasis (one-off, not on the hot path).

Mirrors what a MATLAB caller gets from ``tools/domain_cartesian_dx.m:26-43`` of
the reference: the half-widths ``L`` of the (centred) box domain and the integer
index tuples ``NN`` of the ``m`` Laplace eigenfunctions with the smallest
eigenvalues (stable ascending order).  The per-particle evaluation of the basis
itself happens on the GPU (csrc/basis.cuh); only indices and constants are
prepared here.
"""
import numpy as np


def domain_halfwidths(LL):
    """``L=(max-min)/2`` when ``LL`` is a [2 x d] min/max matrix (domain_cartesian_dx.m:27-29)."""
    LL = np.asarray(LL, dtype=np.float64)
    if LL.ndim == 2 and LL.shape[0] > 1:
        return (LL.max(axis=0) - LL.min(axis=0)) / 2.0
    return LL.reshape(-1).copy()


def eigenvalues(NN, L):
    """lambda_j = sum_k (pi n_jk / (2 L_k))^2 (domain_cartesian_dx.m:40)."""
    NN = np.asarray(NN, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64).reshape(1, -1)
    return np.sum((np.pi * NN / (2.0 * L)) ** 2, axis=1)


def domain_cartesian_dx(m, d, LL):
    """Return (L [d] float64, NN [m x d] int32) of the m lowest eigenfunctions.

    Candidate grid ``ceil(m^(1/d) L / min L)`` per dimension, tuples enumerated
    with the first index slowest (as ``ndgridm``, domain_cartesian_dx.m:195-216),
    stable sort by eigenvalue, first m kept (:33-43).
    """
    L = domain_halfwidths(LL)
    if L.shape[0] != d:
        raise ValueError("domain has %d dims, expected %d" % (L.shape[0], d))
    counts = np.ceil(m ** (1.0 / d) * L / L.min()).astype(np.int64)
    if int(np.prod(counts)) < m:
        raise ValueError("candidate grid smaller than m")
    grids = np.indices(tuple(int(c) for c in counts)).reshape(d, -1).T + 1   # C order: first slowest
    lam = eigenvalues(grids, L)
    order = np.argsort(lam, kind="stable")[:m]
    return L, np.ascontiguousarray(grids[order], dtype=np.int32)


def spectral_density_se(lam, length_scale, magn_sigma2, d):
    """Squared-exponential spectral density at w=sqrt(lambda).

    ``magnSigma2*sqrt(2*pi)^d*lengthScale^d*exp(-w.^2*lengthScale^2/2)``
    (examples/slam-dense-mag/run_dense3D_magfield.m:103-104).
    """
    w = np.sqrt(np.asarray(lam, dtype=np.float64))
    return magn_sigma2 * np.sqrt(2 * np.pi) ** d * length_scale ** d * np.exp(-w ** 2 * length_scale ** 2 / 2)
