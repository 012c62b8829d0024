"""Injected random streams for the oracle (test infrastructure only).

The reference interleaves ``rand`` (inside ``sample``) and ``randn`` (inside
``dynModel``) per particle (src/particleFilter.m:104-109).  MATLAB's ``randn``
cannot be reproduced outside MATLAB, so both the oracle and the CUDA path
consume explicit arrays:

  U[k, t, i]      uniform used by ``sample`` for particle i at step t of sweep k
                  (for the smoother's reference particle at k>=2 it is the
                  uniform of ``ai(N_P) = sample(paNt)``, src/particleSmoother.m:241)
  Z[k, t, i, :]   the ``nz`` standard normals ``dynModel`` consumes for particle i
  Uend[k]         uniform of the sweep-end draw ``ak = sample(w)`` (:346)

``philox_uniforms_normals`` restates, in NumPy, the counter-based generator the
device uses in its free-running mode (Philox4x32-10, counter = (i, t, block,
sweep), key = seed) so free-running device runs can be compared against an
oracle that is fed the very same numbers.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u53(a, b):
    """53-bit integer from two 32-bit words: (a>>5)*2^26 + (b>>6)."""
    return (a >> np.uint64(5)) * np.uint64(67108864) + (b >> np.uint64(6))


def philox_uniforms_normals(seed, k, t, N, nz):
    """Streams of sweep ``k`` step ``t`` for particles 0..N-1.

    Returns (U [N], Z [N, nz]).  Block 0 of counter (i,t,0,k) yields U; block
    1+p yields the Box-Muller pair (Z[2p], Z[2p+1]) with
    u1=(x+1)/2^53 in (0,1], u2=x/2^53 in [0,1).
    """
    i = np.arange(N, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    x0, x1, _, _ = philox4x32_10(i, t, 0, k, k0, k1)
    U = _u53(x0, x1).astype(np.float64) * 2.0 ** -53
    Z = np.empty((N, nz))
    for p in range((nz + 1) // 2):
        x0, x1, x2, x3 = philox4x32_10(i, t, 1 + p, k, k0, k1)
        ua = (_u53(x0, x1).astype(np.float64) + 1.0) * 2.0 ** -53
        ub = _u53(x2, x3).astype(np.float64) * 2.0 ** -53
        r = np.sqrt(-2.0 * np.log(ua))
        ang = 2.0 * np.pi * ub
        Z[:, 2 * p] = r * np.cos(ang)
        if 2 * p + 1 < nz:
            Z[:, 2 * p + 1] = r * np.sin(ang)
    return U, Z


class Streams:
    """Container of injected uniforms / normals for N_K sweeps of T steps."""

    def __init__(self, U, Z, Uend=None):
        self.U = np.asarray(U, dtype=np.float64)      # [K, T, N]   (row t=0 unused)
        self.Z = np.asarray(Z, dtype=np.float64)      # [K, T, N, nz]
        K = self.U.shape[0]
        self.Uend = np.zeros(K) if Uend is None else np.asarray(Uend, dtype=np.float64)

    @classmethod
    def from_numpy_rng(cls, rng, K, T, N, nz):
        U = rng.random((K, T, N))
        Z = rng.standard_normal((K, T, N, nz))
        Uend = rng.random(K)
        return cls(U, Z, Uend)

    @classmethod
    def from_philox(cls, seed, K, T, N, nz):
        U = np.zeros((K, T, N))
        Z = np.zeros((K, T, N, nz))
        Uend = np.zeros(K)
        for k in range(K):
            for t in range(1, T):
                U[k, t], Z[k, t] = philox_uniforms_normals(seed, k, t, N, nz)
            Uend[k] = philox_uniforms_normals(seed, k, T, 1, 0)[0][0]
        return cls(U, Z, Uend)
