"""Oracle restatement of the reference's L0 math primitives (tools/*.m).

Test infrastructure only (see oracle/__init__.py).  Every function cites the
reference lines it restates; paths are relative to /root/reference.
"""
import numpy as np


# --------------------------------------------------------------------------
# tools/sample.m:30-32
# --------------------------------------------------------------------------
def sample(w, u, clamp=True):
    """Multinomial draw.  ``wc=cumsum(w); ind=sum(wc<u)+1`` (tools/sample.m:30-32).

    ``np.cumsum`` accumulates strictly left to right in fp64, like MATLAB's
    ``cumsum``.  Returns a 0-based index.  The reference would raise an index
    error when ``u > wc(end)`` (``ind = N+1``); the device clamps to the last
    particle and so does the oracle when ``clamp`` is set (documented deviation).
    """
    wc = np.cumsum(np.asarray(w, dtype=np.float64))
    ind = int(np.count_nonzero(wc < u))
    if ind >= wc.shape[0]:
        if not clamp:
            raise IndexError("sample: u exceeds cumsum(w)(end) (reference would error)")
        ind = wc.shape[0] - 1
    return ind


def sample_many(w, us, clamp=True):
    """``sample`` for a vector of uniforms against the same weights (same wc)."""
    wc = np.cumsum(np.asarray(w, dtype=np.float64))
    # number of wc strictly below u == lower-bound position of u in wc when wc is
    # non-decreasing; wc IS non-decreasing for w>=0 so searchsorted(left) is exact.
    ind = np.searchsorted(wc, np.asarray(us, dtype=np.float64), side="left")
    if clamp:
        ind = np.minimum(ind, wc.shape[0] - 1)
    elif np.any(ind >= wc.shape[0]):
        raise IndexError("sample: u exceeds cumsum(w)(end)")
    return ind.astype(np.int64)


# --------------------------------------------------------------------------
# quaternion algebra
# --------------------------------------------------------------------------
def mcross(v):
    """Cross-product matrix, single-vector branch (tools/mcross.m:33-36)."""
    v = np.asarray(v, dtype=np.float64).reshape(3)
    return np.array([[0.0, -v[2], v[1]],
                     [v[2], 0.0, -v[0]],
                     [-v[1], v[0], 0.0]])


def expq(phi):
    """Quaternion exponential, single-vector branch (tools/expq.m:22-32).

    The argument is phi (not phi/2).  ``mag_phi + (mag_phi == 0)`` guards the
    division; the sign is flipped iff the scalar part is strictly negative.
    """
    phi = np.asarray(phi, dtype=np.float64).reshape(3)
    mag = np.sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2])
    nphi = phi / (mag + (1.0 if mag == 0 else 0.0))
    eq = np.empty(4)
    eq[0] = np.cos(mag)
    eq[1:] = nphi * np.sin(mag)
    if eq[0] < 0:
        eq = -eq
    return eq


def qLeft(q):
    """Left quaternion-product matrix, single branch (tools/qLeft.m:30-34)."""
    q = np.asarray(q, dtype=np.float64).reshape(4)
    pL = np.empty((4, 4))
    pL[0, 0] = q[0]
    pL[0, 1:] = -q[1:]
    pL[1:, 0] = q[1:]
    pL[1:, 1:] = q[0] * np.eye(3) + mcross(q[1:])
    return pL


def quat2rmat(q):
    """Rotation matrix of a quaternion, no normalisation (tools/quat2rmat.m:27-32)."""
    q0, q1, q2, q3 = np.asarray(q, dtype=np.float64).reshape(4)
    return np.array([
        [q0**2 + q1**2 - q2**2 - q3**2, 2*q1*q2 - 2*q0*q3, 2*q1*q3 + 2*q0*q2],
        [2*q1*q2 + 2*q0*q3, q0**2 - q1**2 + q2**2 - q3**2, 2*q2*q3 - 2*q0*q1],
        [2*q1*q3 - 2*q0*q2, 2*q2*q3 + 2*q0*q1, q0**2 - q1**2 - q2**2 + q3**2]])


def quat2rmat_batch(q):
    """Batched branch (tools/quat2rmat.m:34-39): q [N x 4] -> R [3 x 3 x N]."""
    q = np.asarray(q, dtype=np.float64)
    R = np.empty((3, 3, q.shape[0]))
    for i in range(q.shape[0]):
        R[:, :, i] = quat2rmat(q[i])
    return R


def qInv(q):
    """Quaternion conjugate (tools/qInv.m:27-31)."""
    q = np.array(q, dtype=np.float64).reshape(4)
    q[1:] = -q[1:]
    return q


def logq(q, clamp=True):
    """Quaternion logarithm, single branch (tools/logq.m:25-30).

    Hazard Q6: q is never renormalised, so q0 can exceed 1 by
    round-off and MATLAB's acos turns complex.  The device clamps q0 to [-1,1];
    the oracle does the same when ``clamp`` (documented deviation).
    """
    q = np.array(q, dtype=np.float64).reshape(4)
    if q[0] < 0:
        q = -q
    q0 = min(q[0], 1.0) if clamp else q[0]
    na = np.arccos(q0)
    return na * q[1:] / (np.sin(na) + (1.0 if na == 0 else 0.0))


# --------------------------------------------------------------------------
# Cholesky with MATLAB's two-output semantics
# --------------------------------------------------------------------------
def chol_lower(A):
    """``[cS,flag] = chol(A,'lower')``: returns (L, flag), flag>0 if A is not PD.

    An empty matrix factors to an empty matrix with flag 0 (MATLAB semantics,
    needed for time steps with no observed landmark, src/particleFilter.m:134-148).
    """
    A = np.asarray(A, dtype=np.float64)
    if A.size == 0:
        return np.zeros((0, 0)), 0
    try:
        return np.linalg.cholesky(A), 0
    except np.linalg.LinAlgError:
        return None, 1


def chol_jitter(A, jitter):
    """chol with the reference's retry (src/particleFilter.m:145-148).

    Returns (L, used_jitter).  A second failure raises, as MATLAB would.
    """
    L, flag = chol_lower(A)
    if flag > 0:
        L, flag2 = chol_lower(A + jitter * np.eye(A.shape[0]))
        if flag2 > 0:
            raise np.linalg.LinAlgError("chol failed even with jitter (reference would error)")
        return L, True
    return L, False


def solve_lower(L, b):
    """``L\\b`` for lower-triangular L (forward substitution via LAPACK trtrs)."""
    if L.shape[0] == 0:
        return np.zeros_like(b)
    from scipy.linalg import solve_triangular
    return solve_triangular(L, b, lower=True, check_finite=False)


# --------------------------------------------------------------------------
# tools/domain_cartesian_dx.m
# --------------------------------------------------------------------------
def _ndgridm(N):
    """Index hypercube (tools/domain_cartesian_dx.m:195-216): first index slowest."""
    N = [int(v) for v in N]
    if len(N) == 1:
        return np.arange(1, N[0] + 1, dtype=np.float64).reshape(-1, 1)
    nn = _ndgridm(N[1:])
    rest = int(np.prod(N[1:]))
    NN = np.zeros((N[0] * rest, len(N)))
    NN[:, 0] = np.kron(np.arange(1, N[0] + 1), np.ones(rest))
    NN[:, 1:] = np.tile(nn, (N[0], 1))
    return NN


def eigenval(NN, L):
    """Eigenvalues ``sum((pi*n./(2L)).^2,2)`` (tools/domain_cartesian_dx.m:40)."""
    NN = np.asarray(NN, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64).reshape(1, -1)
    return np.sum((np.pi * NN / (2 * L)) ** 2, axis=1)


def domain_cartesian_dx(m, d, LL):
    """Index selection of the Laplace eigenbasis (tools/domain_cartesian_dx.m:26-43).

    ``LL`` is either the half-widths [d] or the 2 x d matrix of [min; max] rows
    (then L = (max-min)/2, :27-29).  Returns (L, NN) with NN [m x d] as float64
    integers; the sort is stable like MATLAB's ``sort``.
    """
    LL = np.asarray(LL, dtype=np.float64)
    if LL.ndim == 2 and LL.shape[0] > 1:
        L = (LL.max(axis=0) - LL.min(axis=0)) / 2
    else:
        L = LL.reshape(-1)
    N = np.ceil(m ** (1.0 / d) * L / L.min())
    NN = _ndgridm(N)
    lam = eigenval(NN, L)
    ind = np.argsort(lam, kind="stable")
    return L, NN[ind[:m], :]


def eigenfun(NN, x, L):
    """Phi [n_x x m] (tools/domain_cartesian_dx.m:84-93).

    ``v = v .* 1./sqrt(L(j)) .* sin(pi*n.*(x+L)/(2*L))`` evaluated left to right.
    """
    NN = np.asarray(NN, dtype=np.float64)
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    v = np.ones((x.shape[0], NN.shape[0]))
    for j in range(NN.shape[1]):
        arg = (np.pi * NN[None, :, j]) * (x[:, j:j + 1] + L[j]) / (2 * L[j])
        v = v * 1.0 / np.sqrt(L[j]) * np.sin(arg)
    return v


def eigenfun_dx(NN, x, di, L):
    """d Phi / d x_di [n_x x m] (tools/domain_cartesian_dx.m:142-170); di is 0-based."""
    NN = np.asarray(NN, dtype=np.float64)
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    v = np.ones((x.shape[0], NN.shape[0]))
    for j in range(NN.shape[1]):
        arg = (np.pi * NN[None, :, j]) * (x[:, j:j + 1] + L[j]) / (2 * L[j])
        if j == di:
            v = v * np.pi * NN[None, :, j] / (2 * L[j] * np.sqrt(L[j])) * np.cos(arg)
        else:
            v = v * 1.0 / np.sqrt(L[j]) * np.sin(arg)
    return v


def JacobianPhi3D(x, N_m, xl, xu, yl, yu, zl, zu, Indices):
    """J [3 x 3 x N_m x N]: Hessian of each basis function at each point
    (tools/JacobianPhi3D.m:29-64; called from run_dense3D_magfield.m:292-294).

    x is 3 x N.  Products are evaluated left to right as MATLAB does.
    """
    j = np.asarray(Indices, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64).reshape(3, -1)
    N = x.shape[1]
    J = np.zeros((3, 3, N_m, N))
    a = np.array([xl, yl, zl], dtype=np.float64)          # :38
    b = np.array([xu, yu, zu], dtype=np.float64)          # :39
    f = np.zeros((N_m, 3))
    for d in range(3):                                    # :41-44
        f[:, d] = (np.pi * j[:, d]) / (b[d] - a[d])
    for i in range(N):                                    # :47
        s = np.zeros((N_m, 3))
        c = np.zeros((N_m, 3))
        for d in range(3):                                # :51-56
            core = np.pi * j[:, d] * (x[d, i] - a[d]) / (b[d] - a[d])
            mult = 1.0 / np.sqrt(0.5 * (b[d] - a[d]))
            s[:, d] = np.sin(core) * mult
            c[:, d] = np.cos(core) * mult
        J[0, 0, :, i] = -f[:, 0] ** 2 * s[:, 0] * s[:, 1] * s[:, 2]       # :58-66
        J[0, 1, :, i] = f[:, 0] * f[:, 1] * c[:, 0] * c[:, 1] * s[:, 2]
        J[0, 2, :, i] = f[:, 0] * f[:, 2] * c[:, 0] * s[:, 1] * c[:, 2]
        J[1, 0, :, i] = f[:, 1] * f[:, 0] * c[:, 0] * c[:, 1] * s[:, 2]
        J[1, 1, :, i] = -f[:, 1] ** 2 * s[:, 0] * s[:, 1] * s[:, 2]
        J[1, 2, :, i] = f[:, 1] * f[:, 2] * s[:, 0] * c[:, 1] * c[:, 2]
        J[2, 0, :, i] = f[:, 2] * f[:, 0] * c[:, 0] * s[:, 1] * c[:, 2]
        J[2, 1, :, i] = f[:, 2] * f[:, 1] * s[:, 0] * c[:, 1] * c[:, 2]
        J[2, 2, :, i] = -f[:, 2] ** 2 * s[:, 0] * s[:, 1] * s[:, 2]
    return J
