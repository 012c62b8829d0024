"""Synthetic SLAM problems of the reference's shapes (host side, NumPy; not on the hot path).

The recipes follow the reference's data simulators so that trajectories, fields
and noise levels have the shape the paper's experiments use:
  * trajectories  examples/slam-dense-radio/generateData_dense.m:66-214
                  ('bean_6D' :181-213, 'line_3D' :128-147, 'square_3D' :117-127)
  * GP field      tools/gp_rnd_scalar_potential_fast.m:45-102 (3-D curl-free
                  potential), tools/gp_rnd_SE1D_fast.m:47-85 (2-D scalar field)
  * odometry      noisy forward pass of dynModel, generateData_dense.m:294-325
All randomness comes from ``numpy.random.default_rng(seed)`` (MATLAB streams are
not reproducible); nothing here is called from the per-step path.
"""
import numpy as np
from .basis import domain_cartesian_dx, eigenvalues, spectral_density_se


# ----------------------------------------------------------------------------
# small host-side helpers (NumPy)
# ----------------------------------------------------------------------------
def _sincos_arg(NN, x, L, j):
    return (np.pi * NN[None, :, j]) * (x[:, j:j + 1] + L[j]) / (2 * L[j])


def basis_phi(NN, x, L):
    """Phi(x) [n_x x m]: prod_j L_j^-1/2 sin(pi n_j (x_j+L_j)/(2 L_j))."""
    NN = np.asarray(NN, dtype=np.float64)
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    v = np.ones((x.shape[0], NN.shape[0]))
    for j in range(NN.shape[1]):
        v *= np.sin(_sincos_arg(NN, x, L, j)) / np.sqrt(L[j])
    return v


def basis_dphi(NN, x, L, di):
    """d Phi / d x_di [n_x x m]."""
    NN = np.asarray(NN, dtype=np.float64)
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    v = np.ones((x.shape[0], NN.shape[0]))
    for j in range(NN.shape[1]):
        a = _sincos_arg(NN, x, L, j)
        if j == di:
            v *= np.pi * NN[None, :, j] / (2 * L[j] * np.sqrt(L[j])) * np.cos(a)
        else:
            v *= np.sin(a) / np.sqrt(L[j])
    return v


def _qmul(p, q):
    """Hamilton product p (x) q for [..., 4] arrays, scalar first."""
    p0, p1, p2, p3 = np.moveaxis(p, -1, 0)
    q0, q1, q2, q3 = np.moveaxis(q, -1, 0)
    return np.stack([p0*q0 - p1*q1 - p2*q2 - p3*q3,
                     p0*q1 + p1*q0 + p2*q3 - p3*q2,
                     p0*q2 - p1*q3 + p2*q0 + p3*q1,
                     p0*q3 + p1*q2 - p2*q1 + p3*q0], axis=-1)


def _qconj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def _expq(phi):
    phi = np.atleast_2d(phi)
    mag = np.sqrt(np.sum(phi ** 2, axis=1))
    nphi = phi / (mag + (mag == 0))[:, None]
    q = np.concatenate([np.cos(mag)[:, None], nphi * np.sin(mag)[:, None]], axis=1)
    q[q[:, 0] < 0] *= -1.0
    return q


def _rot_nb(q):
    """Body->nav rotation matrices [T,3,3] of quaternions [T,4] (tools/quat2rmat.m)."""
    q0, q1, q2, q3 = q.T
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = q0**2 + q1**2 - q2**2 - q3**2
    R[:, 0, 1] = 2*q1*q2 - 2*q0*q3
    R[:, 0, 2] = 2*q1*q3 + 2*q0*q2
    R[:, 1, 0] = 2*q1*q2 + 2*q0*q3
    R[:, 1, 1] = q0**2 - q1**2 + q2**2 - q3**2
    R[:, 1, 2] = 2*q2*q3 - 2*q0*q1
    R[:, 2, 0] = 2*q1*q3 - 2*q0*q2
    R[:, 2, 1] = 2*q2*q3 + 2*q0*q1
    R[:, 2, 2] = q0**2 - q1**2 - q2**2 + q3**2
    return R


# ----------------------------------------------------------------------------
# C1 / C4 / C5: 6-D pose, magnetic-field potential map
# ----------------------------------------------------------------------------
DEFAULT_MAG_THETA = np.array([650.0, 1.2, 200.0, 10.0])          # slam-dense-mag/main.m:23
DEFAULT_MAG_Q = np.diag(np.concatenate([
    10.0 ** 2 * np.array([0.05, 0.05, 0.01]) ** 2,
    (np.array([0.01, 0.01, 0.3]) * np.pi / 180.0) ** 2]))         # slam-dense-mag/main.m:22


def bean_6d_trajectory(N_T, n_laps=3, a=15.0):
    """'bean_6D' path (generateData_dense.m:181-213) sampled with N_T points."""
    psi = np.linspace(0.0, n_laps * np.pi, N_T)
    r = a * np.sin(psi) ** 3 + a * np.cos(psi) ** 3
    u = r * np.cos(psi) - 0.3
    v = r * np.sin(psi) - 0.3
    th = np.arctan2(np.diff(v), np.diff(u))
    th = np.concatenate([th, th[-1:]])
    pos = np.stack([u, v, np.zeros_like(u)], axis=0)
    # R = [cos th, sin th; -sin th, cos th] is a rotation by -th about z ->
    # q = expq(logR(R)/2) = [cos(th/2), 0, 0, -sin(th/2)] up to the sign flip
    quat = _expq(np.stack([0 * th, 0 * th, -th / 2.0], axis=1))
    pos = pos - 0.5 * (pos.min(axis=1, keepdims=True) + pos.max(axis=1, keepdims=True))
    return pos, quat


def dense_mag_problem(N_T=192, m=512, seed=1, theta=None, Q=None, dt=0.01, n_laps=3, a=15.0,
                      m_sim=2000, nLL=2.0):
    """Magnetic-field SLAM inputs (config C1 with defaults; C4: N_T=2000, m=1024, n_laps=10).

    Returns a dict with odometry [T-1 x 7] (padded to T rows like the reference's
    dx: row t-1 is used at step t), y [T x 3], x0_nonLin [7], x0_lin [M], P0_lin
    [M x M], Q [6 x 6], R [3 x 3], dt, NN [m x 3] int32, L [3] and ground truth.
    """
    rng = np.random.default_rng(seed)
    theta = DEFAULT_MAG_THETA if theta is None else np.asarray(theta, dtype=np.float64)
    Q = DEFAULT_MAG_Q if Q is None else np.asarray(Q, dtype=np.float64)
    pos, quat = bean_6d_trajectory(N_T, n_laps, a)
    ls = theta[1]
    LL = np.array([[pos[0].min() - nLL * ls, pos[1].min() - nLL * ls, -nLL * ls],
                   [pos[0].max() + nLL * ls, pos[1].max() + nLL * ls, nLL * ls]])
    # --- field draw with m_sim basis functions (gp_rnd_scalar_potential_fast.m:45-102)
    centre = LL.mean(axis=0)
    Ls, NNs = domain_cartesian_dx(m_sim, 3, LL)
    lam_s = eigenvalues(NNs, Ls)
    k_s = np.concatenate([np.full(3, theta[0]), spectral_density_se(lam_s, ls, theta[2], 3)])
    coef = np.sqrt(k_s) * rng.standard_normal(k_s.shape[0])
    xs = pos.T - centre
    ones = np.ones((N_T, 1))
    zeros = np.zeros((N_T, 1))
    dPx = np.hstack([ones, zeros, zeros, basis_dphi(NNs, xs, Ls, 0)])
    dPy = np.hstack([zeros, ones, zeros, basis_dphi(NNs, xs, Ls, 1)])
    dPz = np.hstack([zeros, zeros, ones, basis_dphi(NNs, xs, Ls, 2)])
    df = np.stack([dPx @ coef, dPy @ coef, dPz @ coef], axis=1)
    yn = df + np.sqrt(theta[3]) * rng.standard_normal(df.shape)
    Rnb = _rot_nb(quat)
    y = np.einsum("tji,tj->ti", Rnb, yn)           # y_t = R_t' * yn_t (generateData_dense.m:253-257)
    # --- odometry: noisy forward pass of dynModel (generateData_dense.m:302-309)
    dpos = np.diff(pos.T, axis=0)
    dquat = _qmul(_qconj(quat[:-1]), quat[1:])
    Lp = np.linalg.cholesky(dt * Q[0:3, 0:3])
    Lq = np.linalg.cholesky(dt * Q[3:6, 3:6])
    x = np.zeros((N_T, 7))
    x[0] = np.concatenate([pos[:, 0], quat[0]])
    dQ = np.zeros((N_T - 1, 4))
    for i in range(1, N_T):
        x[i, 0:3] = x[i - 1, 0:3] + dpos[i - 1] + Lp @ rng.standard_normal(3)
        dQ[i - 1] = _qmul(dquat[i - 1], _expq(Lq @ rng.standard_normal(3))[0])
        x[i, 3:7] = _qmul(x[i - 1, 3:7], dQ[i - 1])
    odo = np.hstack([np.diff(x[:, 0:3], axis=0), dQ])
    odometry = np.vstack([odo, np.zeros((1, 7))])
    # --- estimation model (run_dense3D_magfield.m:85-131)
    L, NN = domain_cartesian_dx(m, 3, LL)
    lam = eigenvalues(NN, L)
    k = np.concatenate([np.full(3, theta[0]), spectral_density_se(lam, ls, theta[2], 3)])
    M = m + 3
    return dict(family="dense_mag3d", odometry=odometry, y=y, x0_nonLin=x[0].copy(),
                x0_lin=np.zeros(M), P0_lin=np.diag(k), Q=Q, R=theta[3] * np.eye(3), dt=dt,
                NN=NN, L=L, LL=LL, truth=dict(pos=pos, quat=quat, odometry_path=x))


# ----------------------------------------------------------------------------
# C2: 2-D position + heading, scalar RSS field
# ----------------------------------------------------------------------------
def dense_radio_problem(traj="line_3D", m=128, seed=1, theta=(0.25, 1.0, 0.01), m_sim=2000,
                        nLL=2.0):
    """Radio SLAM inputs (run_dense2D_withHeading.m:64-147, generateData_dense.m:117-147,258-325)."""
    rng = np.random.default_rng(seed)
    theta = np.asarray(theta, dtype=np.float64)
    if traj == "line_3D":
        N = 32
        Qv = 1e-6 * np.ones(N)
        Qv[N // 2 - 1] = 0.3 ** 2
        pos = np.stack([np.zeros(N), np.concatenate([np.linspace(0, 3, N // 2),
                                                     np.linspace(3, 0, N // 2)])])
    elif traj == "square_3D":
        N = 48
        Qv = 1e-6 * np.ones(N)
        Qv[N // 4 + N // 4 * np.arange(3) - 1] = 0.1 ** 2
        q = N // 4
        pos = np.stack([
            np.concatenate([np.zeros(q), np.linspace(0, 2, q), 2 * np.ones(q), np.linspace(2, 0, q)]),
            np.concatenate([np.linspace(0, 2, q), 2 * np.ones(q), np.linspace(2, 0, q), np.zeros(q)])])
    else:
        raise ValueError(traj)
    pos = pos - pos.mean(axis=1, keepdims=True)
    ls = theta[0]
    LL = np.array([[pos[0].min() - nLL * ls, pos[1].min() - nLL * ls],
                   [pos[0].max() + nLL * ls, pos[1].max() + nLL * ls]])
    centre = LL.mean(axis=0)
    Ls, NNs = domain_cartesian_dx(m_sim, 2, LL)
    k_s = spectral_density_se(eigenvalues(NNs, Ls), ls, theta[1], 2)
    coef = np.sqrt(k_s) * rng.standard_normal(k_s.shape[0])
    f = basis_phi(NNs, pos.T - centre, Ls) @ coef
    y = (f + np.sqrt(theta[2]) * rng.standard_normal(f.shape)).reshape(-1, 1)
    dx = np.hstack([np.diff(pos.T, axis=0), np.zeros((N - 1, 1))])
    dt = 1.0
    x = np.zeros((N, 3))
    x[0] = np.array([pos[0, 0], pos[1, 0], 0.0])
    for i in range(1, N):
        c, s = np.cos(x[i - 1, 2]), np.sin(x[i - 1, 2])
        x[i, 0] = x[i - 1, 0] + c * dx[i - 1, 0] + s * dx[i - 1, 1]
        x[i, 1] = x[i - 1, 1] - s * dx[i - 1, 0] + c * dx[i - 1, 1]
        x[i, 2] = x[i - 1, 2] + dx[i - 1, 2] + np.sqrt(dt * Qv[i - 1]) * rng.standard_normal()
    odo = np.hstack([dx[:, 0:2], np.diff(x[:, 2]).reshape(-1, 1)])
    odometry = np.vstack([odo, np.zeros((1, 3))])
    L, NN = domain_cartesian_dx(m, 2, LL)
    k = spectral_density_se(eigenvalues(NN, L), ls, theta[1], 2)
    return dict(family="dense_radio2d", odometry=odometry, y=y, x0_nonLin=x[0].copy(),
                x0_lin=np.zeros(m), P0_lin=np.diag(k), Q=Qv.reshape(1, 1, N),
                R=theta[2] * np.eye(1), dt=dt, NN=NN, L=L, LL=LL, truth=dict(pos=pos, f=f))


# ----------------------------------------------------------------------------
# C3: sparse visual SLAM (shape of examples/slam-sparse-visual)
# ----------------------------------------------------------------------------
def sparse_visual_problem(N_T=197, n_landmarks=20, N_P=100, seed=1, f=1.5, fp=0.0, fw=1.0,
                          noise_var=0.01 ** 2, init_map_var=1.0, guess_map_var=0.0,
                          pos_var=0.01 ** 2, angle_var=0.001 ** 2, fixture=None, pos_bias=0.0,
                          meas_noise_std=None):
    """Visual SLAM inputs shaped like pfslam.m:78-97 / load_data.m:58-89.

    With ``fixture`` (a dict holding the arrays of the reference's
    ``curve-x2.mat``: Yclean, map, p, th) the real path/landmarks are used;
    otherwise a loop path with a ring of landmarks is synthesised.  Unobserved
    landmarks are NaN in ``y`` (behind the camera or outside the image width).
    ``c3_problem`` fills in the example's own parameter values.
    """
    rng = np.random.default_rng(seed)
    if fixture is not None:
        p = np.asarray(fixture["p"], dtype=np.float64)
        th = np.asarray(fixture["th"], dtype=np.float64).reshape(-1)
        lm = np.asarray(fixture["map"], dtype=np.float64)
        Yclean = np.asarray(fixture["Yclean"], dtype=np.float64)
        N_T = p.shape[1]
        n_landmarks = lm.shape[1]
    else:
        s = np.linspace(0, 2 * np.pi, N_T)
        p = np.stack([2 * np.cos(s), 2 * np.sin(s)])
        th = s + np.pi / 2
        ang = np.linspace(0, 2 * np.pi, n_landmarks, endpoint=False)
        lm = np.stack([3.5 * np.cos(ang), 3.5 * np.sin(ang)])
        Yclean = np.full((n_landmarks, N_T), np.nan)
        for t in range(N_T):
            c, sn = np.cos(th[t]), np.sin(th[t])
            lx = c * (lm[0] - p[0, t]) + sn * (lm[1] - p[1, t])
            ly = -sn * (lm[0] - p[0, t]) + c * (lm[1] - p[1, t])
            yy = (f * lx + fp * ly) / ly
            vis = (ly > 0) & (np.abs(yy) <= fw)
            Yclean[vis, t] = yy[vis]
    # load_data.m:86 adds noise of a fixed 0.01 std while R = noiseVar*I comes from main.m:28
    sd = np.sqrt(noise_var) if meas_noise_std is None else meas_noise_std
    Y = Yclean + sd * rng.standard_normal(Yclean.shape)
    dth = np.diff(np.unwrap(th))
    u = np.hstack([np.diff(p, axis=1).T, dth.reshape(-1, 1)])
    u[:, 0:2] += np.sqrt(pos_var) * rng.standard_normal((N_T - 1, 2)) + pos_bias   # load_data.m:82
    u[:, 2] += np.sqrt(angle_var) * rng.standard_normal(N_T - 1)
    odometry = np.vstack([u, np.zeros((1, 3))])
    M = 2 * n_landmarks
    x0_lin = lm.T.reshape(-1)[:, None] + np.sqrt(guess_map_var) * rng.standard_normal((M, N_P))
    return dict(family="sparse_visual2d", odometry=odometry, y=Y.T.copy(),
                x0_nonLin=np.array([p[0, 0], p[1, 0], th[0]]), x0_lin=x0_lin,
                P0_lin=init_map_var * np.eye(M), Q=np.diag([0.1 ** 2, 0.1 ** 2, 0.001 ** 2]),
                R=noise_var * np.eye(n_landmarks), dt=1.0, camera=(f, fp, fw),
                n_landmarks=n_landmarks, truth=dict(p=p, th=th, map=lm))


def c3_problem(fixture, N_P=100, seed=42):
    """The sparse visual-SLAM example as its runner configures it (C3): the path, landmarks and
    noise-free observations of ``curve-x2.mat`` (``fixture``), odometry noise + drift and
    observation noise of load_data.m:44-54,80-86, map / noise scales of
    slam-sparse-visual/main.m:27-29, Q and x0_lin of pfslam.m:89-95.  NumPy random numbers
    replace MATLAB's (the streams cannot be reproduced)."""
    return sparse_visual_problem(N_P=N_P, seed=seed, f=1.5, fp=0.0, fw=1.0, noise_var=0.1 ** 2,
                                 init_map_var=4.0 ** 2, guess_map_var=1.0 ** 2, pos_var=0.04 ** 2,
                                 angle_var=(0.001 ** 2) ** 2, fixture=fixture, pos_bias=0.01,
                                 meas_noise_std=0.01)
