"""Oracle restatement of the EKF baseline of the dense magnetic-field example
(examples/slam-dense-mag/ekf_dense.m:37-102 with the closures dynModel_ekf /
measModel_ekf of run_dense3D_magfield.m:281-316).  Test infrastructure only.

State x = [pos(3); orientation deviation(3); map(M)], linearisation point q_nb (unit
quaternion) carried next to it; x0 = [pos0; 0; x0_lin], P0 = blkdiag(zeros(6), P0_lin)
(run_dense3D_magfield.m:248-250).
"""
import numpy as np

from .tools import expq, qLeft, quat2rmat, mcross, eigenfun_dx, JacobianPhi3D, chol_jitter


def dynModel_ekf(x, q, dx):
    """run_dense3D_magfield.m:310-316"""
    xpred = x.copy()
    xpred[0:3] = x[0:3] + dx[0:3]
    qpred = qLeft(q) @ dx[3:7]
    F = np.eye(x.shape[0])
    G = np.zeros((x.shape[0], 6))
    G[0:3, 0:3] = np.eye(3)
    G[3:6, 3:6] = quat2rmat(qpred)
    return xpred, qpred, F, G


def measModel_ekf(x, q, NN, L, LL):
    """run_dense3D_magfield.m:281-299; LL [2 x 3] = the domain bounds handed to JacobianPhi3D"""
    m = NN.shape[0]
    pos = x[0:3].reshape(1, 3)
    dPhi = np.vstack([np.concatenate([[1.0, 0.0, 0.0], eigenfun_dx(NN, pos, 0, L)[0]]),
                      np.concatenate([[0.0, 1.0, 0.0], eigenfun_dx(NN, pos, 1, L)[0]]),
                      np.concatenate([[0.0, 0.0, 1.0], eigenfun_dx(NN, pos, 2, L)[0]])])
    Rnb = quat2rmat(q)
    yhat = Rnb.T @ dPhi @ x[6:]
    dy = np.zeros((3, dPhi.shape[1] + 6))
    J = JacobianPhi3D(x[0:3].reshape(3, 1), m, LL[0, 0], LL[1, 0], LL[0, 1], LL[1, 1], LL[0, 2], LL[1, 2], NN)
    J = J[:, :, :, 0].reshape(9, m, order="F") @ x[9:]
    J = J.reshape(3, 3, order="F")
    dy[:, 0:3] = Rnb.T @ J
    dy[:, 3:6] = Rnb.T @ mcross(dPhi @ x[6:])
    dy[:, 6:] = Rnb.T @ dPhi
    return yhat, dy


def ekf_dense(NN, L, LL, odometry, y, x0, q0, P0, Q, R, dt, keep_P=False):
    """examples/slam-dense-mag/ekf_dense.m:37-102.  Returns xf_traj [nStates x T],
    qnb_traj [4 x T] and the last filtered covariance (all of them with keep_P)."""
    NN = np.asarray(NN, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64).reshape(3)
    LL = np.asarray(LL, dtype=np.float64).reshape(2, 3)
    y = np.asarray(y, dtype=np.float64)
    N_T = y.shape[0]
    xp = np.asarray(x0, dtype=np.float64).reshape(-1).copy()
    Pp = np.asarray(P0, dtype=np.float64).copy()
    q_nb = np.asarray(q0, dtype=np.float64).reshape(4).copy()
    nStates = xp.shape[0]
    Q = np.asarray(Q, dtype=np.float64)
    if Q.ndim == 2:                                            # :46-48
        Q = np.repeat(Q[:, :, None], max(N_T - 1, 1), axis=2)
    dt = np.asarray(dt, dtype=np.float64).reshape(-1)
    if dt.shape[0] == 1:                                       # :51-53
        dt = dt[0] * np.ones(max(N_T - 1, 1))
    jitter = 1e-3                                              # :56
    xf_traj = np.full((nStates, N_T), np.nan)
    qnb_traj = np.full((4, N_T), np.nan)
    Ps = []
    xf, Pf = None, None
    for t in range(N_T):                                       # :66
        if t != 0:                                             # :68-73
            xp, q_nb, F, G = dynModel_ekf(xf, q_nb, odometry[t - 1, :])
            Qt = dt[t - 1] * Q[:, :, t - 1]
            Pp = F @ Pf @ F.T + G @ Qt @ G.T
        yt = y[t, :]                                           # :76
        yhat, dy = measModel_ekf(xp, q_nb, NN, L, LL)          # :77
        e = yt - yhat
        SS = dy @ Pp @ dy.T + R
        cS, _ = chol_jitter(SS, jitter)                        # :82-85
        K = Pp @ np.linalg.solve(cS.T, np.linalg.solve(cS, dy)).T     # :86  Pp*((dy'/cS')/cS)
        xf = xp + K @ e                                        # :89
        Pf = Pp - K @ SS @ K.T
        Pf = 0.5 * (Pf + Pf.T)                                 # :91
        q_nb = qLeft(expq(xf[3:6] / 2)) @ q_nb                 # :94
        xf[3:6] = 0.0
        xf_traj[:, t] = xf
        qnb_traj[:, t] = q_nb
        if keep_P:
            Ps.append(Pf.copy())
    return xf_traj, qnb_traj, (np.stack(Ps, axis=2) if keep_P else Pf)
