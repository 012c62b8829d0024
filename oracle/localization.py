"""Oracle restatement of the localisation-only particle filter
(examples/mag-localization-mapping/particleFilterLocalization.m:50-132) with the model closures of
run_localization.m:241-280.  Test infrastructure only.

The map is FIXED: ``foo`` [M] are the posterior-mean basis weights and ``dVarft`` [>= N_P x 3] the
predictive variances the reference's measModel reads (row i for particle i, run_localization.m:263-270).
"""
import numpy as np

from .tools import expq, qLeft, quat2rmat, eigenfun_dx, sample


def qRight(q):
    """tools/qRight.m:30: [q0 -qv'; qv q0*I - [qv x]]"""
    q0, qv = q[0], q[1:4]
    cr = np.array([[0, -qv[2], qv[1]], [qv[2], 0, -qv[0]], [-qv[1], qv[0], 0]])
    out = np.zeros((4, 4))
    out[0, 0] = q0
    out[0, 1:] = -qv
    out[1:, 0] = qv
    out[1:, 1:] = q0 * np.eye(3) - cr
    return out


def dynModel_loc(xn, dx, dt, Q, z):
    """run_localization.m:274-280 (element-wise sqrt of the 3 x 3 blocks)"""
    pos = xn[0:3] + dx[0:3] + np.sqrt(dt * Q[0:3, 0:3]) @ z[0:3]
    quat = qLeft(qRight(xn[3:7]) @ dx[3:7]) @ expq(np.sqrt(dt * Q[3:6, 3:6]) @ z[3:6])
    return np.concatenate([pos, quat])


def measModel_loc(yt, xn, NN, L, foo, dVarft, sigma2):
    """run_localization.m:241-272: w_i = sum(normpdf(yt, (Rnb_i' * dEft_i')', sqrt(dVarft(i,:) + sigma2)))"""
    N = xn.shape[1]
    pos = xn[0:3, :].T
    dPhix = np.hstack([np.ones((N, 1)), np.zeros((N, 2)), eigenfun_dx(NN, pos, 0, L)])
    dPhiy = np.hstack([np.zeros((N, 1)), np.ones((N, 1)), np.zeros((N, 1)), eigenfun_dx(NN, pos, 1, L)])
    dPhiz = np.hstack([np.zeros((N, 2)), np.ones((N, 1)), eigenfun_dx(NN, pos, 2, L)])
    dEft = np.stack([dPhix @ foo, dPhiy @ foo, dPhiz @ foo], axis=1)
    w = np.full(N, np.nan)
    for i in range(N):
        Rnb = quat2rmat(xn[3:7, i])
        mu = Rnb.T @ dEft[i]
        sd = np.sqrt(dVarft[i, :] + sigma2)
        w[i] = np.sum(np.exp(-0.5 * ((yt - mu) / sd) ** 2) / (sd * np.sqrt(2 * np.pi)))   # normpdf
    return w


def particleFilterLocalization(NN, L, foo, dVarft, sigma2, odometry, y, x0_nonLin, Q, N_P, dt, U, Z, tap=None):
    """particleFilterLocalization.m:50-132.  U [T, N], Z [T, N, 6]: injected uniforms / normals.
    Returns traj_max, traj_mean [7 x T], xn_traj [7 x N x T]."""
    NN = np.asarray(NN, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64).reshape(3)
    y = np.asarray(y, dtype=np.float64)
    N_T = y.shape[0]
    w = 1.0 / N_P * np.ones(N_P)                                   # :53
    x0 = np.asarray(x0_nonLin, dtype=np.float64)
    xn = x0.copy() if x0.ndim == 2 and x0.shape[1] > 1 else np.repeat(x0.reshape(-1, 1), N_P, axis=1)   # :56-60
    Q = np.asarray(Q, dtype=np.float64)
    if Q.ndim == 2:                                                # :67-69
        Q = np.repeat(Q[:, :, None], N_T, axis=2)
    dt = np.asarray(dt, dtype=np.float64).reshape(-1)
    if dt.shape[0] == 1:                                           # :72-74
        dt = dt[0] * np.ones(N_T)
    traj_max = np.full((7, N_T), np.nan)
    traj_mean = np.full((7, N_T), np.nan)
    xn_traj = np.zeros((7, N_P, N_T))
    xn_traj[:, :, 0] = xn
    ai = np.zeros(N_P, dtype=np.int64)
    for t in range(N_T):                                           # :84
        xn_ = xn.copy()
        if t != 0:                                                 # :90-97
            for i in range(N_P):
                ai[i] = sample(w, U[t, i])
                xn[:, i] = dynModel_loc(xn_[:, ai[i]], odometry[t - 1, :], dt[t - 1], Q[:, :, t - 1], Z[t, i])
            xn_traj[:, :, t] = xn                                  # :101-104
            xn_traj[:, :, :t] = xn_traj[:, ai, :t]
        w = measModel_loc(y[t, :], xn, NN, L, foo, dVarft, sigma2)  # :107-110
        w = w / np.sum(w)                                          # :118
        iw_max = int(np.argmax(w))                                 # :121
        traj_max[:, t] = xn[:, iw_max]
        traj_mean[:, t] = np.sum(xn * w, axis=1)
        if tap is not None:
            tap(t, dict(w=w.copy(), ai=ai.copy(), xn=xn.copy()))
    return traj_max, traj_mean, xn_traj
