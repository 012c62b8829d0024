"""Oracle restatement of src/particleFilter.m (test infrastructure only).

Loop structure and arithmetic follow src/particleFilter.m:52-233 line by line;
comments give the reference lines.  P is stored [N, M, M] (particle first) but
means the reference's P(:,:,i).
"""
import numpy as np
from .tools import sample, chol_jitter, solve_lower

LOG2PI = np.log(2 * np.pi)


def _expand_Q_dt(Q, dt, N_T):
    Q = np.asarray(Q, dtype=np.float64)
    if Q.ndim < 2:
        Q = Q.reshape(1, 1)
    if Q.ndim == 2:                       # :75-77
        Q = np.repeat(Q[:, :, None], max(N_T - 1, 1), axis=2)
    dt = np.asarray(dt, dtype=np.float64).reshape(-1)
    if dt.size == 1:                      # :80-82
        dt = dt[0] * np.ones(max(N_T - 1, 1))
    return Q, dt


def innovation(yt, dyt, xl_i, P_i, R, yhat=None):
    """e, SS (and the observed-row mask) of one particle.

    Dense: src/particleFilter.m:139-141.  Sparse (yhat given): :129-136 with the
    unobserved (NaN) rows stripped.
    """
    if yhat is None:
        e = yt - dyt @ xl_i
        SS = dyt @ P_i @ dyt.T + R
        ind = np.ones(e.shape[0], dtype=bool)
    else:
        e = yt - yhat
        SS = dyt @ P_i @ dyt.T + R
        ind = ~np.isnan(yt)
        e = e[ind]
        SS = SS[np.ix_(ind, ind)]
    return e, SS, ind


def log_weight(e, SS, jitter):
    """src/particleFilter.m:144-150.  Returns (logw, cS)."""
    cS, _ = chol_jitter(SS, jitter)
    v = solve_lower(cS, e)
    logw = -np.sum(np.log(np.diag(cS))) - 0.5 * (v @ v) - 0.5 * e.size * LOG2PI
    return logw, cS


def kalman_gain(P_i, dyt_obs, cS):
    """K = P*((dy'/cS')/cS)  (src/particleFilter.m:180,194)."""
    if cS.shape[0] == 0:
        return np.zeros((P_i.shape[0], 0))
    from scipy.linalg import solve_triangular
    # dy'/cS'  ==  solve  X cS' = dy'  ==  cS X' = dy  -> X' = cS\dy
    X1 = solve_triangular(cS, dyt_obs, lower=True, check_finite=False).T      # dy'/cS'
    # X1/cS  ==  solve  Y cS = X1  ==  cS' Y' = X1'
    X2 = solve_triangular(cS.T, X1.T, lower=False, check_finite=False).T      # (dy'/cS')/cS
    return P_i @ X2


def normalise(logw):
    """Log-sum-exp normalisation (src/particleFilter.m:153-156)."""
    c = np.max(logw)
    lse = c + np.log(np.sum(np.exp(logw - c)))
    return np.exp(logw - lse)


def particleFilter(model, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, dt,
                   streams, sparseFeatures=None, makePlots=None, sweep=0,
                   forced_ancestors=None, tap=None, jitter=1e-3):
    """Rao-Blackwellized particle filter (src/particleFilter.m:1-3).

    ``model`` supplies the dynModel/measModel closures; ``streams`` the injected
    uniforms/normals (sweep index ``sweep``).  ``forced_ancestors`` [T, N]
    (0-based, row 0 unused) overrides the draws (teacher forcing).  ``tap`` is
    called as tap(t, dict(...)) after each step with copies of the state.
    Returns the reference's 8 outputs in order.
    """
    if sparseFeatures is None:
        sparseFeatures = model.sparse
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y.reshape(-1, 1)
    odometry = np.asarray(odometry, dtype=np.float64)
    R = np.atleast_2d(np.asarray(R, dtype=np.float64))
    x0_nonLin = np.asarray(x0_nonLin, dtype=np.float64).reshape(-1)
    x0_lin = np.asarray(x0_lin, dtype=np.float64)
    if x0_lin.ndim == 1:
        x0_lin = x0_lin.reshape(-1, 1)
    P0_lin = np.asarray(P0_lin, dtype=np.float64)

    # :55-67 initial weights, states, covariances
    w = 1.0 / N_P * np.ones(N_P)
    logw = np.log(w)
    xn = np.repeat(x0_nonLin[:, None], N_P, axis=1)
    xl = x0_lin.copy() if x0_lin.shape[1] > 1 else np.repeat(x0_lin, N_P, axis=1)
    P = np.repeat(P0_lin[None, :, :], N_P, axis=0)

    nNonLin = x0_nonLin.shape[0]
    N_T = y.shape[0]
    Q, dt = _expand_Q_dt(Q, dt, N_T)

    # :92-97
    traj_max = np.full((nNonLin, N_T), np.nan)
    traj_mean = np.full((nNonLin, N_T), np.nan)
    yhattraj = np.full((y.shape[1], N_T), np.nan)
    xn_traj = np.zeros((nNonLin, N_P, N_T))
    xn_traj[:, :, 0] = xn
    ai = np.zeros(N_P, dtype=np.int64)
    iw_max = 0

    for t in range(N_T):                                          # :100
        xn_ = xn.copy()
        if t != 0:                                                # :103
            for i in range(N_P):                                  # :104-109
                if forced_ancestors is not None:
                    ai[i] = forced_ancestors[t][i]
                else:
                    ai[i] = sample(w, streams.U[sweep, t, i])
                xn[:, i] = model.dynModel(xn_[:, ai[i]], odometry[t - 1, :], dt[t - 1],
                                          Q[:, :, t - 1], streams.Z[sweep, t, i])
            xl = xl[:, ai]                                        # :112
            P = P[ai]                                             # :113
            xn_traj[:, :, t] = xn                                 # :117
            xn_traj[:, :, :t] = xn_traj[:, ai, :t]                # :118

        yt = y[t, :]                                              # :122
        if not sparseFeatures:
            dy = model.measModel(xn)                              # :124
        for i in range(N_P):                                      # :126-151
            if sparseFeatures:
                yhat, dyi = model.measModel_sparse(xn[:, i], xl[:, i])
                e, SS, ind = innovation(yt, dyi, xl[:, i], P[i], R, yhat)
            else:
                e, SS, ind = innovation(yt, dy[i], xl[:, i], P[i], R)
            logw[i], _ = log_weight(e, SS, jitter)

        w = normalise(logw)                                       # :153-156

        iw_max = int(np.argmax(w))                                # :159 first index on ties
        traj_max[:, t] = xn[:, iw_max]                            # :160
        traj_mean[:, t] = np.sum(xn * w, axis=1)                  # :161

        for i in range(N_P):                                      # :164-204
            if sparseFeatures:
                yhat, dyi = model.measModel_sparse(xn[:, i], xl[:, i])
                e, SS, ind = innovation(yt, dyi, xl[:, i], P[i], R, yhat)
                yhat_full = yhat
            else:
                dyi = dy[i]
                yhat_full = dyi @ xl[:, i]
                e, SS, ind = innovation(yt, dyi, xl[:, i], P[i], R)
            cS, _ = chol_jitter(SS, jitter)
            K = kalman_gain(P[i], dyi[ind, :], cS)
            xl[:, i] = xl[:, i] + K @ e                           # :197
            P[i] = P[i] - K @ SS @ K.T                            # :198
            if i == iw_max:
                yhattraj[:, t] = yhat_full                        # :201-203

        if tap is not None:
            tap(t, dict(xn=xn.copy(), xl=xl.copy(), P=P, logw=logw.copy(), w=w.copy(),
                        ai=ai.copy(), iw_max=iw_max))
        if makePlots is not None:                                 # :215-217
            makePlots(xn, xl[:, iw_max], P[iw_max], traj_max, yhattraj, xn_traj, traj_mean, xl, P)

    # :220-233 final extraction
    xl_max = xl[:, iw_max].copy()
    P_max = P[iw_max].copy()
    xl_mean = np.sum(xl * w, axis=1)
    P_mean = np.zeros((xl_mean.shape[0], xl_mean.shape[0]))
    for i in range(N_P):
        # quirk Q1: assigned, not accumulated (src/particleFilter.m:228-230)
        dxl = xl_mean - xl[:, i]
        P_mean = w[i] * (P[i] + np.outer(dxl, dxl))
    traj_sample_iwmax = xn_traj[:, iw_max, :].copy()
    return traj_max, traj_mean, xl_max, xl_mean, P_max, P_mean, traj_sample_iwmax, xn_traj
