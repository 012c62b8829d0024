"""Oracle restatement of src/particleSmoother.m (test infrastructure only).

Conditional particle filter with ancestor sampling (CPF-AS), N_K sweeps.
Follows src/particleSmoother.m:48-366; comments give the reference lines.
"""
import numpy as np
from .tools import sample, chol_jitter, solve_lower
from .particle_filter import (_expand_Q_dt, innovation, log_weight, kalman_gain,
                              normalise, LOG2PI)


def default_dyn_res_norm(xnkt, xni, odo, dt, Q):
    """Default transition residual when dynResNorm is [] (src/particleSmoother.m:175-177)."""
    Lc = np.linalg.cholesky(dt * Q)
    return np.linalg.solve(Lc.T, xnkt - xni - odo)


def stacked_future_jacobian(dy_xnk, t, ny):
    """D = future Jacobians stacked time-major, measurement index inner (:162-167).

    dy_xnk is [N_T x ny x nLin]; permute([2 1 3]) + reshape gives rows ordered
    (t, ny-block), (t+1, ny-block), ...
    """
    blk = dy_xnk[t:, :, :]
    return blk.reshape(blk.shape[0] * ny, blk.shape[2])


def particleSmoother(model, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, N_K, dt,
                     streams, sparseFeatures=None, makePlots=None, forced=None, tap=None, record=None,
                     jitter=1e-2, verbose=False):
    """Rao-Blackwellized particle smoother (src/particleSmoother.m:1-2).

    ``forced`` (optional) is a dict with 'ai' [K,T,N] and 'ak' [K] (0-based) to
    teacher-force every ancestor draw.  ``tap(k, t, dict)`` receives per-step
    state copies (paNt of the reference particle is included for k>=1, t>=1).
    Returns XNK [n x T x K], XLK [M x K], PK [M x M x K].
    """
    if sparseFeatures is None:
        sparseFeatures = model.sparse
    dynResNorm = getattr(model, "dynResNorm", None)
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

    nNonLin = x0_nonLin.shape[0]                                  # :50-53
    nLin = x0_lin.shape[0]
    N_T = y.shape[0]
    ny = y.shape[1]
    Q, dt = _expand_Q_dt(Q, dt, N_T)

    xn_traj = np.zeros((nNonLin, N_P, N_T))                       # :73
    ai = np.zeros(N_P, dtype=np.int64)
    XNK = np.full((nNonLin, N_T, N_K), np.nan)                    # :80-82
    XLK = np.full((nLin, N_K), np.nan)
    PK = np.full((nLin, nLin, N_K), np.nan)
    xnk = None

    for k in range(N_K):                                          # :88
        xn = np.repeat(x0_nonLin[:, None], N_P, axis=1)           # :91
        if k != 0:
            xn[:, N_P - 1] = xnk[:, 0]                            # :95
        xl = x0_lin.copy() if x0_lin.shape[1] > 1 else np.repeat(x0_lin, N_P, axis=1)
        P = np.repeat(P0_lin[None, :, :], N_P, axis=0)            # :104
        w = 1.0 / N_P * np.ones(N_P)                              # :107-108
        logw = np.log(w)
        if k != 0:
            xn_traj[:, N_P - 1, :] = xnk                          # :111-113
        xn_traj[:, :, 0] = xn                                     # :116
        if k != 0 and not sparseFeatures:
            dy_xnk = model.measModel(xnk)                         # :119-121  [N_T x ny x nLin]

        for t in range(N_T):                                      # :124
            paNt = None
            if t != 0:                                            # :127
                xn_pred = np.zeros_like(xn)
                xl_pred = np.zeros_like(xl)
                P_pred = np.zeros_like(P)
                for i in range(N_P - 1):                          # :132-137
                    if forced is not None:
                        ai[i] = forced["ai"][k, t, i]
                    else:
                        ai[i] = sample(w, streams.U[k, t, i])
                    xn_pred[:, i] = model.dynModel(xn[:, ai[i]], odometry[t - 1, :], dt[t - 1],
                                                   Q[:, :, t - 1], streams.Z[k, t, i])
                xl_pred[:, :-1] = xl[:, ai[:-1]]                  # :140
                P_pred[:-1] = P[ai[:-1]]                          # :141

                if k == 0:                                        # :145-155
                    i = N_P - 1
                    if forced is not None:
                        ai[i] = forced["ai"][k, t, i]
                    else:
                        ai[i] = sample(w, streams.U[k, t, i])
                    xn_pred[:, i] = model.dynModel(xn[:, ai[i]], odometry[t - 1, :], dt[t - 1],
                                                   Q[:, :, t - 1], streams.Z[k, t, i])
                    xl_pred[:, i] = xl[:, ai[i]]
                    P_pred[i] = P[ai[i]]
                else:                                             # :156-245
                    paNtLog = np.zeros(N_P)
                    if not sparseFeatures:
                        D = stacked_future_jacobian(dy_xnk, t, ny)        # :162-167
                        yfut = y[t:, :].reshape(-1)                       # :192
                        RR = np.kron(np.eye(N_T - t), R)                  # :191
                    xnkt = xnk[:, t]                                      # :170
                    for i in range(N_P):                                  # :171
                        if dynResNorm is None:                            # :175-180
                            eDyn = default_dyn_res_norm(xnkt, xn[:, i], odometry[t - 1, :],
                                                        dt[t - 1], Q[:, :, t - 1])
                        else:
                            eDyn = dynResNorm(xnkt, xn[:, i], odometry[t - 1, :], dt[t - 1],
                                              Q[:, :, t - 1])
                        logwDyn = -0.5 * (eDyn @ eDyn)                    # :182
                        if not sparseFeatures:
                            SS = D @ P[i] @ D.T + RR                      # :191
                            e = yfut - D @ xl[:, i]                       # :193
                        else:                                             # :197-216
                            es, dys, inds = [], [], []
                            for ti in range(t, N_T):
                                yt_ = y[ti, :]
                                ind = ~np.isnan(yt_)
                                yhat, dyi = model.measModel_sparse(xnk[:, ti], xl[:, i])
                                ei = yt_ - yhat
                                es.append(ei[ind])
                                dys.append(dyi[ind, :])
                                inds.append(np.flatnonzero(ind))
                            e = np.concatenate(es) if es else np.zeros(0)
                            Dsp = np.vstack(dys) if dys else np.zeros((0, nLin))
                            RS = np.zeros((e.size, e.size))
                            o = 0
                            for idx in inds:                              # blkdiag(RS, R(ind,ind))
                                RS[o:o + idx.size, o:o + idx.size] = R[np.ix_(idx, idx)]
                                o += idx.size
                            SS = Dsp @ P[i] @ Dsp.T + RS                  # :214
                        cS, _ = chol_jitter(SS, jitter)                   # :221-224
                        v = solve_lower(cS, e)                            # :225
                        logwMeas = (-np.sum(np.log(np.diag(cS))) - 0.5 * (v @ v)
                                    - e.shape[0] / 2 * LOG2PI)            # :229
                        paNtLog[i] = np.log(w[i]) + logwDyn + logwMeas    # :232
                    c = np.max(paNtLog)                                   # :236-238
                    lse = c + np.log(np.sum(np.exp(paNtLog - c)))
                    paNt = np.exp(paNtLog - lse)
                    i = N_P - 1
                    if forced is not None:
                        ai[i] = forced["ai"][k, t, i]
                    else:
                        ai[i] = sample(paNt, streams.U[k, t, i])          # :241
                    xn_pred[:, i] = xnkt
                    xl_pred[:, i] = xl[:, ai[i]]
                    P_pred[i] = P[ai[i]]

                xn, xl, P = xn_pred, xl_pred, P_pred                      # :251-253
                xn_traj[:, :, t] = xn                                     # :256
                xn_traj[:, :, :t] = xn_traj[:, ai, :t]                    # :257

            yt = y[t, :]                                                  # :262
            if not sparseFeatures:
                dy = model.measModel(xn)                                  # :264
            for i in range(N_P):                                          # :266-294
                if sparseFeatures:
                    yhat, dyi = model.measModel_sparse(xn[:, i], xl[:, i])
                    e, SS, ind = innovation(yt, dyi, xl[:, i], P[i], R, yhat)
                else:
                    e, SS, ind = innovation(yt, dy[i], xl[:, i], P[i], R)
                logw[i], _ = log_weight(e, SS, jitter)
            w = normalise(logw)                                           # :300-302

            for i in range(N_P):                                          # :305-340
                if sparseFeatures:
                    yhat, dyi = model.measModel_sparse(xn[:, i], xl[:, i])
                    e, SS, ind = innovation(yt, dyi, xl[:, i], P[i], R, yhat)
                else:
                    dyi = dy[i]
                    e, SS, ind = innovation(yt, dyi, xl[:, i], P[i], R)
                cS, _ = chol_jitter(SS, jitter)
                K = kalman_gain(P[i], dyi[ind, :], cS)
                xl[:, i] = xl[:, i] + K @ e
                P[i] = P[i] - K @ SS @ K.T

            if record is not None:
                record.setdefault("ai", {})[(k, t)] = ai.copy()
                record.setdefault("paNt", {})[(k, t)] = None if paNt is None else paNt.copy()
            if tap is not None:
                tap(k, t, dict(xn=xn.copy(), xl=xl.copy(), P=P, logw=logw.copy(), w=w.copy(),
                               ai=ai.copy(), paNt=None if paNt is None else paNt.copy()))

        if forced is not None:                                            # :346
            ak = int(forced["ak"][k])
        else:
            ak = sample(w, streams.Uend[k])
        if record is not None:
            record.setdefault("ak", {})[k] = ak
        xnk = xn_traj[:, ak, :].copy()                                    # :347
        XNK[:, :, k] = xnk                                                # :352-354
        XLK[:, k] = xl[:, ak]
        PK[:, :, k] = P[ak]
        if makePlots is not None:
            makePlots(xnk, xl[:, ak], k, XNK, XLK, PK)
        if verbose:
            print("Particle smoother iteration %i/%i done." % (k + 1, N_K))   # :365
    return XNK, XLK, PK
