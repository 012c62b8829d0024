"""Oracle restatement of src/particleSmootherInformationForm.m (test infrastructure only).

Same CPF-AS recursion as particle_smoother.py but the ancestor weights of the
reference trajectory are evaluated in information form, which makes their cost
independent of the number of future time steps.  Follows
src/particleSmootherInformationForm.m:54-361.
"""
import numpy as np
from .tools import sample, chol_lower, chol_jitter, solve_lower
from .particle_filter import _expand_Q_dt, kalman_gain, normalise
from .particle_smoother import default_dyn_res_norm


def particleSmootherInformationForm(model, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P,
                                    N_K, dt, streams, sparseFeatures=None, makePlots=None,
                                    forced=None, tap=None, record=None, jitter=1e-2, verbose=False):
    """Information-form RBPS (src/particleSmootherInformationForm.m:1-2).

    Dense models only (:77-80).  Quirk Q8: a per-particle x0_lin is ignored
    (only column semantics of ``repmat(x0_lin,1,N_P)`` with an [M x 1] x0_lin are
    meaningful, :109).  Quirk Q7: if ``chol(ImatEnd)`` fails the reference
    re-factors the partial factor and then hits a size error (:229-231); the
    oracle raises LinAlgError in that case.
    """
    if sparseFeatures is None:
        sparseFeatures = model.sparse
    if sparseFeatures:
        print("This code has only been implemented for dense features")   # :77-80
        return None
    dynResNorm = getattr(model, "dynResNorm", None)
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y.reshape(-1, 1)
    odometry = np.asarray(odometry, dtype=np.float64)
    R = np.atleast_2d(np.asarray(R, dtype=np.float64))
    Rinv = np.linalg.inv(R)
    x0_nonLin = np.asarray(x0_nonLin, dtype=np.float64).reshape(-1)
    x0_lin = np.asarray(x0_lin, dtype=np.float64).reshape(-1, 1)
    P0_lin = np.asarray(P0_lin, dtype=np.float64)

    nNonLin = x0_nonLin.shape[0]
    nLin = x0_lin.shape[0]
    N_T = y.shape[0]
    ny = y.shape[1]
    Q, dt = _expand_Q_dt(Q, dt, N_T)
    logdetR = np.log(np.linalg.det(R))

    xn_traj = np.zeros((nNonLin, N_P, N_T))
    ai = np.zeros(N_P, dtype=np.int64)
    XNK = np.full((nNonLin, N_T, N_K), np.nan)
    XLK = np.full((nLin, N_K), np.nan)
    PK = np.full((nLin, nLin, N_K), np.nan)
    xnk = None

    def rank_terms(H, yrow):
        """(H'/R*y', H'/R*H) of one time step (:136-144, :193-201, :292, :333-334)."""
        HtRi = H.T @ Rinv
        return HtRi @ yrow, HtRi @ H

    for k in range(N_K):                                          # :98
        xn = np.repeat(x0_nonLin[:, None], N_P, axis=1)           # :101
        if k != 0:
            xn[:, N_P - 1] = xnk[:, 0]                            # :105
        xl = np.repeat(x0_lin, N_P, axis=1)                       # :109
        ivec0 = np.diag(1.0 / np.diag(P0_lin)) @ x0_lin[:, 0]     # :110
        ivec = np.repeat(ivec0[:, None], N_P, axis=1)             # :111
        P = np.repeat(P0_lin[None, :, :], N_P, axis=0)            # :112
        Imat = np.repeat(np.diag(1.0 / np.diag(P0_lin))[None, :, :], N_P, axis=0)   # :113
        halfLogDetP = np.sum(np.log(np.sqrt(np.diag(P0_lin)))) * np.ones(N_P)       # :115
        w = 1.0 / N_P * np.ones(N_P)                              # :118-119
        logw = np.log(w)
        if k != 0:
            xn_traj[:, N_P - 1, :] = xnk                          # :122-124
        xn_traj[:, :, 0] = xn                                     # :127

        if k != 0:                                                # :132-146
            dy_xnk = model.measModel(xnk)                         # [N_T x ny x nLin]
            ImatAddt = np.zeros((nLin, nLin))
            ivecAddt = np.zeros(nLin)
            for jj in range(N_T):
                dv, dM = rank_terms(dy_xnk[jj], y[jj, :])
                ivecAddt = ivecAddt + dv
                ImatAddt = ImatAddt + dM

        for t in range(N_T):                                      # :149
            paNt = None
            if t != 0:                                            # :152
                xn_pred = np.zeros_like(xn)
                xl_pred = np.zeros_like(xl)
                P_pred = np.zeros_like(P)
                ivec_pred = np.zeros_like(ivec)
                Imat_pred = np.zeros_like(Imat)
                for i in range(N_P - 1):                          # :159-164
                    if forced is not None:
                        ai[i] = forced["ai"][k, t, i]
                    else:
                        ai[i] = sample(w, streams.U[k, t, i])
                    xn_pred[:, i] = model.dynModel(xn[:, ai[i]], odometry[t - 1, :], dt[t - 1],
                                                   Q[:, :, t - 1], streams.Z[k, t, i])
                xl_pred[:, :-1] = xl[:, ai[:-1]]                  # :167-170
                P_pred[:-1] = P[ai[:-1]]
                ivec_pred[:, :-1] = ivec[:, ai[:-1]]
                Imat_pred[:-1] = Imat[ai[:-1]]

                if k == 0:                                        # :174-186
                    i = N_P - 1
                    if forced is not None:
                        ai[i] = forced["ai"][k, t, i]
                    else:
                        ai[i] = sample(w, streams.U[k, t, i])
                    xn_pred[:, i] = model.dynModel(xn[:, ai[i]], odometry[t - 1, :], dt[t - 1],
                                                   Q[:, :, t - 1], streams.Z[k, t, i])
                else:                                             # :187-254
                    paNtLog = np.zeros(N_P)
                    dv, dM = rank_terms(dy_xnk[t - 1], y[t - 1, :])       # :192-201
                    ivecAddt = ivecAddt - dv
                    ImatAddt = ImatAddt - dM
                    xnkt = xnk[:, t]                                      # :204
                    for i in range(N_P):                                  # :205
                        if dynResNorm is None:                            # :209-214
                            eDyn = default_dyn_res_norm(xnkt, xn[:, i], odometry[t - 1, :],
                                                        dt[t - 1], Q[:, :, t - 1])
                        else:
                            eDyn = dynResNorm(xnkt, xn[:, i], odometry[t - 1, :], dt[t - 1],
                                              Q[:, :, t - 1])
                        logwDyn = -0.5 * (eDyn @ eDyn)                    # :216
                        ivecEnd = ivec[:, i] + ivecAddt                   # :224
                        ImatEnd = Imat[i] + ImatAddt                      # :225
                        cIend, flag = chol_lower(ImatEnd)                 # :228
                        if flag > 0:                                      # :229-231 (quirk Q7)
                            raise np.linalg.LinAlgError(
                                "chol(ImatEnd) failed; the reference's jitter branch is broken")
                        vIend = solve_lower(cIend, ivecEnd)               # :233
                        logwMeas = (-0.5 * (ivec[:, i] @ P[i] @ ivec[:, i]) - halfLogDetP[i]
                                    - np.sum(np.log(np.diag(cIend))) + 0.5 * (vIend @ vIend))  # :234-236
                        paNtLog[i] = np.log(w[i]) + logwDyn + logwMeas    # :239
                    c = np.max(paNtLog)                                   # :243-245
                    lse = c + np.log(np.sum(np.exp(paNtLog - c)))
                    paNt = np.exp(paNtLog - lse)
                    i = N_P - 1
                    if forced is not None:
                        ai[i] = forced["ai"][k, t, i]
                    else:
                        ai[i] = sample(paNt, streams.U[k, t, i])          # :248
                    xn_pred[:, i] = xnkt
                i = N_P - 1
                xl_pred[:, i] = xl[:, ai[i]]                              # :183-186 / :250-253
                P_pred[i] = P[ai[i]]
                ivec_pred[:, i] = ivec[:, ai[i]]
                Imat_pred[i] = Imat[ai[i]]

                xn, xl, P, ivec, Imat = xn_pred, xl_pred, P_pred, ivec_pred, Imat_pred   # :260-264
                halfLogDetP = halfLogDetP[ai]                             # :267
                xn_traj[:, :, t] = xn                                     # :270-271
                xn_traj[:, :, :t] = xn_traj[:, ai, :t]

            yt = y[t, :]                                                  # :275
            dy = model.measModel(xn)                                      # :276
            halfLogDetPplus = np.zeros_like(halfLogDetP)
            ivecPlus = np.zeros_like(ivec)
            ytRy = yt @ Rinv @ yt
            for i in range(N_P):                                          # :279-305
                dyi = dy[i]
                SS = dyi @ P[i] @ dyi.T + R
                cS, _ = chol_jitter(SS, jitter)
                dv, _ = rank_terms(dyi, yt)
                ivecPlus[:, i] = ivec[:, i] + dv                          # :292
                K = kalman_gain(P[i], dyi, cS)                            # :293
                Pplus = P[i] - K @ SS @ K.T                               # :294
                halfLogDetPplus[i] = (-np.sum(np.log(np.diag(cS))) + 0.5 * logdetR
                                      + halfLogDetP[i])                   # :298
                logw[i] = (-0.5 * (ivec[:, i] @ P[i] @ ivec[:, i]) - halfLogDetP[i]
                           + halfLogDetPplus[i]
                           + 0.5 * (ivecPlus[:, i] @ Pplus @ ivecPlus[:, i]) - 0.5 * ytRy
                           - 0.5 * np.log((2 * np.pi) ** yt.size * np.linalg.det(R)))   # :301-304
            halfLogDetP = halfLogDetPplus                                 # :308
            w = normalise(logw)                                           # :311-313

            for i in range(N_P):                                          # :316-335
                dyi = dy[i]
                e = yt - dyi @ xl[:, i]
                SS = dyi @ P[i] @ dyi.T + R
                cS, _ = chol_jitter(SS, jitter)
                K = kalman_gain(P[i], dyi, cS)
                xl[:, i] = xl[:, i] + K @ e
                P[i] = P[i] - K @ SS @ K.T
                dv, dM = rank_terms(dyi, yt)
                ivec[:, i] = ivec[:, i] + dv                              # :333
                Imat[i] = Imat[i] + dM                                    # :334

            if record is not None:
                record.setdefault("ai", {})[(k, t)] = ai.copy()
                record.setdefault("paNt", {})[(k, t)] = None if paNt is None else paNt.copy()
            if tap is not None:
                tap(k, t, dict(xn=xn.copy(), xl=xl.copy(), P=P, ivec=ivec.copy(), Imat=Imat,
                               halfLogDetP=halfLogDetP.copy(), logw=logw.copy(), w=w.copy(),
                               ai=ai.copy(), paNt=None if paNt is None else paNt.copy()))

        if forced is not None:                                            # :341
            ak = int(forced["ak"][k])
        else:
            ak = sample(w, streams.Uend[k])
        if record is not None:
            record.setdefault("ak", {})[k] = ak
        xnk = xn_traj[:, ak, :].copy()
        XNK[:, :, k] = xnk
        XLK[:, k] = xl[:, ak]
        PK[:, :, k] = P[ak]
        if makePlots is not None:
            makePlots(xnk, xl[:, ak], k, XNK, XLK, PK)
        if verbose:
            print("Particle smoother iteration %i/%i done." % (k + 1, N_K))
    return XNK, XLK, PK
