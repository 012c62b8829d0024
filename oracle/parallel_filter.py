"""Multi-threaded driver of the oracle particle filter (test infrastructure only).

Used solely as the CPU baseline of bench.py (``cpu_baseline`` leg and
``--impl reference``).  The reference itself (MATLAB) cannot run in this image, so
the baseline is this NumPy/OpenBLAS restatement with the reference's loop
structure (src/particleFilter.m:100-204): the two per-particle loops (weights
:126-151, Kalman update :164-204) are spread over a thread pool, one BLAS thread
per worker (NumPy releases the GIL inside BLAS and ufunc loops), everything else
is as in oracle/particle_filter.py.
"""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .particle_filter import (_expand_Q_dt, innovation, log_weight, kalman_gain, normalise)
from .tools import sample_many, chol_jitter


def filter_steps_timed(model, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, dt, streams,
                       n_steps, warmup=1, cores=None, jitter=1e-3):
    """Run ``warmup + n_steps`` time steps of the dense filter on N_P particles.

    Returns (seconds for the n_steps timed steps, cores used).  Arithmetic per
    particle is identical to oracle.particleFilter.
    """
    try:
        from threadpoolctl import threadpool_limits
    except Exception:   # pragma: no cover
        threadpool_limits = None
    cores = cores or os.cpu_count() or 1
    y = np.asarray(y, dtype=np.float64)
    R = np.atleast_2d(np.asarray(R, dtype=np.float64))
    x0_nonLin = np.asarray(x0_nonLin, dtype=np.float64).reshape(-1)
    x0_lin = np.asarray(x0_lin, dtype=np.float64).reshape(-1, 1)
    N_T = warmup + n_steps
    Q, dt = _expand_Q_dt(Q, dt, y.shape[0])
    w = np.ones(N_P) / N_P
    logw = np.log(w)
    xn = np.repeat(x0_nonLin[:, None], N_P, axis=1)
    xl = np.repeat(x0_lin, N_P, axis=1)
    P = np.repeat(np.asarray(P0_lin, dtype=np.float64)[None], N_P, axis=0)
    chunks = np.array_split(np.arange(N_P), cores)

    def weights(idx, dy, yt):
        for i in idx:
            e, SS, _ = innovation(yt, dy[i], xl[:, i], P[i], R)
            logw[i], _ = log_weight(e, SS, jitter)

    def update(idx, dy, yt):
        for i in idx:
            e, SS, ind = innovation(yt, dy[i], xl[:, i], P[i], R)
            cS, _ = chol_jitter(SS, jitter)
            K = kalman_gain(P[i], dy[i], cS)
            xl[:, i] = xl[:, i] + K @ e
            P[i] = P[i] - K @ SS @ K.T

    ctxm = threadpool_limits(limits=1) if threadpool_limits else None
    t0 = None
    step_secs = []
    try:
        with ThreadPoolExecutor(max_workers=cores) as pool:
            for t in range(N_T):
                ts = time.perf_counter()
                if t == warmup:
                    t0 = ts
                if t != 0:
                    ai = sample_many(w, streams.U[0, t, :N_P])
                    xn_ = xn.copy()
                    for i in range(N_P):
                        xn[:, i] = model.dynModel(xn_[:, ai[i]], odometry[t - 1, :], dt[t - 1],
                                                  Q[:, :, t - 1], streams.Z[0, t, i])
                    xl = xl[:, ai]
                    P = P[ai]
                yt = y[t, :]
                dy = model.measModel(xn)
                list(pool.map(lambda c: weights(c, dy, yt), chunks))
                w = normalise(logw)
                list(pool.map(lambda c: update(c, dy, yt), chunks))
                if t >= warmup:
                    step_secs.append(time.perf_counter() - ts)
        secs = time.perf_counter() - t0
    finally:
        if ctxm is not None:
            ctxm.__exit__(None, None, None)
    filter_steps_timed.last_step_secs = step_secs   # per-step wall times of the timed steps (for a robust rate)
    return secs, cores
