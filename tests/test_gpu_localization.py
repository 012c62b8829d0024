"""The localisation-only particle filter (examples/mag-localization-mapping/particleFilterLocalization.m:50-132,
closures of run_localization.m:241-280) on the device against its oracle restatement."""
import numpy as np
import pytest

import oracle
from conftest import assert_close_norm

pytestmark = pytest.mark.gpu


def _setup(rb, m, T, N, seed):
    pr = rb.synth.dense_mag_problem(N_T=T, m=m, seed=seed, m_sim=300)
    gm = rb.models.from_problem(pr)
    rng = np.random.default_rng(seed)
    M = m + 3
    foo = np.sqrt(np.diag(pr["P0_lin"])) * rng.standard_normal(M)      # a map drawn from the GP prior
    dVar = 0.5 + rng.random((N, 3))                                    # predictive variances, one row per particle
    sigma2 = float(pr["R"][0, 0])
    om = oracle.DenseMag3D(pr["NN"], pr["L"])
    pos = pr["truth"]["pos"].T
    y = np.zeros((T, 3))
    for t in range(T):                                                 # measurements of that map along the true path
        xt = np.concatenate([pos[t], pr["x0_nonLin"][3:7]])[:, None]
        y[t] = om.measModel(xt)[0] @ foo + 0.05 * rng.standard_normal(3)
    return pr, gm, foo, dVar, sigma2, y


@pytest.mark.parametrize("m,T,N", [(64, 20, 50), (253, 12, 200), (1000, 6, 1000)])   # last: the example's N_P = 1000, M = 1003
def test_localization_filter_matches_oracle(rbslam_lib, m, T, N):
    rb = rbslam_lib
    pr, gm, foo, dVar, sigma2, y = _setup(rb, m, T, N, seed=5)
    rng = np.random.default_rng(1)

    class S:
        U = rng.random((T, N))
        Z = rng.standard_normal((T, N, 6))
    taps = []
    ref = oracle.particleFilterLocalization(pr["NN"], pr["L"], foo, dVar, sigma2, pr["odometry"], y, pr["x0_nonLin"],
                                            pr["Q"], N, pr["dt"], S.U, S.Z, tap=lambda t, d: taps.append(d))
    tm, tmean, ex = rb.particleFilterLocalization(gm.dynModel, gm.measModel, pr["odometry"], y, pr["x0_nonLin"], pr["Q"],
                                                  pr["R"], N, pr["dt"], None, map_mean=foo, var_rows=dVar,
                                                  sigma2=sigma2, rng=S, want_xn_traj=True, taps=True)
    for t in range(T):
        assert_close_norm(ex["w_hist"][:, t], taps[t]["w"], 1e-8, "w t=%d" % t)
        if t:
            assert np.array_equal(ex["ancestors"][:, t], taps[t]["ai"]), "ancestors t=%d" % t
    assert_close_norm(tm, ref[0], 1e-10, "traj_max")
    assert_close_norm(tmean, ref[1], 1e-8, "traj_mean")
    assert_close_norm(ex["xn_traj"], ref[2], 1e-10, "xn_traj")
    assert ex["n_diverged"] == 0


def test_localization_philox_and_argument_errors(rbslam_lib):
    rb = rbslam_lib
    pr, gm, foo, dVar, sigma2, y = _setup(rb, 64, 8, 40, seed=7)
    a = rb.particleFilterLocalization(gm.dynModel, gm.measModel, pr["odometry"], y, pr["x0_nonLin"], pr["Q"], pr["R"], 40,
                                      pr["dt"], map_mean=foo, var_rows=dVar, sigma2=sigma2, rng=3)
    b = rb.particleFilterLocalization(gm.dynModel, gm.measModel, pr["odometry"], y, pr["x0_nonLin"], pr["Q"], pr["R"], 40,
                                      pr["dt"], map_mean=foo, var_rows=dVar, sigma2=sigma2, rng=3)
    assert np.array_equal(a[0], b[0]) and np.all(np.isfinite(a[1]))        # device stream: reproducible
    with pytest.raises(ValueError):
        rb.particleFilterLocalization(gm.dynModel, gm.measModel, pr["odometry"], y, pr["x0_nonLin"], pr["Q"], pr["R"], 40,
                                      pr["dt"], map_mean=foo[:-1], var_rows=dVar, sigma2=sigma2, rng=3)
