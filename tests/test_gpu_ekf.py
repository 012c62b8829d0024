"""The EKF baseline of the dense magnetic-field example (examples/slam-dense-mag/ekf_dense.m:37-102,
run_dense3D_magfield.m:281-316) on the device (rbslam_ekf_run) against its oracle restatement."""
import numpy as np
import pytest

import oracle
from conftest import assert_close_norm

pytestmark = pytest.mark.gpu


def _ekf_inputs(pr):
    M = pr["P0_lin"].shape[0]
    x0 = np.concatenate([pr["x0_nonLin"][:3], np.zeros(3), pr["x0_lin"].reshape(-1)])   # run_dense3D_magfield.m:248
    q0 = pr["x0_nonLin"][3:7]                                                           # :249
    P0 = np.zeros((M + 6, M + 6))
    P0[6:, 6:] = pr["P0_lin"]                                                           # :250
    return x0, q0, P0


@pytest.mark.parametrize("m,T", [(64, 40), (253, 30), (512, 12)])
def test_ekf_dense_matches_oracle(rbslam_lib, m, T):
    rb = rbslam_lib
    pr = rb.synth.dense_mag_problem(N_T=T, m=m, seed=3, m_sim=300)
    gm = rb.models.from_problem(pr)
    x0, q0, P0 = _ekf_inputs(pr)
    LL = np.vstack([-pr["L"], pr["L"]])
    LL[1, 0] += 0.7            # measModel_ekf hands the data's bounds to JacobianPhi3D, not [-L, L]
    ref = oracle.ekf_dense(pr["NN"], pr["L"], LL, pr["odometry"], pr["y"], x0, q0, P0, pr["Q"], pr["R"], pr["dt"],
                           keep_P=True)
    xf, qn, Pt = rb.ekf_dense(gm, pr["odometry"], pr["y"], x0, q0, P0, pr["Q"], pr["R"], pr["dt"], LL, keep_P=True)
    assert_close_norm(xf, ref[0], 1e-8, "xf_traj")
    assert_close_norm(qn, ref[1], 1e-10, "qnb_traj")
    assert_close_norm(Pt, ref[2], 1e-8, "Pf_traj")
    assert np.array_equal(Pt[:, :, -1], Pt[:, :, -1].T)          # ekf_dense.m:91 symmetrises
    xf2, qn2, Pl = rb.ekf_dense(gm, pr["odometry"], pr["y"], x0, q0, P0, pr["Q"], pr["R"], pr["dt"], LL)
    assert np.array_equal(Pl, Pt[:, :, -1]) and np.array_equal(xf2, xf)


def test_ekf_dense_time_varying_q_and_errors(rbslam_lib):
    rb = rbslam_lib
    T = 10
    pr = rb.synth.dense_mag_problem(N_T=T, m=64, seed=5, m_sim=300)
    gm = rb.models.from_problem(pr)
    x0, q0, P0 = _ekf_inputs(pr)
    rng = np.random.default_rng(0)
    Qt = np.repeat(pr["Q"][:, :, None], T - 1, axis=2) * (1 + rng.random(T - 1))
    dtv = pr["dt"] * (1 + 0.1 * rng.random(T - 1))
    LL = np.vstack([-pr["L"], pr["L"]])
    ref = oracle.ekf_dense(pr["NN"], pr["L"], LL, pr["odometry"], pr["y"], x0, q0, P0, Qt, pr["R"], dtv)
    xf, qn, Pl = rb.ekf_dense(gm, pr["odometry"], pr["y"], x0, q0, P0, Qt, pr["R"], dtv)
    assert_close_norm(xf, ref[0], 1e-8, "xf_traj")
    assert_close_norm(Pl, ref[2], 1e-8, "Pf")
    with pytest.raises(ValueError):
        rb.ekf_dense(gm, pr["odometry"], pr["y"], x0[:-1], q0, P0, pr["Q"], pr["R"], pr["dt"])
    radio = rb.models.from_problem(rb.synth.dense_radio_problem("line_3D", m=32, seed=2, m_sim=100))
    with pytest.raises(rb.UnsupportedModelError):
        rb.ekf_dense(radio, pr["odometry"], pr["y"], x0, q0, P0, pr["Q"], pr["R"], pr["dt"])
