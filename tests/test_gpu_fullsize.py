"""Parity at BASELINE.json's full C4 size (N = 10^4 particles, m = 1024 -> M = 1027, 84.8 GB
of covariance slabs), where the oracle cannot run the whole population:
  * particles evolve independently once ancestors and noise are fixed, so the oracle runs a
    SAMPLE of the particles (teacher-forced identity ancestors, injected normals) and must
    match the full-size CUDA run on exactly those particles, every step;
  * size-independent properties of a free-running full-size run: ancestors bit-exact against
    a NumPy recomputation from the run's own weights and the Philox uniforms, weights
    normalised, everything finite."""
import numpy as np
import pytest

import oracle
from conftest import assert_close_norm

pytestmark = pytest.mark.gpu

N, M_BASIS, T = 10000, 1024, 4


def _problem(rb):
    pr = rb.synth.dense_mag_problem(N_T=2000, m=M_BASIS, seed=1, n_laps=10, m_sim=2000)
    pr["y"] = pr["y"][:T].copy()
    pr["odometry"] = pr["odometry"][:T].copy()
    return pr


def _args(pr):
    return (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])


def _need_big_gpu():
    import subprocess
    try:
        mib = int(subprocess.check_output(["nvidia-smi", "--query-gpu=memory.total", "--format=csv,noheader,nounits"],
                                          text=True).splitlines()[0])
    except Exception:
        pytest.skip("nvidia-smi unavailable")
    if mib < 120000:
        pytest.skip("needs a >=120 GB GPU for the full C4 state")


def test_c4_fullsize_sampled_particles_match_oracle(rbslam_lib):
    _need_big_gpu()
    rb = rbslam_lib
    pr = _problem(rb)
    gm = rb.models.from_problem(pr)
    om = oracle.DenseMag3D(pr["NN"], pr["L"])
    rng = np.random.default_rng(3)

    class S:
        U = np.zeros((1, T, N))
        Z = rng.standard_normal((1, T, N, 6))
    forced = np.tile(np.arange(N, dtype=np.int32), (1, T, 1))        # identity ancestors
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=S, forced_ancestors=forced, want_xn_traj=False,
                           taps=True)
    w_last = o["w_hist"][:, T - 1]
    iw = int(np.argmax(w_last))
    sample = sorted({0, 1, 4999, iw, N - 1})
    st = oracle.Streams(np.zeros((1, T, len(sample))), S.Z[:, :, sample, :])
    taps = []
    oracle.particleFilter(om, *_args(pr), len(sample), pr["dt"], st,
                          forced_ancestors=np.tile(np.arange(len(sample)), (T, 1)),
                          tap=lambda t, d: taps.append(dict(logw=d["logw"].copy(), xl=d["xl"].copy(),
                                                            P=np.array(d["P"]), xn=d["xn"].copy())))
    for t in range(T):      # every step's (unnormalised) log-weight of the sampled particles
        ref = taps[t]["logw"]
        got = o["logw_hist"][sample, t]
        assert np.all(np.abs(got - ref) <= 1e-8 * np.maximum(1.0, np.abs(ref))), (t, got, ref)
    k = sample.index(iw)
    assert_close_norm(o["xl_max"], taps[-1]["xl"][:, k], 1e-8, "xl_max")
    assert_close_norm(o["P_max"], taps[-1]["P"][k], 1e-8, "P_max")
    assert_close_norm(o["traj_max"][:, -1], taps[-1]["xn"][:, k], 1e-12, "traj_max")
    kl = sample.index(N - 1)     # quirk Q1: P_mean = w(N)*(P_N + dx dx')
    dxl = o["xl_mean"] - taps[-1]["xl"][:, kl]
    assert_close_norm(o["P_mean"], w_last[N - 1] * (taps[-1]["P"][kl] + np.outer(dxl, dxl)), 1e-8, "P_mean")


def test_c4_fullsize_free_running_properties(rbslam_lib):
    _need_big_gpu()
    rb = rbslam_lib
    pr = _problem(rb)
    gm = rb.models.from_problem(pr)
    seed = 11
    with rb.Context(gm, N, T, rng_mode=1, seed=seed) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], want_xn_traj=False, taps=True)
    assert np.all(np.isfinite(o["logw_hist"])) and np.all(np.isfinite(o["P_max"]))
    for t in range(T):
        assert abs(o["w_hist"][:, t].sum() - 1.0) < 1e-10
    for t in range(1, T):    # bit-exact multinomial resampling at N = 10^4 (tools/sample.m:30-32)
        U, _ = oracle.philox_uniforms_normals(seed, 0, t, N, 0)
        ref = oracle.tools.sample_many(o["w_hist"][:, t - 1], U)
        assert np.array_equal(o["ancestors"][:, t], ref), t
    # P_max symmetric to rounding, and the Kalman update can only shrink the prior variances
    assert np.max(np.abs(o["P_max"] - o["P_max"].T)) <= 1e-9 * np.max(np.abs(o["P_max"]))
    assert np.all(np.diag(o["P_max"]) <= np.diag(pr["P0_lin"]) * (1 + 1e-12))
    assert np.all(np.diag(o["P_max"]) > 0)
