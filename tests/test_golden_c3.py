"""The C3 configuration (examples/slam-sparse-visual) on the reference's own input fixture.

tests/golden/curve_x2.npz holds the arrays of examples/slam-sparse-visual/curve-x2.mat
(made by tests/golden/make_curve_x2.py); it is the only fixture the reference ships - there
are no golden OUTPUTS, so parity stays oracle-against-CUDA on these real inputs."""
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT, assert_close_norm

FIX = os.path.join(ROOT, "tests", "golden", "curve_x2.npz")


def _fixture():
    return dict(np.load(FIX))


def _args(pr):
    return (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])


def test_fixture_has_the_surveyed_shape():
    """20 landmarks, T = 197, 0..11 observed per step (mean 4.06), 8 steps with none."""
    fx = _fixture()
    assert fx["Yclean"].shape == (20, 197) and fx["map"].shape == (2, 20) and fx["p"].shape == (2, 197)
    obs = np.sum(~np.isnan(fx["Yclean"]), axis=0)
    assert obs.min() == 0 and obs.max() == 11 and np.sum(obs == 0) == 8
    assert abs(obs.mean() - 4.06) < 0.01


def test_oracle_projection_reproduces_the_fixture_observations():
    """measurement.m:36-50 restated in the oracle, evaluated at the fixture's true poses and
    landmarks, must reproduce the fixture's noise-free observations wherever one is recorded:
    a known-answer check of the oracle's measurement model against reference DATA."""
    fx = _fixture()
    om = oracle.SparseVisual2D(20, 1.5, 0.0, 1.0)
    xl = fx["map"].T.reshape(-1)
    worst = 0.0
    for t in range(197):
        xn = np.array([fx["p"][0, t], fx["p"][1, t], fx["th"][t, 0]])
        yhat, _ = om.measModel_sparse(xn, xl)
        seen = ~np.isnan(fx["Yclean"][:, t])
        if seen.any():
            worst = max(worst, np.max(np.abs(yhat[seen] - fx["Yclean"][seen, t])))
    assert worst < 1e-12, worst


def test_odometry_recipe_reproduces_the_fixture_increments():
    """load_data.m:73-78 (dPos = diff(p,[],2), dTheta = diff(unwrap(th))) as restated in
    rbslam/synth.py, with the noise switched off, against the increments stored in the fixture."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200"))
    from rbslam import synth
    fx = _fixture()
    pr = synth.sparse_visual_problem(fixture=fx, pos_var=0.0, angle_var=0.0, pos_bias=0.0, noise_var=0.0)
    assert_close_norm(pr["odometry"][:196, :2], fx["dPos"].T, 1e-14, "dPos")
    assert_close_norm(pr["odometry"][:196, 2], fx["dTheta"].reshape(-1), 1e-14, "dTheta")
    assert np.array_equal(np.isnan(pr["y"]), np.isnan(fx["Yclean"].T))


def test_oracle_filter_on_fixture_tracks_the_path():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200"))
    from rbslam import synth     # pure NumPy generators: no GPU needed
    N = 30
    pr = synth.c3_problem(_fixture(), N_P=N)
    om = oracle.SparseVisual2D(pr["n_landmarks"], *pr["camera"])
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(2), 1, T, N, om.nz)
    out = oracle.particleFilter(om, *_args(pr), N, pr["dt"], st)
    traj_mean = out[1]
    assert np.all(np.isfinite(traj_mean))
    err = np.sqrt(np.mean(np.sum((traj_mean[:2] - pr["truth"]["p"]) ** 2, axis=0)))
    dead_reckoning = pr["x0_nonLin"][:2, None] + np.cumsum(pr["odometry"][:T - 1, :2], axis=0).T
    err_dr = np.sqrt(np.mean(np.sum((dead_reckoning - pr["truth"]["p"][:, 1:]) ** 2, axis=0)))
    assert err < err_dr      # SLAM beats integrating the drifting odometry (load_data.m:82)


@pytest.mark.gpu
def test_c3_filter_on_fixture(rbslam_lib):
    """pfslam.m configuration: N_P = 100, M = 40, d = 20 with NaNs, T = 197."""
    rb = rbslam_lib
    N = 100
    pr = rb.synth.c3_problem(_fixture(), N_P=N)
    gm = rb.models.from_problem(pr)
    om = oracle.SparseVisual2D(pr["n_landmarks"], *pr["camera"])
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(3), 1, T, N, om.nz)
    taps = []
    ref = oracle.particleFilter(om, *_args(pr), N, pr["dt"], st,
                                tap=lambda t, d: taps.append(dict(logw=d["logw"], ai=d["ai"])))
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    for t in range(1, T):
        assert np.array_equal(o["ancestors"][:, t], taps[t]["ai"]), "ancestors t=%d" % t
    for t in range(T):
        assert_close_norm(o["logw_hist"][:, t], taps[t]["logw"], 1e-8, "logw t=%d" % t)
    names = ["traj_max", "traj_mean", "xl_max", "xl_mean", "P_max", "P_mean", "traj_sample_iwmax", "xn_traj"]
    for k, r in zip(names, ref):
        assert_close_norm(o[k], r, 1e-8 if k != "traj_mean" else 1e-7, k)


@pytest.mark.gpu
def test_c3_smoother_on_fixture(rbslam_lib):
    """psslam.m configuration (N_P = 10, dynResNorm = []), 3 of its 10 sweeps over the first 100
    steps of the fixture (the oracle's stacked future system grows as O(T^4) per sweep)."""
    rb = rbslam_lib
    N, K, Tcut = 10, 3, 100
    pr = rb.synth.c3_problem(_fixture(), N_P=N)
    pr["y"] = pr["y"][:Tcut].copy()
    pr["odometry"] = pr["odometry"][:Tcut].copy()
    gm = rb.models.from_problem(pr)
    om = oracle.SparseVisual2D(pr["n_landmarks"], *pr["camera"])
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(4), K, T, N, om.nz)
    rec = {}
    XNK, XLK, PK = oracle.particleSmoother(om, *_args(pr), N, K, pr["dt"], st, record=rec)
    ai = np.zeros((K, T, N), dtype=np.int32)
    for (k, t), v in rec["ai"].items():
        ai[k, t] = v
    ak = np.array([rec["ak"][k] for k in range(K)], dtype=np.int32)
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.smoother_run(*_args(pr), pr["dt"], K, 0, streams=st, forced_ancestors=ai, forced_ak=ak)
    assert_close_norm(o["XNK"], XNK, 1e-8, "XNK")
    assert_close_norm(o["XLK"], XLK, 1e-8, "XLK")
    assert_close_norm(o["PK"], PK, 1e-8, "PK")
