"""End-to-end parity of the CUDA particle filter (rbslam_filter_run) against the
oracle restatement of src/particleFilter.m, on the reference's three example
configurations (C1 dense-mag, C2 dense-radio, C3 sparse-visual) at reduced and
full example sizes."""
import numpy as np
import pytest

import oracle
from conftest import assert_close_norm

pytestmark = pytest.mark.gpu

TOL = 1e-8   # north_star: means, covariances and log-weights within 1e-8 relative


def _setup(rb, fam, N, **kw):
    s = rb.synth
    if fam == "radio":
        pr = s.dense_radio_problem(kw.get("traj", "line_3D"), m=kw.get("m", 128), seed=2, m_sim=400)
        om = oracle.DenseRadio2D(pr["NN"], pr["L"])
    elif fam == "mag":
        pr = s.dense_mag_problem(N_T=kw.get("T", 24), m=kw.get("m", 64), seed=3, m_sim=300)
        om = oracle.DenseMag3D(pr["NN"], pr["L"])
    else:
        pr = s.sparse_visual_problem(N_T=kw.get("T", 60), n_landmarks=20, N_P=N, seed=4,
                                     guess_map_var=0.01)
        om = oracle.SparseVisual2D(pr["n_landmarks"], *pr["camera"])
    return pr, om, rb.models.from_problem(pr)


def _args(pr):
    return (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])


def _run_oracle(om, pr, N, st, forced=None):
    taps = []
    out = oracle.particleFilter(om, *_args(pr), N, pr["dt"], st, forced_ancestors=forced,
                                tap=lambda t, d: taps.append(dict(logw=d["logw"], w=d["w"],
                                                                  ai=d["ai"], xl=d["xl"])))
    return out, taps


def _compare(o, ref, taps, T, tol=TOL):
    names = ["traj_max", "traj_mean", "xl_max", "xl_mean", "P_max", "P_mean", "traj_sample_iwmax",
             "xn_traj"]
    lw = np.stack([tp["logw"] for tp in taps], axis=1)          # [N, T]
    for t in range(T):
        # log-weights may be hugely negative; compare shifted like the normalisation does
        assert_close_norm(o["logw_hist"][:, t], lw[:, t], tol, "logw t=%d" % t)
    ww = np.stack([tp["w"] for tp in taps], axis=1)
    assert_close_norm(o["w_hist"], ww, 1e-7, "w")
    for t in range(1, T):
        assert np.array_equal(o["ancestors"][:, t], taps[t]["ai"]), "ancestors t=%d" % t
    for k, r in zip(names, ref):
        assert_close_norm(o[k], r, tol if k != "traj_mean" else 1e-7, k)


@pytest.mark.parametrize("fam,N,kw", [
    ("radio", 100, {}),                                  # C2 full size: N=100, M=128, T=32
    ("radio", 64, {"traj": "square_3D", "m": 60}),
    ("mag", 24, {"m": 64, "T": 24}),                     # small-M kernel
    ("mag", 16, {"m": 253, "T": 12}),                    # streaming kernels (M=256)
    ("mag", 40, {"m": 1024, "T": 5}),                    # C4 slab size: k_stream_fam<3,3,8,2,3>, real families
    ("sparse", 40, {"T": 60}),                           # C3 shape: M=40, d=20 with NaNs
])
def test_filter_teacher_forced(rbslam_lib, fam, N, kw):
    """Ancestors taken from the oracle run; every step's logw, and all 8 outputs, must match."""
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(5), 1, T, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    forced = np.stack([tp["ai"] for tp in taps])[None].astype(np.int32)    # [1, T, N]
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, forced_ancestors=forced, taps=True)
    _compare(o, ref, taps, T)


@pytest.mark.parametrize("fam,N,kw", [("radio", 100, {}), ("mag", 24, {"m": 64, "T": 24}),
                                      ("sparse", 40, {"T": 60}),
                                      ("mag", 33000, {"m": 5, "T": 3})])      # chunked normalisation, split scan / search
def test_filter_free_running_injected(rbslam_lib, fam, N, kw):
    """Same injected uniforms/normals, device draws its own ancestors: bit-exact indices."""
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(6), 1, T, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    _compare(o, ref, taps, T)


def test_filter_philox_matches_oracle_philox(rbslam_lib):
    """Device Philox stream == oracle/streams.py restatement of it (free-running mode)."""
    rb = rbslam_lib
    N = 50
    pr, om, gm = _setup(rb, "radio", N)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_philox(1234, 1, T, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    with rb.Context(gm, N, T, rng_mode=1, seed=1234) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], taps=True)
    _compare(o, ref, taps, T, tol=1e-7)


def test_particlefilter_dropin_signature(rbslam_lib):
    """The host mirror keeps the reference's positional signature and 8 outputs."""
    rb = rbslam_lib
    N = 30
    pr, om, gm = _setup(rb, "radio", N)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(9), 1, T, N, om.nz)
    outs = rb.particleFilter(gm.dynModel, gm.measModel, pr["odometry"], pr["y"], pr["x0_nonLin"],
                             pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"], N, pr["dt"], False, None,
                             rng=st)
    ref, _ = _run_oracle(om, pr, N, st)
    assert len(outs) == 8
    for a, b in zip(outs, ref):
        assert_close_norm(a, b, 1e-7)
    with pytest.raises(rb.UnsupportedModelError):
        rb.particleFilter(lambda *a: 0, gm.measModel, pr["odometry"], pr["y"], pr["x0_nonLin"],
                          pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"], N, pr["dt"])


def test_filter_step_callback_and_read_particles(rbslam_lib):
    rb = rbslam_lib
    N = 20
    pr, om, gm = _setup(rb, "radio", N, m=40)
    T = 6
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(3), 1, T, N, om.nz)
    states = {}
    oracle.particleFilter(om, pr["odometry"][:T], pr["y"][:T], pr["x0_nonLin"], pr["x0_lin"],
                          pr["P0_lin"], pr["Q"], pr["R"], N, pr["dt"], st,
                          tap=lambda t, d: states.__setitem__(t, dict(xl=d["xl"].copy(),
                                                                      P=np.array(d["P"]),
                                                                      xn=d["xn"].copy())))
    seen = []
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        def cb(k, t):
            s = ctx.read_particles()
            assert_close_norm(s["xl"], states[t]["xl"], TOL, "xl@%d" % t)
            assert_close_norm(s["P"], states[t]["P"].transpose(1, 2, 0), TOL, "P@%d" % t)
            assert_close_norm(s["xn"], states[t]["xn"], 1e-12, "xn@%d" % t)
            seen.append(t)
        ctx.set_step_callback(cb)
        ctx.filter_run(pr["odometry"][:T], pr["y"][:T], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"],
                       pr["Q"], pr["R"], pr["dt"], streams=st)
    assert seen == list(range(T))


@pytest.mark.parametrize("N,m,T", [(100, 512, 6)])
def test_filter_c1_shape(rbslam_lib, N, m, T):
    """C1 dense-mag at the example's N=100, M=515 (few steps): streaming kernels."""
    rb = rbslam_lib
    pr, om, gm = _setup(rb, "mag", N, m=m, T=T)
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(1), 1, T, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    _compare(o, ref, taps, T)


@pytest.mark.parametrize("variant", [2, 7], ids=["full", "packed"])
def test_filter_heavy_families(rbslam_lib, variant):
    """Forced ancestors with one ancestor taking most offspring (families far above the 4 / 6
    siblings a main family holds -> surplus families, several batches, copies landing in many
    dead slabs) next to singletons and childless particles."""
    rb = rbslam_lib
    N = 32
    pr, om, gm = _setup(rb, "mag", N, m=253, T=6)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(8), 1, T, N, om.nz)
    rng = np.random.default_rng(4)
    forced = np.zeros((1, T, N), dtype=np.int32)
    for t in range(1, T):
        heavy = int(rng.integers(0, N))
        a = rng.integers(0, N, N)
        a[rng.permutation(N)[:19]] = heavy          # 19+ offspring of one ancestor
        a[rng.permutation(N)[:7]] = (heavy + 5) % N  # and a 7-sibling family
        forced[0, t] = a
    ref, taps = _run_oracle(om, pr, N, st, forced=forced[0])
    with rb.Context(gm, N, T, rng_mode=0, kalman_variant=variant) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, forced_ancestors=forced, taps=True)
    _compare(o, ref, taps, T)


@pytest.mark.parametrize("variant", [2, 7], ids=["full", "packed"])
def test_filter_long_horizon_deferred_downdate(rbslam_lib, variant):
    """T = 600 teacher-forced steps at N = 8, M = 256: the pending (G, KS) pair rides through
    every resampling for the whole run; the drift against the oracle stays inside 1e-8."""
    rb = rbslam_lib
    N, T = 8, 600
    pr, om, gm = _setup(rb, "mag", N, m=253, T=T)
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(12), 1, T, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    forced = np.stack([tp["ai"] for tp in taps])[None].astype(np.int32)
    with rb.Context(gm, N, T, rng_mode=0, kalman_variant=variant) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, forced_ancestors=forced, taps=True)
    _compare(o, ref, taps, T)


@pytest.mark.parametrize("fam,N,kw", [("radio", 20, {"m": 40}), ("mag", 12, {"m": 253, "T": 6}),
                                      ("sparse", 16, {"T": 12})])
def test_particlefilter_makeplots_hook(rbslam_lib, fam, N, kw):
    """The drop-in forwards makePlots every step with the reference's nine arguments
    (src/particleFilter.m:215-217): xn, xl(:,iw_max), P(:,:,iw_max), traj_max, yhattraj, xn_traj,
    traj_mean, xl, P -- NaN / zero in the columns of steps not yet run, as in the reference."""
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = min(pr["y"].shape[0], 8)
    pr = dict(pr, y=pr["y"][:T], odometry=pr["odometry"][:T])
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(2), 1, T, N, om.nz)
    want, got = [], []

    def keep(dst):
        def f(xn, xl_max, P_max, traj_max, yhattraj, xn_traj, traj_mean, xl, P):
            dst.append([np.array(v, dtype=np.float64).copy() for v in
                        (xn, xl_max, P_max, traj_max, yhattraj, xn_traj, traj_mean, xl, P)])
        return f
    oracle.particleFilter(om, *_args(pr), N, pr["dt"], st, makePlots=keep(want))
    rb.particleFilter(gm.dynModel, gm.measModel, pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"],
                      pr["P0_lin"], pr["Q"], pr["R"], N, pr["dt"], bool(gm.sparse), keep(got), rng=st)
    assert len(got) == len(want) == T
    names = ["xn", "xl_max", "P_max", "traj_max", "yhattraj", "xn_traj", "traj_mean", "xl", "P"]
    for t in range(T):
        for k, a, b in zip(names, got[t], want[t]):
            if k == "P":
                b = b.transpose(1, 2, 0)            # the oracle keeps P as [N, M, M]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (t, k)
            assert_close_norm(np.nan_to_num(a), np.nan_to_num(b), 1e-7 if k == "traj_mean" else TOL, "%s@%d" % (k, t))
