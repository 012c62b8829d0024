"""Parity of the CUDA smoothers (rbslam_smoother_run, covariance and information form)
against the oracle restatements of src/particleSmoother.m and
src/particleSmootherInformationForm.m."""
import numpy as np
import pytest

import oracle
from conftest import assert_close_norm

pytestmark = pytest.mark.gpu


def _setup(rb, fam, N, **kw):
    s = rb.synth
    if fam == "radio":
        pr = s.dense_radio_problem(kw.get("traj", "line_3D"), m=kw.get("m", 128), seed=2, m_sim=400)
        om = oracle.DenseRadio2D(pr["NN"], pr["L"])
    elif fam == "mag":
        pr = s.dense_mag_problem(N_T=kw.get("T", 12), m=kw.get("m", 64), seed=3, m_sim=300)
        om = oracle.DenseMag3D(pr["NN"], pr["L"])
    else:
        pr = s.sparse_visual_problem(N_T=kw.get("T", 30), n_landmarks=kw.get("nl", 20), N_P=N, seed=4,
                                     guess_map_var=0.01)
        om = oracle.SparseVisual2D(pr["n_landmarks"], *pr["camera"])
    return pr, om, rb.models.from_problem(pr)


def _args(pr):
    return (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])


def _oracle_run(form, om, pr, N, K, st):
    rec = {}
    fn = oracle.particleSmootherInformationForm if form else oracle.particleSmoother
    out = fn(om, *_args(pr), N, K, pr["dt"], st, record=rec)
    T = pr["y"].shape[0]
    ai = np.zeros((K, T, N), dtype=np.int32)
    for (k, t), v in rec["ai"].items():
        ai[k, t] = v
    ak = np.array([rec["ak"][k] for k in range(K)], dtype=np.int32)
    return out, ai, ak, rec["paNt"]


def _check(o, ref, paNt, K, T, tol=1e-8, ai_tol=1e-6):
    XNK, XLK, PK = ref
    assert_close_norm(o["XNK"], XNK, tol, "XNK")
    assert_close_norm(o["XLK"], XLK, tol, "XLK")
    assert_close_norm(o["PK"], PK, tol, "PK")
    n = 0
    for (k, t), p in paNt.items():
        if p is not None:
            assert_close_norm(o["AI"][:, t, k], p, ai_tol, "AI k=%d t=%d" % (k, t))
            n += 1
    assert n == (K - 1) * (T - 1)


CASES = [
    ("radio", 30, 3, {}),                         # C2 shape (M=128, T=32): shared-memory Kalman kernel
    ("radio", 16, 2, {"traj": "square_3D", "m": 40}),
    ("mag", 10, 2, {"m": 64, "T": 10}),           # d=3, future system up to 27 x 27
    ("mag", 8, 2, {"m": 253, "T": 8}),            # streaming Kalman kernel + flush before K6
    ("sparse", 8, 2, {"T": 30}),                  # re-linearised future Jacobians, NaN rows
]


@pytest.mark.parametrize("fam,N,K,kw", CASES)
def test_smoother_cov_teacher_forced(rbslam_lib, fam, N, K, kw):
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(7), K, T, N, om.nz)
    ref, ai, ak, paNt = _oracle_run(0, om, pr, N, K, st)
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.smoother_run(*_args(pr), pr["dt"], K, 0, streams=st, forced_ancestors=ai, forced_ak=ak,
                             want_AI=True)
    _check(o, ref, paNt, K, T)


@pytest.mark.parametrize("fam,N,K,kw", [CASES[0], CASES[2], CASES[4]])
def test_smoother_cov_free_running(rbslam_lib, fam, N, K, kw):
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(8), K, T, N, om.nz)
    ref, ai, ak, paNt = _oracle_run(0, om, pr, N, K, st)
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.smoother_run(*_args(pr), pr["dt"], K, 0, streams=st, want_AI=True)
    assert np.array_equal(o["ak"], ak)
    _check(o, ref, paNt, K, T)


@pytest.mark.parametrize("fam,N,K,kw", [("radio", 20, 3, {"m": 40}), ("radio", 12, 2, {}),
                                        ("mag", 8, 2, {"m": 40, "T": 8}), ("mag", 6, 2, {"m": 253, "T": 6})])
def test_smoother_information_form(rbslam_lib, chol_shape, fam, N, K, kw):
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(9), K, T, N, om.nz)
    ref, ai, ak, paNt = _oracle_run(1, om, pr, N, K, st)
    with rb.Context(gm, N, T, rng_mode=0, information_form=True) as ctx:
        o = ctx.smoother_run(*_args(pr), pr["dt"], K, 1, streams=st, forced_ancestors=ai, forced_ak=ak,
                             want_AI=True)
    _check(o, ref, paNt, K, T, ai_tol=1e-5)
    # the information-form state itself (Imat, ivec, halfLogDetP) after the last sweep
    st2, lws = {}, {}

    def tap(k, t, d):
        st2.update(ivec=d["ivec"], Imat=np.array(d["Imat"]), hld=d["halfLogDetP"])
        lws[(k, t)] = d["w"].copy()
    oracle.particleSmootherInformationForm(om, *_args(pr), N, K, pr["dt"], st,
                                           forced=dict(ai=ai, ak=ak), tap=tap)
    with rb.Context(gm, N, T, rng_mode=0, information_form=True) as ctx:
        def cb(k, t):
            if t == T:        # sweep-complete notification (outputs of sweep k are written)
                return
            assert_close_norm(ctx.read_particles(P=False)["w"], lws[(k, t)], 1e-6, "w k=%d t=%d" % (k, t))
        ctx.set_step_callback(cb)
        ctx.smoother_run(*_args(pr), pr["dt"], K, 1, streams=st, forced_ancestors=ai, forced_ak=ak)
        inf = ctx.read_information()
    assert_close_norm(inf["ivec"], st2["ivec"], 1e-8, "ivec")
    assert_close_norm(inf["Imat"], st2["Imat"].transpose(1, 2, 0), 1e-8, "Imat")
    assert_close_norm(inf["halfLogDetP"], st2["hld"], 1e-8, "halfLogDetP")


def test_smoother_dropin_signatures(rbslam_lib):
    rb = rbslam_lib
    N, K = 12, 2
    pr, om, gm = _setup(rb, "radio", N, m=30)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(1), K, T, N, om.nz)
    ref, ai, ak, _ = _oracle_run(0, om, pr, N, K, st)
    XNK, XLK, PK = rb.particleSmoother(gm.dynModel, gm.measModel, gm.dynResNorm, pr["odometry"], pr["y"],
                                       pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"], N, K,
                                       pr["dt"], False, None, rng=st)
    assert XNK.shape == (3, T, K) and XLK.shape == (30, K) and PK.shape == (30, 30, K)
    assert_close_norm(XNK, ref[0], 1e-7)
    out = rb.particleSmootherInformationForm(gm.dynModel, gm.measModel, gm.dynResNorm, pr["odometry"],
                                             pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"],
                                             pr["R"], N, K, pr["dt"], False, None, rng=st)
    assert_close_norm(out[0], ref[0], 1e-7)      # cross-form identity on the device
    pr2, om2, gm2 = _setup(rb, "sparse", 6, T=10)
    assert rb.particleSmootherInformationForm(gm2.dynModel, gm2.measModel, None, pr2["odometry"], pr2["y"],
                                              pr2["x0_nonLin"], pr2["x0_lin"], pr2["P0_lin"], pr2["Q"],
                                              pr2["R"], 6, 2, pr2["dt"], True) is None


# ---------------------------------------------------------------------------
# K6 / K7 alone at the BASELINE shapes (rbslam_op_ancestor_weights): the oracle's per-particle
# formulas for 8 particles; a whole smoother run at these sizes is out of the oracle's reach.
# ---------------------------------------------------------------------------
def _c1_state(rb, m, T, N, seed):
    """A plausible mid-run particle state at the C1 / C5 slab size: N particles after a few
    filter steps of the oracle (different poses -> different P_i, xl_i, Imat_i, ivec_i)."""
    pr = rb.synth.dense_mag_problem(N_T=T, m=m, seed=seed, m_sim=300)
    om = oracle.DenseMag3D(pr["NN"], pr["L"])
    rng = np.random.default_rng(seed)
    M = m + 3
    xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1) + 0.3 * rng.standard_normal((7, N))
    P = np.repeat(pr["P0_lin"][None], N, axis=0).copy()
    xl = np.zeros((M, N))
    Imat = np.repeat(np.diag(1.0 / np.diag(pr["P0_lin"]))[None], N, axis=0).copy()
    ivec = np.zeros((M, N))
    Ri = np.linalg.inv(pr["R"])
    for t in range(4):                       # four Kalman updates at perturbed poses
        xs = xn + 0.05 * t
        H = om.measModel(xs)
        for i in range(N):
            S = H[i] @ P[i] @ H[i].T + pr["R"]
            K = P[i] @ H[i].T @ np.linalg.inv(S)
            xl[:, i] += K @ (pr["y"][t] - H[i] @ xl[:, i])
            P[i] = P[i] - K @ S @ K.T
            Imat[i] += H[i].T @ Ri @ H[i]
            ivec[:, i] += H[i].T @ Ri @ pr["y"][t]
    return pr, om, P, xl, Imat, ivec


@pytest.fixture(params=["wide", "narrow"])
def chol_shape(request, monkeypatch):
    """The factorisation kernel runs with 512 threads per matrix for small batches and with 128 for large ones
    (launch_chol); every kernel-level case below runs with both."""
    if request.param == "narrow":
        monkeypatch.setenv("RBSLAM_CHOL_WIDE_MAX", "0")
    else:
        monkeypatch.delenv("RBSLAM_CHOL_WIDE_MAX", raising=False)
    return request.param


def test_ancestor_weights_cov_c1_size(rbslam_lib, chol_shape):
    """K6 at the C1 shape: M = 515, the full future system of T = 192 steps (d tau = 576)."""
    rb = rbslam_lib
    N, m, T = 8, 512, 193
    pr, om, P, xl, _, _ = _c1_state(rb, m, T, N, seed=21)
    gm = rb.models.from_problem(pr)
    rng = np.random.default_rng(2)
    xnk = np.repeat(pr["x0_nonLin"][:, None], T - 1, axis=1) + 0.2 * rng.standard_normal((7, T - 1))
    D = om.measModel(xnk).reshape(-1, m + 3)            # stacked future Jacobians, time-major, d rows per step
    yfut = pr["y"][1:T].reshape(-1)
    ne = D.shape[0]
    assert ne == 576
    RR = np.kron(np.eye(T - 1), pr["R"])
    ref = np.zeros(N)
    for i in range(N):                                  # src/particleSmoother.m:191-229
        SS = D @ P[i] @ D.T + RR
        e = yfut - D @ xl[:, i]
        cS = np.linalg.cholesky(SS)
        v = np.linalg.solve(cS, e)
        ref[i] = -np.sum(np.log(np.diag(cS))) - 0.5 * (v @ v) - ne / 2 * np.log(2 * np.pi)
    with rb.Context(gm, N, 4) as ctx:
        got = ctx.op_ancestor_weights(0, P.transpose(1, 2, 0), xl, D, yfut, R=pr["R"], jitter=1e-2)
    assert np.all(np.abs(got - ref) <= 1e-8 * np.maximum(1.0, np.abs(ref))), (got, ref)


def test_ancestor_weights_info_c5_size(rbslam_lib, chol_shape):
    """K7 at the C1 / C5 shape: batched 515 x 515 Cholesky of Imat_i + ImatAddt, forward solve,
    quadratic forms (src/particleSmootherInformationForm.m:225-236)."""
    rb = rbslam_lib
    N, m = 8, 512
    pr, om, P, xl, Imat, ivec = _c1_state(rb, m, 40, N, seed=22)
    gm = rb.models.from_problem(pr)
    rng = np.random.default_rng(3)
    M = m + 3
    xnk = np.repeat(pr["x0_nonLin"][:, None], 30, axis=1) + 0.2 * rng.standard_normal((7, 30))
    Hk = om.measModel(xnk)
    Ri = np.linalg.inv(pr["R"])
    ImatAddt = sum(Hk[j].T @ Ri @ Hk[j] for j in range(30))
    ivecAddt = sum(Hk[j].T @ Ri @ pr["y"][5 + j] for j in range(30))
    q2 = np.array([ivec[:, i] @ P[i] @ ivec[:, i] for i in range(N)])
    hld = np.array([0.5 * np.linalg.slogdet(P[i])[1] for i in range(N)])
    ref = np.zeros(N)
    for i in range(N):
        cI = np.linalg.cholesky(Imat[i] + ImatAddt)
        vI = np.linalg.solve(cI, ivec[:, i] + ivecAddt)
        ref[i] = -0.5 * q2[i] - hld[i] - np.sum(np.log(np.diag(cI))) + 0.5 * (vI @ vI)
    with rb.Context(gm, N, 4, information_form=True) as ctx:
        got = ctx.op_ancestor_weights(1, Imat.transpose(1, 2, 0), ivec, ImatAddt, ivecAddt, q2=q2, hld=hld,
                                      jitter=-1.0)
    assert np.all(np.abs(got - ref) <= 1e-8 * np.maximum(1.0, np.abs(ref))), (got, ref)


@pytest.mark.parametrize("m,nfut", [(5, 1), (29, 2), (30, 11), (61, 21), (62, 22), (93, 32), (125, 43), (126, 5)])
def test_ancestor_weights_panel_edges(rbslam_lib, chol_shape, m, nfut):
    """The blocked factorisation at orders around its block sizes (panels of 32 columns, row tiles of 64, the
    right-hand side as an extra row): K7 with n = M = m + 3 in {8, 32, 33, 64, 65, 96, 128, 129} and K6 with
    n = 3 nfut in {3, 6, 33, 63, 66, 96, 129, 15}: single narrow panel, exact multiples (the right-hand-side row
    opens a new tile), one past a multiple (one-column last panel)."""
    rb = rbslam_lib
    N = 5
    pr, om, P, xl, Imat, ivec = _c1_state(rb, m, max(8, nfut + 2), N, seed=100 + m)
    gm = rb.models.from_problem(pr)
    rng = np.random.default_rng(m)
    M = m + 3
    xnk = np.repeat(pr["x0_nonLin"][:, None], nfut, axis=1) + 0.2 * rng.standard_normal((7, nfut))
    Hk = om.measModel(xnk)
    Ri = np.linalg.inv(pr["R"])
    yk = pr["y"][1:1 + nfut]
    # information form
    ImatAddt = sum(Hk[j].T @ Ri @ Hk[j] for j in range(nfut))
    ivecAddt = sum(Hk[j].T @ Ri @ yk[j] for j in range(nfut))
    q2 = np.array([ivec[:, i] @ P[i] @ ivec[:, i] for i in range(N)])
    hld = np.array([0.5 * np.linalg.slogdet(P[i])[1] for i in range(N)])
    ref1 = np.zeros(N)
    for i in range(N):
        cI = np.linalg.cholesky(Imat[i] + ImatAddt)
        vI = np.linalg.solve(cI, ivec[:, i] + ivecAddt)
        ref1[i] = -0.5 * q2[i] - hld[i] - np.sum(np.log(np.diag(cI))) + 0.5 * (vI @ vI)
    with rb.Context(gm, N, 4, information_form=True) as ctx:
        got1 = ctx.op_ancestor_weights(1, Imat.transpose(1, 2, 0), ivec, ImatAddt, ivecAddt, q2=q2, hld=hld, jitter=-1.0)
    assert np.all(np.abs(got1 - ref1) <= 1e-8 * np.maximum(1.0, np.abs(ref1))), (got1, ref1)
    # covariance form
    D = Hk.reshape(-1, M)
    yfut = yk.reshape(-1)
    ne = D.shape[0]
    RR = np.kron(np.eye(nfut), pr["R"])
    ref0 = np.zeros(N)
    for i in range(N):
        cS = np.linalg.cholesky(D @ P[i] @ D.T + RR)
        v = np.linalg.solve(cS, yfut - D @ xl[:, i])
        ref0[i] = -np.sum(np.log(np.diag(cS))) - 0.5 * (v @ v) - ne / 2 * np.log(2 * np.pi)
    with rb.Context(gm, N, 4) as ctx:
        got0 = ctx.op_ancestor_weights(0, P.transpose(1, 2, 0), xl, D, yfut, R=pr["R"], jitter=1e-2)
    assert np.all(np.abs(got0 - ref0) <= 1e-8 * np.maximum(1.0, np.abs(ref0))), (got0, ref0)


def test_ancestor_weights_not_positive_definite_is_reported(rbslam_lib, chol_shape):
    """An indefinite matrix in the batch: the information form raises (quirk Q7: no retry), with the particle
    named; the others in the batch are unaffected by it on the next call."""
    rb = rbslam_lib
    N, m = 4, 61
    pr, om, P, xl, Imat, ivec = _c1_state(rb, m, 8, N, seed=7)
    gm = rb.models.from_problem(pr)
    M = m + 3
    bad = Imat.copy()
    bad[2, 40, 40] = -5.0                      # a negative pivot in the second panel of particle 2
    q2, hld = np.zeros(N), np.zeros(N)
    with rb.Context(gm, N, 4, information_form=True) as ctx:
        with pytest.raises(rb.RbslamError, match="particle 2"):
            ctx.op_ancestor_weights(1, bad.transpose(1, 2, 0), ivec, np.zeros((M, M)), np.zeros(M), q2=q2, hld=hld, jitter=-1.0)
        ok = ctx.op_ancestor_weights(1, Imat.transpose(1, 2, 0), ivec, np.zeros((M, M)), np.zeros(M), q2=q2, hld=hld, jitter=-1.0)
    ref = np.array([-np.sum(np.log(np.diag(np.linalg.cholesky(Imat[i])))) +
                    0.5 * np.sum(np.linalg.solve(np.linalg.cholesky(Imat[i]), ivec[:, i]) ** 2) for i in range(N)])
    assert np.all(np.abs(ok - ref) <= 1e-8 * np.maximum(1.0, np.abs(ref)))


def test_ancestor_weights_cov_jitter_retry(rbslam_lib, chol_shape):
    """K6 with a system matrix that is not positive definite until the reference's retry adds jitter * I
    (src/particleSmoother.m:221-226): one particle's covariance gets a small negative direction; the batch
    must factor that matrix in the second attempt (same result as the oracle's chol_jitter) and the others
    in the first."""
    rb = rbslam_lib
    N, m, nfut = 4, 61, 30                       # n = 90: three panels, the last one ragged
    pr, om, P, xl, _, _ = _c1_state(rb, m, nfut + 2, N, seed=31)
    gm = rb.models.from_problem(pr)
    rng = np.random.default_rng(5)
    M = m + 3
    xnk = np.repeat(pr["x0_nonLin"][:, None], nfut, axis=1) + 0.2 * rng.standard_normal((7, nfut))
    D = om.measModel(xnk).reshape(-1, M)
    yfut = pr["y"][1:1 + nfut].reshape(-1)
    ne = D.shape[0]
    R = 1e-3 * np.eye(3)                         # small measurement noise: the perturbation below decides the sign
    RR = np.kron(np.eye(nfut), R)
    # particle 1: P1 <- P1 - alpha v v' with D v = z (a unit vector), i.e. S <- S - alpha z z'; alpha by bisection
    # so that the smallest eigenvalue of S lands at -0.3 * jitter
    jitter = 1e-2
    U, sv, Vt = np.linalg.svd(D, full_matrices=False)
    v = Vt[0] / sv[0]
    S1 = D @ P[1] @ D.T + RR
    lo, hi = 0.0, 10.0 * np.linalg.eigvalsh(S1)[-1]
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if np.linalg.eigvalsh(S1 - mid * np.outer(U[:, 0], U[:, 0]))[0] > -0.3 * jitter:
            lo = mid
        else:
            hi = mid
    P = P.copy()
    P[1] = P[1] - lo * np.outer(v, v)
    lam = np.linalg.eigvalsh(D @ P[1] @ D.T + RR)
    assert lam[0] < 0 and lam[0] > -jitter, lam[:3]
    ref = np.zeros(N)
    used = []
    for i in range(N):
        SS = D @ P[i] @ D.T + RR
        cS, uj = oracle.tools.chol_jitter(SS, jitter)
        used.append(uj)
        v = np.linalg.solve(cS, yfut - D @ xl[:, i])
        ref[i] = -np.sum(np.log(np.diag(cS))) - 0.5 * (v @ v) - ne / 2 * np.log(2 * np.pi)
    assert used == [False, True, False, False]
    with rb.Context(gm, N, 4) as ctx:
        got = ctx.op_ancestor_weights(0, P.transpose(1, 2, 0), xl, D, yfut, R=R, jitter=jitter)
    assert np.all(np.abs(got - ref) <= 1e-7 * np.maximum(1.0, np.abs(ref))), (got, ref)


def test_particlesmoother_makeplots_and_progress(rbslam_lib, capsys):
    """The smoother drop-in calls makePlots(xnk,xlk,k,XNK,XLK,PK) and prints the progress line once
    per sweep, when that sweep's outputs exist (src/particleSmoother.m:359-365)."""
    rb = rbslam_lib
    N, K = 10, 3
    pr, om, gm = _setup(rb, "radio", N, m=40)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(4), K, T, N, om.nz)
    want, got = [], []
    oracle.particleSmoother(om, *_args(pr), N, K, pr["dt"], st,
                            makePlots=lambda xnk, xlk, k, XNK, XLK, PK: want.append((xnk.copy(), xlk.copy(), k, XNK.copy())))
    XNK, XLK, PK = rb.particleSmoother(gm.dynModel, gm.measModel, gm.dynResNorm, pr["odometry"], pr["y"],
                                       pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"], N, K,
                                       pr["dt"], False,
                                       lambda xnk, xlk, k, XNK, XLK, PK: got.append((xnk.copy(), xlk.copy(), k, XNK.copy())),
                                       rng=st)
    assert [g[2] for g in got] == [w[2] for w in want] == list(range(K))
    for g, w in zip(got, want):
        assert_close_norm(g[0], w[0], 1e-8, "xnk")
        assert_close_norm(g[1], w[1], 1e-8, "xlk")
        assert_close_norm(g[3][:, :, :g[2] + 1], w[3][:, :, :w[2] + 1], 1e-8, "XNK so far")
    out = capsys.readouterr().out
    assert out.count("Particle smoother iteration") == K and "iteration %d/%d done." % (K, K) in out


@pytest.mark.parametrize("form,fam,N,K,kw,world", [
    (1, "mag", 10, 3, {"m": 64, "T": 10}, 2),       # information form: K7 split over two replicas
    (0, "mag", 9, 2, {"m": 64, "T": 8}, 2),         # covariance form, N not divisible by the replica count
    (1, "radio", 16, 2, {"traj": "square_3D", "m": 40}, 4),
    (0, "sparse", 8, 2, {"T": 20}, 2),              # re-linearised future Jacobians per particle
])
def test_smoother_replicas_match_single_gpu(rbslam_lib, form, fam, N, K, kw, world):
    """rbslam_create_replicas: the ancestor weights of each time step are evaluated block-wise on
    several devices (all on device 0 on a 1-GPU box) and all-gathered over peer memory; the run
    must reproduce the single-GPU smoother exactly (same kernels on the same numbers)."""
    rb = rbslam_lib
    from rbslam import _capi
    ndev = _capi.lib().rbslam_device_count()
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(11), K, T, N, om.nz)
    with rb.Context(gm, N, T, rng_mode=0, information_form=(form == 1)) as ctx:
        one = ctx.smoother_run(*_args(pr), pr["dt"], K, form, streams=st, want_AI=True)
    with rb.Context(gm, N, T, rng_mode=0, information_form=(form == 1), replicas=True,
                    devices=[r % ndev for r in range(world)]) as ctx:
        rep = ctx.smoother_run(*_args(pr), pr["dt"], K, form, streams=st, want_AI=True)
        rep2 = ctx.smoother_run(*_args(pr), pr["dt"], K, form, streams=st)      # reusable
    assert np.array_equal(rep["ak"], one["ak"])
    for k in ["XNK", "XLK", "PK"]:
        assert_close_norm(rep[k], one[k], 1e-12, "replicas vs single: " + k)
        assert np.array_equal(rep2[k], rep[k]), k
    assert_close_norm(np.nan_to_num(rep["AI"]), np.nan_to_num(one["AI"]), 1e-12, "AI")
