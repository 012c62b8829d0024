"""Kernel-level parity: each CUDA kernel, driven alone through the C ABI
(rbslam_op_*), against the NumPy oracle on the same seeded inputs."""
import numpy as np
import pytest

import oracle
from conftest import assert_close_norm

pytestmark = pytest.mark.gpu


def _problem(rbslam, fam, **kw):
    s = rbslam.synth
    if fam == "radio":
        pr = s.dense_radio_problem("line_3D", m=kw.get("m", 128), seed=2, m_sim=400)
        om = oracle.DenseRadio2D(pr["NN"], pr["L"])
    elif fam == "mag":
        pr = s.dense_mag_problem(N_T=kw.get("T", 24), m=kw.get("m", 64), seed=3, m_sim=300)
        om = oracle.DenseMag3D(pr["NN"], pr["L"])
    else:
        pr = s.sparse_visual_problem(N_T=kw.get("T", 40), n_landmarks=20, N_P=kw.get("N", 32), seed=4,
                                     guess_map_var=0.01)
        om = oracle.SparseVisual2D(pr["n_landmarks"], *pr["camera"])
    return pr, om, rbslam.models.from_problem(pr)


@pytest.mark.parametrize("N", [1, 7, 1000, 10000, 40000])
def test_resample_bit_exact(rbslam_lib, N):
    rb = rbslam_lib
    pr, om, gm = _problem(rb, "radio")
    rng = np.random.default_rng(N)
    w = rng.random(N) ** 4
    w[rng.random(N) < 0.2] = 0.0          # zero-weight particles (ties in cumsum)
    if w.sum() == 0:
        w[0] = 1.0
    w = w / w.sum()
    wc = np.cumsum(w)
    u = np.concatenate([rng.random(3 * N + 5), [0.0, wc[0], wc[N // 2], wc[-1], np.nextafter(wc[-1], 2.0), 1.0],
                        wc[rng.integers(0, N, 50)]])
    with rb.Context(gm, 8, 4) as ctx:
        ai = ctx.op_resample(w, u)
    ref = oracle.tools.sample_many(w, u)
    assert np.array_equal(ai, ref)


@pytest.mark.parametrize("N", [4096, 10000, 80000])
def test_resample_fast_path_is_exact_and_usually_stands(rbslam_lib, N):
    """Large populations: a parallel prefix sum first; every draw must PROVE that the strict left-to-right
    cumsum would give the same ancestor (error bounds of both summations), else the exact scan runs.
    (1) random draws: bit-exact against the sequential oracle, and the exact scan is (almost) never needed;
    (2) draws placed ON the boundaries wc(j) and one ulp beside them: still bit-exact -- through the exact path."""
    rb = rbslam_lib
    pr, om, gm = _problem(rb, "radio")
    rng = np.random.default_rng(N + 1)
    w = rng.random(N) ** 4
    w[rng.random(N) < 0.2] = 0.0
    w = w / w.sum()
    wc = np.cumsum(w)
    with rb.Context(gm, 8, 4) as ctx:
        runs = 0
        for rep in range(8):
            u = rng.random(N)
            ai = ctx.op_resample(w, u)
            assert np.array_equal(ai, oracle.tools.sample_many(w, u))
            runs = ctx.status_counters()["exact_scan_runs"]
        assert runs <= 2, runs                       # expected: ~0.04 ambiguous steps per 8 at N = 80 000
        j = rng.integers(0, N, 400)
        u = np.concatenate([rng.random(N), wc[j], np.nextafter(wc[j], 2.0), np.nextafter(wc[j], -1.0)])
        ai = ctx.op_resample(w, u)
        assert np.array_equal(ai, oracle.tools.sample_many(w, u))
        assert ctx.status_counters()["exact_scan_runs"] == runs + 1


def test_resample_frequencies(rbslam_lib):
    """The reference's own (commented-out) self-test of sample: tools/sample.m:36-64."""
    rb = rbslam_lib
    pr, om, gm = _problem(rb, "radio")
    rng = np.random.default_rng(0)
    w = rng.random(10)
    w /= w.sum()
    with rb.Context(gm, 8, 4) as ctx:
        ai = ctx.op_resample(w, rng.random(100000))
    freq = np.bincount(ai, minlength=10) / 1e5
    assert np.max(np.abs(freq - w)) < 5e-3


@pytest.mark.parametrize("N", [1, 33, 5000, 32768, 50001])   # >= 32768: the chunked multi-CTA form
def test_normalize(rbslam_lib, N):
    rb = rbslam_lib
    pr, om, gm = _problem(rb, "radio")
    rng = np.random.default_rng(N)
    logw = -50 * rng.random(N) - 1000.0
    if N > 2:
        logw[2] = logw[1]                  # tie: first index must win
    with rb.Context(gm, 8, 4) as ctx:
        w, im = ctx.op_normalize(logw)
    ref = oracle.particle_filter.normalise(logw)
    assert_close_norm(w, ref, 1e-12, "w")
    assert im == int(np.argmax(w))
    assert abs(w.sum() - 1) < 1e-12


@pytest.mark.parametrize("fam", ["radio", "mag", "sparse"])
def test_propagate(rbslam_lib, fam):
    rb = rbslam_lib
    pr, om, gm = _problem(rb, fam)
    N = 257
    rng = np.random.default_rng(11)
    xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1) + 0.1 * rng.standard_normal((gm.n, N))
    ai = rng.integers(0, N, N)
    Z = rng.standard_normal((N, gm.nz))
    Q = np.asarray(pr["Q"])
    Q = Q[:, :, 3] if Q.ndim == 3 else Q
    dt = pr["dt"]
    dx = pr["odometry"][2]
    with rb.Context(gm, N, 4) as ctx:
        out = ctx.op_propagate(xn, ai, dx, dt, Q, Z)
    ref = np.stack([om.dynModel(xn[:, ai[i]], dx, dt, Q, Z[i]) for i in range(N)], axis=1)
    assert_close_norm(out, ref, 1e-12, "xn")


@pytest.mark.parametrize("fam", ["radio", "mag", "sparse"])
def test_meas_jacobian(rbslam_lib, fam):
    rb = rbslam_lib
    pr, om, gm = _problem(rb, fam)
    N = 19
    rng = np.random.default_rng(5)
    xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1) + 0.3 * rng.standard_normal((gm.n, N))
    with rb.Context(gm, N, 4) as ctx:
        if fam == "sparse":
            xl = pr["x0_lin"][:, :N] + 0.05 * rng.standard_normal((gm.M, N))
            dy, yhat = ctx.op_meas_jacobian(xn, xl)
            for i in range(N):
                yr, dr = om.measModel_sparse(xn[:, i], xl[:, i])
                assert_close_norm(dy[i], dr, 1e-12, "dy")
                assert_close_norm(yhat[:, i], yr, 1e-12, "yhat")
        else:
            dy, _ = ctx.op_meas_jacobian(xn)
            ref = om.measModel(xn)
            assert_close_norm(dy, ref, 1e-11, "dy")


@pytest.mark.parametrize("m", [64, 512])
def test_jacobian_phi3d(rbslam_lib, m):
    """k_jacobian_phi3d against the oracle's tools/JacobianPhi3D.m restatement, on a centred and
    on an off-centre domain (the reference takes the six bounds as separate scalars)."""
    rb = rbslam_lib
    pr, om, gm = _problem(rb, "mag", m=m)
    rng = np.random.default_rng(6)
    N = 23
    L = np.asarray(pr["L"], dtype=np.float64)
    with rb.Context(gm, 4, 4) as ctx:
        for lo, hi in ((-L, L), (-L + np.array([0.3, -0.2, 0.1]), L + np.array([0.5, 0.4, 0.25]))):
            x = lo[:, None] + (hi - lo)[:, None] * rng.random((3, N))
            J = ctx.op_jacobian_phi3d(x, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])
            ref = oracle.JacobianPhi3D(x, m, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], pr["NN"])
            assert J.shape == ref.shape == (3, 3, m, N)
            assert_close_norm(J, ref, 1e-12, "JacobianPhi3D")
    prr, _, gmr = _problem(rb, "radio")
    with rb.Context(gmr, 4, 4) as ctx:     # 2-D basis: refused, not silently wrong
        with pytest.raises(rb.UnsupportedModelError):
            ctx.op_jacobian_phi3d(np.zeros((3, 1)), -1, 1, -1, 1, -1, 1)


def _rand_spd(rng, M, scale=1.0):
    A = rng.standard_normal((M, M))
    return scale * (A @ A.T / M + 0.5 * np.eye(M))


def _oracle_update(xl, P, H, yt, R, jitter, yhat=None):
    """One particle of src/particleFilter.m:126-151,164-204 via the oracle's helpers."""
    pf = oracle.particle_filter
    e, SS, ind = pf.innovation(yt, H, xl, P, R, yhat)
    logw, cS = pf.log_weight(e, SS, jitter)
    K = pf.kalman_gain(P, H[ind, :], cS)
    return xl + K @ e, P - K @ SS @ K.T, logw


@pytest.mark.parametrize("fam,m,variant", [("radio", 128, 0), ("radio", 50, 0), ("mag", 64, 0),
                                           ("mag", 64, 2), ("mag", 64, 3), ("mag", 253, 0),
                                           ("mag", 253, 3), ("mag", 512, 0), ("radio", 300, 0),
                                           ("radio", 300, 3), ("mag", 1024, 0)])
def test_kalman_update_dense(rbslam_lib, fam, m, variant):
    rb = rbslam_lib
    pr, om, gm = _problem(rb, fam, m=m)
    N = 12
    M, d = gm.M, gm.d
    rng = np.random.default_rng(m)
    xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1) + 0.2 * rng.standard_normal((gm.n, N))
    H = om.measModel(xn)                                    # [N, d, M]
    P = np.stack([_rand_spd(rng, M, 10.0) for _ in range(N)])
    xl = rng.standard_normal((M, N))
    yt = pr["y"][3]
    R = pr["R"]
    with rb.Context(gm, N, 4, kalman_variant=variant) as ctx:
        xl2, P2, logw = ctx.op_kalman_update(xl, P.transpose(1, 2, 0), yt, R, 1e-3, H=H)
        # also with the Jacobian evaluated on the device
        xl3, P3, logw3 = ctx.op_kalman_update(xl, P.transpose(1, 2, 0), yt, R, 1e-3, xn=xn)
    for i in range(N):
        xr, Pr, lr = _oracle_update(xl[:, i], P[i], H[i], yt, R, 1e-3)
        assert_close_norm(xl2[:, i], xr, 1e-8, "xl")
        assert_close_norm(P2[:, :, i], Pr, 1e-8, "P")
        assert abs(logw[i] - lr) <= 1e-8 * max(1.0, abs(lr))
        assert_close_norm(P3[:, :, i], Pr, 1e-8, "P(dev H)")
        assert abs(logw3[i] - lr) <= 1e-8 * max(1.0, abs(lr))


def test_kalman_update_sparse_nan_rows(rbslam_lib):
    rb = rbslam_lib
    pr, om, gm = _problem(rb, "sparse")
    N, M = 32, gm.M
    rng = np.random.default_rng(8)
    xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1) + 0.05 * rng.standard_normal((3, N))
    xl = pr["x0_lin"][:, :N].copy()
    P = np.stack([_rand_spd(rng, M, 0.3) for _ in range(N)])
    for yt in (pr["y"][5], np.full(gm.d, np.nan)):          # some observed / none observed
        with rb.Context(gm, N, 4) as ctx:
            xl2, P2, logw = ctx.op_kalman_update(xl, P.transpose(1, 2, 0), yt, pr["R"], 1e-3, xn=xn)
        for i in range(N):
            yhat, dy = om.measModel_sparse(xn[:, i], xl[:, i])
            xr, Pr, lr = _oracle_update(xl[:, i], P[i], dy, yt, pr["R"], 1e-3, yhat)
            assert_close_norm(xl2[:, i], xr, 1e-8, "xl")
            assert_close_norm(P2[:, :, i], Pr, 1e-8, "P")
            assert abs(logw[i] - lr) <= 1e-8 * max(1.0, abs(lr))


def test_kalman_jitter_and_not_pd(rbslam_lib):
    """chol failure -> retry with jitter (src/particleFilter.m:145-148); second failure -> status 3."""
    rb = rbslam_lib
    pr, om, gm = _problem(rb, "radio", m=50)
    N, M = 4, gm.M
    rng = np.random.default_rng(1)
    xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1)
    H = om.measModel(xn)
    h = H[0, 0]
    # P such that H P H' + R is slightly negative: jitter 1e-3 rescues it
    base = _rand_spd(rng, M, 1.0)
    s = h @ base @ h
    R = np.array([[1e-4]])
    Pbad = base - (1.0 + (R[0, 0] + 5e-4) / s) * np.outer(base @ h, base @ h) / s
    P = np.stack([Pbad] * N)
    xl = np.zeros((M, N))
    with rb.Context(gm, N, 4) as ctx:
        xl2, P2, logw = ctx.op_kalman_update(xl, P.transpose(1, 2, 0), np.array([0.3]), R, 1e-3, H=H)
        xr, Pr, lr = _oracle_update(xl[:, 0], P[0], H[0], np.array([0.3]), R, 1e-3)
        assert abs(logw[0] - lr) <= 1e-8 * max(1.0, abs(lr))
        assert_close_norm(P2[:, :, 0], Pr, 1e-8, "P jitter")
        Pworse = base - 3.0 * np.outer(base @ h, base @ h) / s
        with pytest.raises(rb.RbslamError) as ei:
            ctx.op_kalman_update(xl, np.stack([Pworse] * N).transpose(1, 2, 0), np.array([0.3]), R,
                                 1e-3, H=H)
        assert ei.value.code == 3


@pytest.mark.parametrize("fam", ["radio", "mag", "sparse"])
def test_dyn_logweight(rbslam_lib, fam):
    rb = rbslam_lib
    pr, om, gm = _problem(rb, fam)
    N = 50
    rng = np.random.default_rng(2)
    xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1) + 0.01 * rng.standard_normal((gm.n, N))
    xk = pr["x0_nonLin"] + 0.01 * rng.standard_normal(gm.n)
    Q = np.asarray(pr["Q"])
    Q = Q[:, :, 15] if Q.ndim == 3 else Q
    dx = pr["odometry"][1]
    with rb.Context(gm, N, 4) as ctx:
        out = ctx.op_dyn_logweight(xk, xn, dx, pr["dt"], Q, use_default=(fam == "sparse"))
    for i in range(N):
        if fam == "sparse":
            e = oracle.particle_smoother.default_dyn_res_norm(xk, xn[:, i], dx, pr["dt"], Q)
        else:
            e = om.dynResNorm(xk, xn[:, i], dx, pr["dt"], Q)
        ref = -0.5 * (e @ e)
        assert abs(out[i] - ref) <= 1e-8 * max(1.0, abs(ref)), (i, out[i], ref)
