"""CPU tests that pin the oracle as far as the reference allows.

The reference ships no tests or golden vectors (parity unpinned), so the oracle is
checked against the analytical identities the reference itself states and against
independent extended-precision re-derivations."""
import numpy as np
import pytest

import oracle
from oracle import tools
from oracle.particle_filter import innovation, log_weight, kalman_gain, normalise
from conftest import assert_close_norm, PKG  # noqa: F401

from rbslam import synth, basis


# ---------------------------------------------------------------- sample.m
def test_sample_is_count_of_cumsum_below_u():
    rng = np.random.default_rng(0)
    w = rng.random(37)
    w /= w.sum()
    for u in rng.random(200):
        wc = 0.0
        cnt = 0
        for wi in w:              # literal restatement: sum(cumsum(w) < u) + 1, left to right
            wc += wi
            cnt += wc < u
        assert tools.sample(w, u) == min(cnt, 36)
    assert np.array_equal(tools.sample_many(w, [0.0, 1.0 - 1e-17, 0.5]),
                          [tools.sample(w, 0.0), tools.sample(w, 1.0 - 1e-17), tools.sample(w, 0.5)])


def test_sample_frequencies_selftest():
    """tools/sample.m:36-64 (commented-out self test): empirical frequency ~ w."""
    rng = np.random.default_rng(1)
    w = rng.random(10)
    w /= w.sum()
    idx = tools.sample_many(w, rng.random(100000))
    assert np.max(np.abs(np.bincount(idx, minlength=10) / 1e5 - w)) < 5e-3


# ---------------------------------------------------------------- quaternions
def test_quaternion_identities():
    rng = np.random.default_rng(2)
    for _ in range(50):
        phi = rng.standard_normal(3) * 0.3
        assert np.linalg.norm(phi) < np.pi / 2
        q = tools.expq(phi)
        assert abs(np.linalg.norm(q) - 1) < 1e-14
        assert_close_norm(tools.logq(q), phi, 1e-12)           # log(exp(phi)) = phi (|phi|<pi/2)
        R = tools.quat2rmat(q)
        assert_close_norm(R @ R.T, np.eye(3), 1e-13)
        assert abs(np.linalg.det(R) - 1) < 1e-13
        p = tools.expq(rng.standard_normal(3))
        # rotation of a product = product of rotations; qLeft(q)*qInv(q) = identity
        assert_close_norm(tools.quat2rmat(tools.qLeft(q) @ p), R @ tools.quat2rmat(p), 1e-13)
        assert_close_norm(tools.qLeft(q) @ tools.qInv(q), [1, 0, 0, 0], 1e-14)
    assert np.array_equal(tools.expq(np.zeros(3)), [1, 0, 0, 0])   # mag_phi == 0 guard
    assert tools.expq(np.array([0, 0, 2.0]))[0] > 0                # sign flip when cos < 0
    v = rng.standard_normal(3)
    u = rng.standard_normal(3)
    assert_close_norm(tools.mcross(v) @ u, np.cross(v, u), 1e-15)


# ---------------------------------------------------------------- eigenbasis
def _basis_by_definition(m, LL):
    """The index selection of tools/domain_cartesian_dx.m:26-43 straight from its definition, written
    independently of both restatements: all index tuples of the grid, eigenvalues
    sum((pi n_j / (2 L_j))^2), stable ascending sort, first m."""
    import itertools
    L = (LL[1] - LL[0]) / 2.0
    d = L.shape[0]
    grid = np.ceil(m ** (1.0 / d) * L / L.min()).astype(int)                  # :33
    # ndgridm enumerates (1,1),(1,2),...,(1,N2),(2,1),...: the LAST index runs fastest (:186-188, 207-214)
    tuples = list(itertools.product(*[range(1, g + 1) for g in grid]))
    lam = [sum((np.pi * n / (2.0 * Lj)) ** 2 for n, Lj in zip(t, L)) for t in tuples]
    order = sorted(range(len(tuples)), key=lambda k: lam[k])                  # Python's sort is stable, like MATLAB's
    return L, np.array([tuples[k] for k in order[:m]])


def test_domain_cartesian_dx_against_its_definition():
    """Both restatements of the index selection (oracle/tools.py and the product's rbslam/basis.py) against
    an independent enumeration from the definition, incl. the stable order among equal eigenvalues
    (a square domain makes (a,b) / (b,a) ties)."""
    for m, LL in [(100, np.array([[-3.0, -2.0, -1.0], [5.0, 2.5, 1.0]])),
                  (60, np.array([[-2.0, -2.0], [2.0, 2.0]])),                 # ties: L_1 == L_2
                  (200, np.array([[-12.02, -12.02, -2.4], [12.02, 12.02, 2.4]]))]:
        d = LL.shape[1]
        Lr, NNr = _basis_by_definition(m, LL)
        for impl in (tools.domain_cartesian_dx, basis.domain_cartesian_dx):
            L, NN = impl(m, d, LL)
            assert np.allclose(L, Lr) and np.array_equal(np.asarray(NN, dtype=int), NNr), impl.__module__
        lam = tools.eigenval(NNr.astype(float), Lr)
        assert np.all(np.diff(lam) >= 0) and len({tuple(r) for r in NNr}) == m
    # C1/C4 claims of SURVEY 8a row A12: max index (18,18,3) for m=512 and (22,22,4) for m=1024
    pr_LL = np.array([[-12.02, -12.02, -2.4], [12.02, 12.02, 2.4]])
    assert tuple(basis.domain_cartesian_dx(512, 3, pr_LL)[1].max(0)) == (18, 18, 3)
    assert tuple(basis.domain_cartesian_dx(1024, 3, pr_LL)[1].max(0)) == (22, 22, 4)


def test_eigenfun_dx_is_gradient_and_basis_is_orthonormal():
    LL = np.array([[-2.0, -1.5], [2.0, 1.5]])
    L, NN = tools.domain_cartesian_dx(12, 2, LL)
    rng = np.random.default_rng(3)
    x = (rng.random((5, 2)) - 0.5) * L
    h = 1e-6
    for di in range(2):
        e = np.zeros(2)
        e[di] = h
        fd = (tools.eigenfun(NN, x + e, L) - tools.eigenfun(NN, x - e, L)) / (2 * h)
        assert_close_norm(tools.eigenfun_dx(NN, x, di, L), fd, 1e-8)
    g = np.linspace(-1, 1, 401)
    X, Y = np.meshgrid(g * L[0], g * L[1], indexing="ij")
    Phi = tools.eigenfun(NN, np.stack([X.ravel(), Y.ravel()], 1), L)
    wq = np.ones(401)
    wq[0] = wq[-1] = 0.5
    W = np.outer(wq, wq).ravel() * (2 * L[0] / 400) * (2 * L[1] / 400)
    assert_close_norm(Phi.T @ (Phi * W[:, None]), np.eye(12), 1e-3)


def test_jacobian_phi3d_is_hessian_of_the_basis():
    """tools/JacobianPhi3D.m against central differences of eigenfun_dx (centred domain:
    a = -L, b = +L, so both files describe the same basis), and symmetric in (r, c)."""
    LL = np.array([[-3.0, -2.0, -1.0], [3.0, 2.0, 1.0]])
    L, NN = tools.domain_cartesian_dx(20, 3, LL)
    rng = np.random.default_rng(11)
    x = (rng.random((4, 3)) - 0.5) * 1.6 * L
    J = tools.JacobianPhi3D(x.T, 20, LL[0, 0], LL[1, 0], LL[0, 1], LL[1, 1], LL[0, 2], LL[1, 2], NN)
    assert J.shape == (3, 3, 20, 4)
    assert_close_norm(J, J.transpose(1, 0, 2, 3), 1e-15)
    h = 1e-6
    for c in range(3):
        e = np.zeros(3)
        e[c] = h
        for r in range(3):
            fd = (tools.eigenfun_dx(NN, x + e, r, L) - tools.eigenfun_dx(NN, x - e, r, L)) / (2 * h)
            assert_close_norm(J[r, c].T, fd, 1e-7, "J[%d,%d]" % (r, c))


# ---------------------------------------------------------------- Kalman update / log-weight
def _rand_spd(rng, M):
    A = rng.standard_normal((M, M))
    return A @ A.T / M + 0.3 * np.eye(M)


def test_logweight_matches_direct_formula():
    """The formula the reference leaves as a comment: -.5*log(det(SS)) - .5*(e'*(SS\\e))
    (src/particleSmoother.m:228,285) minus the 2*pi term."""
    rng = np.random.default_rng(4)
    for d in (1, 3, 7):
        M = 11
        H = rng.standard_normal((d, M))
        P = _rand_spd(rng, M)
        R = _rand_spd(rng, d) * 0.1
        xl = rng.standard_normal(M)
        y = rng.standard_normal(d)
        e, SS, _ = innovation(y, H, xl, P, R)
        lw, _ = log_weight(e, SS, 1e-3)
        direct = -0.5 * np.log(np.linalg.det(SS)) - 0.5 * e @ np.linalg.solve(SS, e) \
            - 0.5 * d * np.log(2 * np.pi)
        assert abs(lw - direct) < 1e-11 * max(1, abs(direct))


def test_kalman_update_extended_precision():
    """Independent re-derivation of src/particleFilter.m:184-198 in 50-digit arithmetic:
    P+ = P - P H'(H P H' + R)^-1 H P,  xl+ = xl + P H' (..)^-1 e."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(5)
    M, d = 6, 3
    H = rng.standard_normal((d, M))
    P = _rand_spd(rng, M)
    R = 0.2 * np.eye(d)
    xl = rng.standard_normal(M)
    y = rng.standard_normal(d)
    e, SS, ind = innovation(y, H, xl, P, R)
    _, cS = log_weight(e, SS, 1e-3)
    K = kalman_gain(P, H, cS)
    xl_new = xl + K @ e
    P_new = P - K @ SS @ K.T
    mH, mP, mR = mp.matrix(H.tolist()), mp.matrix(P.tolist()), mp.matrix(R.tolist())
    mS = mH * mP * mH.T + mR
    mK = mP * mH.T * mp.inverse(mS)
    me = mp.matrix(y.tolist()) - mH * mp.matrix(xl.tolist())
    ref_xl = np.array((mp.matrix(xl.tolist()) + mK * me).tolist(), dtype=float).ravel()
    ref_P = np.array((mP - mK * mS * mK.T).tolist(), dtype=float)
    assert_close_norm(xl_new, ref_xl, 1e-13)
    assert_close_norm(P_new, ref_P, 1e-13)


def test_sequential_filter_equals_batch_gp():
    """Closed-form anchor (tools/gp_scalar_potential_fast.m:190-193): with one particle,
    fixed poses and no process noise the sequential Kalman updates of
    src/particleFilter.m:184-198 reproduce the batch reduced-rank GP posterior
    mean (Phi'Phi + s2 diag(1/k))^-1 Phi' y  and covariance  s2 (Phi'Phi + s2 diag(1/k))^-1."""
    pr = synth.dense_radio_problem("line_3D", m=24, seed=7, m_sim=200)
    model = oracle.DenseRadio2D(pr["NN"], pr["L"])
    T = pr["y"].shape[0]
    # exact odometry, negligible process noise -> the single particle follows the true path
    pos = pr["truth"]["pos"]
    odo = np.vstack([np.hstack([np.diff(pos.T, axis=0), np.zeros((T - 1, 1))]), np.zeros((1, 3))])
    st = oracle.Streams(np.full((1, T, 1), 0.5), np.zeros((1, T, 1, 1)))
    x0 = np.array([pos[0, 0], pos[1, 0], 0.0])
    out = oracle.particleFilter(model, odo, pr["y"], x0, pr["x0_lin"], pr["P0_lin"],
                                1e-300 * np.ones((1, 1)), pr["R"], 1, 1.0, st)
    xl_seq, P_seq = out[2], out[4]
    Phi = tools.eigenfun(pr["NN"], pos.T, pr["L"])
    s2 = pr["R"][0, 0]
    k = np.diag(pr["P0_lin"])
    A = Phi.T @ Phi + s2 * np.diag(1.0 / k)
    assert_close_norm(xl_seq, np.linalg.solve(A, Phi.T @ pr["y"][:, 0]), 1e-8)
    assert_close_norm(P_seq, s2 * np.linalg.inv(A), 1e-8)


def test_normalise_and_quirk_Q1():
    lw = np.array([-1000.0, -1001.0, -1000.0])
    w = normalise(lw)
    assert abs(w.sum() - 1) < 1e-12 and np.argmax(w) == 0   # lse loses ~|c|*eps
    pr = synth.dense_radio_problem("line_3D", m=10, seed=1, m_sim=100)
    model = oracle.DenseRadio2D(pr["NN"], pr["L"])
    N, T = 5, 4
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(0), 1, T, N, 1)
    taps = {}
    out = oracle.particleFilter(model, pr["odometry"][:T], pr["y"][:T], pr["x0_nonLin"], pr["x0_lin"],
                                pr["P0_lin"], pr["Q"], pr["R"], N, pr["dt"], st,
                                tap=lambda t, d: taps.__setitem__(t, d))
    last = taps[T - 1]
    dxl = out[3] - last["xl"][:, N - 1]
    # P_mean is ASSIGNED in the loop (src/particleFilter.m:228-230): last particle's term only
    assert_close_norm(out[5], last["w"][N - 1] * (last["P"][N - 1] + np.outer(dxl, dxl)), 1e-14)


# ---------------------------------------------------------------- smoothers
@pytest.mark.parametrize("fam", ["radio", "mag"])
def test_information_form_equals_covariance_form(fam):
    """src/particleSmootherInformationForm.m:34-37 claims identity with particleSmoother.m:
    the ancestor probabilities AI(:,t) and all outputs must agree (diagonal P0)."""
    if fam == "radio":
        pr = synth.dense_radio_problem("line_3D", m=20, seed=3, m_sim=200)
        model = oracle.DenseRadio2D(pr["NN"], pr["L"])
        N, K = 12, 3
    else:
        pr = synth.dense_mag_problem(N_T=10, m=20, seed=3, m_sim=100)
        model = oracle.DenseMag3D(pr["NN"], pr["L"])
        N, K = 8, 2
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(11), K, T, N, model.nz)
    args = (model, pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"],
            pr["R"], N, K, pr["dt"], st)
    ta, tb = {}, {}
    A = oracle.particleSmoother(*args, tap=lambda k, t, d: ta.__setitem__((k, t), d["paNt"]))
    B = oracle.particleSmootherInformationForm(*args, tap=lambda k, t, d: tb.__setitem__((k, t), d["paNt"]))
    n_checked = 0
    for key, pa in ta.items():
        if pa is not None:
            assert_close_norm(tb[key], pa, 1e-6, "paNt %s" % (key,))
            n_checked += 1
    assert n_checked == (K - 1) * (T - 1)
    for a, b in zip(A, B):
        assert_close_norm(b, a, 1e-9)


def test_smoother_first_sweep_is_a_plain_filter():
    """Sweep k=1 treats particle N_P like the others (src/particleSmoother.m:145-155)."""
    pr = synth.dense_radio_problem("line_3D", m=16, seed=5, m_sim=100)
    model = oracle.DenseRadio2D(pr["NN"], pr["L"])
    N = 9
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(2), 1, T, N, 1)
    args = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
    f = oracle.particleFilter(model, *args, N, pr["dt"], st, jitter=1e-2)
    XNK, XLK, PK = oracle.particleSmoother(model, *args, N, 1, pr["dt"], st)
    # the sampled trajectory is one of the filter's genealogies at time T
    ak = tools.sample(normalise(np.zeros(N)) * 0 + 1.0 / N, st.Uend[0])  # placeholder index range
    assert any(np.allclose(XNK[:, :, 0], f[7][:, i, :]) for i in range(N))
    assert 0 <= ak < N


def test_sparse_filter_runs_with_unobserved_steps():
    pr = synth.sparse_visual_problem(N_T=30, n_landmarks=6, N_P=7, seed=1, guess_map_var=0.01)
    pr["y"][3, :] = np.nan          # a step without any observation: empty-matrix algebra, logw = 0
    model = oracle.SparseVisual2D(6, *pr["camera"])
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(0), 1, 30, 7, 3)
    taps = {}
    oracle.particleFilter(model, pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"],
                          pr["Q"], pr["R"], 7, pr["dt"], st, tap=lambda t, d: taps.__setitem__(t, d))
    assert np.all(taps[3]["logw"] == 0.0)
    assert np.allclose(taps[3]["w"], 1.0 / 7)


# ---------------------------------------------------------------- Philox
def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    from oracle.streams import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        out = philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(o[0]) for o in out) == exp


def test_philox_streams_are_uniform_and_normal():
    U, Z = oracle.philox_uniforms_normals(42, 0, 3, 20000, 6)
    assert abs(U.mean() - 0.5) < 0.01 and U.min() >= 0 and U.max() < 1
    assert abs(Z.mean()) < 0.01 and abs(Z.std() - 1) < 0.01
    U2, Z2 = oracle.philox_uniforms_normals(42, 0, 3, 100, 6)
    assert np.array_equal(U[:100], U2) and np.array_equal(Z[:100], Z2)   # counter-based: N-invariant


# ---------------------------------------------------------------- the (f) rows: EKF baseline, localisation filter
def test_ekf_map_block_equals_single_particle_rbpf_when_the_pose_is_known():
    """Two independent restatements against each other: with a perfectly known pose (zero pose covariance,
    zero process noise) the EKF of ekf_dense.m:67-102 never moves its pose states, so its map block must be
    the Rao-Blackwellized filter's map update (src/particleFilter.m:184-198) of ONE particle that follows the
    same dead-reckoned poses."""
    from oracle.ekf import ekf_dense
    T, m = 12, 30
    pr = synth.dense_mag_problem(N_T=T, m=m, seed=4, m_sim=100)
    M = m + 3
    L = pr["L"]
    x0 = np.concatenate([pr["x0_nonLin"][:3], np.zeros(3), pr["x0_lin"].reshape(-1)])
    P0 = np.zeros((M + 6, M + 6))
    P0[6:, 6:] = pr["P0_lin"]
    Q0 = np.zeros((6, 6))
    xf, qn, Pf = ekf_dense(pr["NN"], L, np.vstack([-L, L]), pr["odometry"], pr["y"], x0, pr["x0_nonLin"][3:7], P0, Q0,
                           pr["R"], pr["dt"])
    om = oracle.DenseMag3D(pr["NN"], L)
    st = oracle.Streams(np.zeros((1, T, 1)), np.zeros((1, T, 1, 6)))
    Qeps = 1e-300 * np.eye(6)          # chol(dt*Q) must exist; the noise it scales is zero anyway
    out = oracle.particleFilter(om, pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], Qeps, pr["R"],
                                1, pr["dt"], st)
    traj_max, xl_max, P_max = out[0], out[2], out[4]
    assert_close_norm(xf[:3], traj_max[:3], 1e-12, "EKF pose == dead-reckoned pose")
    assert_close_norm(qn, traj_max[3:7], 1e-12, "linearisation point == dead-reckoned orientation")
    assert_close_norm(xf[6:, -1], xl_max, 1e-9, "map mean")
    assert_close_norm(Pf[6:, 6:], P_max, 1e-9, "map covariance")
    assert np.max(np.abs(Pf[:6, :])) == 0.0


def test_localization_oracle_pieces():
    """run_localization.m:241-280 restated: qRight(q)*p = p (x) q, the weight of one particle from the
    normal densities written out, and weights that prefer the particle sitting on the true pose."""
    from oracle.localization import qRight, dynModel_loc, measModel_loc
    rng = np.random.default_rng(2)
    q, p = rng.standard_normal(4), rng.standard_normal(4)
    assert_close_norm(qRight(q) @ p, tools.qLeft(p) @ q, 1e-14, "qRight")
    Q = np.diag([1e-2, 2e-2, 3e-2, 1e-4, 2e-4, 3e-4])
    xn = np.concatenate([rng.standard_normal(3), q / np.linalg.norm(q)])
    dx = np.concatenate([0.1 * rng.standard_normal(3), p / np.linalg.norm(p)])
    z = rng.standard_normal(6)
    out = dynModel_loc(xn, dx, 0.5, Q, z)
    assert_close_norm(out[:3], xn[:3] + dx[:3] + np.sqrt(0.5 * np.diag(Q)[:3]) * z[:3], 1e-14, "position")
    assert_close_norm(out[3:], tools.qLeft(tools.qLeft(dx[3:]) @ xn[3:]) @ tools.expq(np.sqrt(0.5 * np.diag(Q)[3:]) * z[3:]),
                      1e-14, "orientation")
    pr = synth.dense_mag_problem(N_T=4, m=40, seed=6, m_sim=100)
    foo = rng.standard_normal(43)
    om = oracle.DenseMag3D(pr["NN"], pr["L"])
    x_true = pr["x0_nonLin"]
    y = om.measModel(x_true[:, None])[0] @ foo
    xs = np.stack([x_true, x_true + np.array([1.5, -1.0, 0.2, 0, 0, 0, 0])], axis=1)
    dVar = np.full((2, 3), 0.3)
    w = measModel_loc(y, xs, pr["NN"], pr["L"], foo, dVar, 0.1)
    sd = np.sqrt(0.4)
    assert abs(w[0] - 3.0 / (sd * np.sqrt(2 * np.pi))) < 1e-12       # zero residual: three density peaks added up
    assert w[0] > w[1] > 0


# ---------------------------------------------------------------------------
# the error bound behind the device's parallel resampling path (csrc/step_kernels.cuh: k_scan_approx,
# k_search_checked): a NumPy restatement of the blocked summation order, against the strict
# left-to-right cumsum (np.cumsum: tools/sample.m:30)
# ---------------------------------------------------------------------------
def _scan_approx_numpy(w):
    """k_scan_approx, operation by operation: 1024 threads with contiguous segments (sequential sums), a
    Hillis-Steele inclusive scan inside each warp (shfl_up by 1, 2, 4, 8, 16), the same over the 32 warp
    totals, then every thread walks its segment from its exclusive prefix."""
    N = len(w)
    per = (N + 1023) // 1024
    seg = [w[min(N, t * per):min(N, (t + 1) * per)] for t in range(1024)]
    s = np.array([np.cumsum(x)[-1] if len(x) else 0.0 for x in seg])     # sequential within a thread

    def warp_scan(x):                      # x: [n_warps, 32]
        x = x.copy()
        for o in (1, 2, 4, 8, 16):
            y = np.zeros_like(x)
            y[:, o:] = x[:, :-o]
            x[:, o:] = x[:, o:] + y[:, o:]
        return x
    x = warp_scan(s.reshape(32, 32))
    tot = warp_scan(x[:, 31].reshape(1, 32))[0]
    excl = np.zeros_like(x)
    excl[:, 1:] = x[:, :-1]
    base = np.concatenate([[0.0], tot[:-1]])[:, None] + excl           # exclusive prefix of every thread
    out = np.empty(N)
    for t in range(1024):
        b, e = min(N, t * per), min(N, (t + 1) * per)
        if e > b:
            run = base[t // 32, t % 32]
            for i in range(b, e):
                run = run + w[i]
                out[i] = run
    return out


@pytest.mark.parametrize("N", [4096, 10000, 80000])
@pytest.mark.parametrize("kind", ["uniform", "heavy", "zeros", "dominant", "ascending", "descending"])
def test_parallel_scan_stays_inside_the_bound_the_device_uses(N, kind):
    """|blocked prefix sum - sequential cumsum| <= delta(j) for every j, with the delta of k_search_checked:
    1.05 eps ((j+2) wc'(j) + (2 ceil(N/1024) + 16) sum(w)).  That inequality is what lets a draw prove that
    count(wc < u) does not depend on the rounding order (wc'(idx-1) + delta < u <= wc'(idx) - delta)."""
    rng = np.random.default_rng(N + len(kind))
    w = rng.random(N)
    if kind == "heavy":
        w = w ** 12
    elif kind == "zeros":
        w[rng.random(N) < 0.7] = 0.0
    elif kind == "dominant":
        w = w * 1e-9
        w[N // 3] = 1.0
    elif kind == "ascending":
        w = np.sort(w ** 4)
    elif kind == "descending":
        w = np.sort(w ** 4)[::-1].copy()
    w = w / w.sum()
    seq = np.cumsum(w)                      # strict left-to-right: the contract
    par = _scan_approx_numpy(w)
    eps = 2.0 ** -53
    per = (N + 1023) // 1024
    j = np.arange(N)
    delta = 1.05 * eps * ((j + 2) * par + (2 * per + 16) * par[-1])
    assert np.all(np.abs(par - seq) <= delta)
    # the decision rule is not vacuous: for random draws almost every one proves itself
    u = rng.random(2000)
    idx = np.searchsorted(par, u, side="left")               # count(par < u)
    jj = np.minimum(idx, N - 1)
    d = 1.05 * eps * ((idx + 2) * par[jj] + (2 * per + 16) * par[-1])
    below = (idx == 0) | (par[np.maximum(idx - 1, 0)] + d < u)
    above = (idx == N) | (par[jj] - d >= u)
    sure = below & above
    assert sure.mean() > 0.99
    assert np.array_equal(idx[sure], np.searchsorted(seq, u, side="left")[sure])
