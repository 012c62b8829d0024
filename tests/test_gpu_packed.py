"""kalman_variant 7 (and the filter-only auto mode -1): packed symmetric tile slabs streamed by
k_stream_fam_pt on the fp64 tensor cores, against the oracle.  Same bars as the full-storage
kernels: ancestor indices bit-exact, means / covariances / log-weights within 1e-8 (norm-wise
for matrices).  The reference never symmetrises P (src/particleFilter.m:198), so its two
triangles differ by rounding; the tolerance absorbs that."""
import numpy as np
import pytest

import oracle
from conftest import assert_close_norm
from test_gpu_filter import _setup, _args, _run_oracle, _compare, TOL
from test_gpu_kernels import _problem, _rand_spd, _oracle_update

pytestmark = pytest.mark.gpu
PT = 7


@pytest.mark.parametrize("fam,m", [("mag", 13), ("mag", 64), ("mag", 253), ("mag", 512), ("mag", 1024),
                                   ("radio", 300)])
def test_packed_kalman_update(rbslam_lib, fam, m):
    """One update from a zero pending pair: packing, both tensor-core products, the column-side
    reduction, the fold in k_innov4, mirror on read-out (rbslam_op_kalman_update flushes the
    deferred downdate and unpacks the slabs)."""
    rb = rbslam_lib
    pr, om, gm = _problem(rb, fam, m=m)
    N = 12
    M = gm.M
    rng = np.random.default_rng(m)
    xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1) + 0.2 * rng.standard_normal((gm.n, N))
    H = om.measModel(xn)
    P = np.stack([_rand_spd(rng, M, 10.0) for _ in range(N)])
    xl = rng.standard_normal((M, N))
    yt, R = pr["y"][3], pr["R"]
    with rb.Context(gm, N, 4, kalman_variant=PT) as ctx:
        xl2, P2, logw = ctx.op_kalman_update(xl, P.transpose(1, 2, 0), yt, R, 1e-3, H=H)
    for i in range(N):
        xr, Pr, lr = _oracle_update(xl[:, i], P[i], H[i], yt, R, 1e-3)
        assert_close_norm(xl2[:, i], xr, 1e-8, "xl")
        assert_close_norm(P2[:, :, i], Pr, 1e-8, "P")
        assert abs(logw[i] - lr) <= 1e-8 * max(1.0, abs(lr))


@pytest.mark.parametrize("cfg", ["48,4", "7,3", "100,2"])
def test_packed_stage_shapes(rbslam_lib, cfg, monkeypatch):
    """Stage size / ring depth do not change results: stages that end inside a panel, stages
    that span many of the short panels at the narrow end, a ring of two."""
    rb = rbslam_lib
    monkeypatch.setenv("RBSLAM_PT_CFG", cfg)
    pr, om, gm = _setup(rb, "mag", 12, m=253, T=6)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(5), 1, T, 12, om.nz)
    ref, taps = _run_oracle(om, pr, 12, st)
    with rb.Context(gm, 12, T, rng_mode=0, kalman_variant=PT) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    _compare(o, ref, taps, T)


@pytest.mark.parametrize("fam,N,kw", [
    ("mag", 16, {"m": 253, "T": 12}),      # M=256: nb = 32 row blocks
    ("mag", 100, {"m": 512, "T": 6}),      # C1 shape, M=515
    ("mag", 40, {"m": 1024, "T": 5}),      # C4 slab size, M=1027: nine row blocks per warp
    ("radio", 64, {"m": 300}),             # d=1
    ("mag", 24, {"m": 64, "T": 24}),       # small M through the streaming path
])
def test_packed_filter_teacher_forced(rbslam_lib, fam, N, kw):
    """Ancestors from the oracle run (families of 1..many siblings, surplus families, copies
    and in-place offspring): every step's logw and all 8 outputs."""
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(5), 1, T, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    forced = np.stack([tp["ai"] for tp in taps])[None].astype(np.int32)
    with rb.Context(gm, N, T, rng_mode=0, kalman_variant=PT) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, forced_ancestors=forced, taps=True)
    _compare(o, ref, taps, T)


def test_packed_filter_free_running_and_read_particles(rbslam_lib):
    """Device draws its own ancestors (bit-exact), and the per-step tap returns full
    covariances although only the block triangle is stored."""
    rb = rbslam_lib
    N = 24
    pr, om, gm = _setup(rb, "mag", N, m=253, T=8)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(6), 1, T, N, om.nz)
    states = {}
    taps = []

    def tap(t, d):
        taps.append(dict(logw=d["logw"], w=d["w"], ai=d["ai"], xl=d["xl"]))
        states[t] = dict(xl=d["xl"].copy(), P=np.array(d["P"]))
    ref = oracle.particleFilter(om, *_args(pr), N, pr["dt"], st, tap=tap)
    seen = []
    with rb.Context(gm, N, T, rng_mode=0, kalman_variant=PT) as ctx:
        def cb(k, t):
            s = ctx.read_particles()
            assert_close_norm(s["xl"], states[t]["xl"], TOL, "xl@%d" % t)
            assert_close_norm(s["P"], states[t]["P"].transpose(1, 2, 0), TOL, "P@%d" % t)
            seen.append(t)
        ctx.set_step_callback(cb)
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    assert seen == list(range(T))
    _compare(o, ref, taps, T)


def test_packed_is_filter_only_and_auto_mode(rbslam_lib):
    rb = rbslam_lib
    pr, om, gm = _setup(rb, "mag", 8, m=253, T=4)
    with pytest.raises(rb.RbslamError):
        rb.Context(gm, 8, 4, kalman_variant=PT, information_form=True)
    with rb.Context(gm, 8, 4, rng_mode=1, seed=1, kalman_variant=PT) as ctx:
        with pytest.raises(rb.RbslamError):
            ctx.smoother_run(*_args(pr), pr["dt"], 2)
    # -1 = "auto, filter entry points only": what the particleFilter drop-in passes
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(9), 1, 4, 8, om.nz)
    outs = rb.particleFilter(gm.dynModel, gm.measModel, pr["odometry"], pr["y"], pr["x0_nonLin"],
                             pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"], 8, pr["dt"], False, None, rng=st)
    ref = oracle.particleFilter(om, *_args(pr), 8, pr["dt"], st)
    for a, b in zip(outs, ref):
        assert_close_norm(a, b, 1e-7)


def test_packed_filter_is_bit_reproducible(rbslam_lib):
    """The column-side tiles of the packed kernel travel from 15 consumer warps to the reducer lanes
    through mbarrier-ordered shared-memory buffers and are added in fixed warp order: three runs on
    the same inputs must agree in every bit (a stale or early read of a hand-off buffer would not),
    with many more work items than SMs so that the schedule differs from run to run."""
    rb = rbslam_lib
    N, T = 600, 5
    pr, om, gm = _setup(rb, "mag", N, m=253, T=T)
    runs = []
    for _ in range(3):
        with rb.Context(gm, N, T, rng_mode=1, seed=11, kalman_variant=PT) as ctx:
            runs.append(ctx.filter_run(*_args(pr), pr["dt"]))
    for o in runs[1:]:
        for k in ("traj_max", "traj_mean", "xl_max", "xl_mean", "P_max", "P_mean"):
            assert np.array_equal(o[k], runs[0][k]), k
