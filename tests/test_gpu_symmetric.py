"""kalman_variant 4 / 5: the symmetric streaming passes (lower triangle of every covariance slab
only; k_stream_fam_sym in SIMT form, k_stream_fam_symt on the fp64 tensor cores) against the oracle.  Same bars as the full-storage kernels: ancestor
indices bit-exact, means / covariances / log-weights within 1e-8 (norm-wise for matrices).
The reference never symmetrises P (src/particleFilter.m:198), so its two triangles differ by
rounding; the tolerance absorbs that."""
import os

import numpy as np
import pytest

import oracle
from conftest import assert_close_norm
from test_gpu_filter import _setup, _args, _run_oracle, _compare, TOL
from test_gpu_kernels import _problem, _rand_spd, _oracle_update

pytestmark = pytest.mark.gpu


# 4 = k_stream_fam_sym (SIMT, shuffle reduction), 5 = k_stream_fam_symt (fp64 tensor cores): both
# verified on a B200.  6 = k_stream_fam_symp (variant 5 with a producer warp and a six-slot ring) was
# written after the round's GPU budget was spent and has NOT run yet: its cases are collected only
# with RBSLAM_TEST_UNVERIFIED=1 so that the default suite contains verified kernels only.
_VARIANTS = [(4, "simt"), (5, "dmma")] + ([(6, "dmma-pipelined")] if os.environ.get("RBSLAM_TEST_UNVERIFIED") else [])


@pytest.fixture(params=[v for v, _ in _VARIANTS], ids=[n for _, n in _VARIANTS])
def SYM(request):
    return request.param


@pytest.mark.parametrize("fam,m", [("mag", 64), ("mag", 253), ("mag", 512), ("mag", 1024), ("radio", 300)])
def test_sym_kalman_update(rbslam_lib, SYM, fam, m):
    """One update from a zero pending pair: triangle streaming, column-side reduction, mirror on
    read-out (rbslam_op_kalman_update flushes the deferred downdate and packs the slabs)."""
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
    with rb.Context(gm, N, 4, kalman_variant=SYM) as ctx:
        xl2, P2, logw = ctx.op_kalman_update(xl, P.transpose(1, 2, 0), yt, R, 1e-3, H=H)
    for i in range(N):
        xr, Pr, lr = _oracle_update(xl[:, i], P[i], H[i], yt, R, 1e-3)
        assert_close_norm(xl2[:, i], xr, 1e-8, "xl")
        assert_close_norm(P2[:, :, i], Pr, 1e-8, "P")
        assert np.array_equal(P2[:, :, i], P2[:, :, i].T)       # read-out mirrors the triangle
        assert abs(logw[i] - lr) <= 1e-8 * max(1.0, abs(lr))


@pytest.mark.parametrize("fam,N,kw", [
    ("mag", 16, {"m": 253, "T": 12}),      # M=256: one row pair per thread, column splits
    ("mag", 100, {"m": 512, "T": 6}),      # C1 shape, M=515
    ("mag", 40, {"m": 1024, "T": 5}),      # C4 slab size, M=1027: two row pairs per thread
    ("radio", 64, {"m": 300}),             # d=1
])
def test_sym_filter_teacher_forced(rbslam_lib, SYM, fam, N, kw):
    """Ancestors from the oracle run (families of 1..many siblings, surplus families, copies
    and in-place offspring): every step's logw and all 8 outputs."""
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, N, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(5), 1, T, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    forced = np.stack([tp["ai"] for tp in taps])[None].astype(np.int32)
    with rb.Context(gm, N, T, rng_mode=0, kalman_variant=SYM) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, forced_ancestors=forced, taps=True)
    _compare(o, ref, taps, T)


def test_sym_filter_free_running_and_read_particles(rbslam_lib, SYM):
    """Device draws its own ancestors (bit-exact), and the per-step tap returns full symmetric
    covariances although only the triangle is maintained."""
    rb = rbslam_lib
    N = 24
    pr, om, gm = _setup(rb, "mag", N, m=253, T=8)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(6), 1, T, N, om.nz)
    states = {}
    ref, taps = None, []

    def tap(t, d):
        taps.append(dict(logw=d["logw"], w=d["w"], ai=d["ai"], xl=d["xl"]))
        states[t] = dict(xl=d["xl"].copy(), P=np.array(d["P"]))
    ref = oracle.particleFilter(om, *_args(pr), N, pr["dt"], st, tap=tap)
    seen = []
    with rb.Context(gm, N, T, rng_mode=0, kalman_variant=SYM) as ctx:
        def cb(k, t):
            s = ctx.read_particles()
            assert_close_norm(s["xl"], states[t]["xl"], TOL, "xl@%d" % t)
            assert_close_norm(s["P"], states[t]["P"].transpose(1, 2, 0), TOL, "P@%d" % t)
            seen.append(t)
        ctx.set_step_callback(cb)
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    assert seen == list(range(T))
    _compare(o, ref, taps, T)


def test_sym_is_filter_only(rbslam_lib, SYM):
    rb = rbslam_lib
    pr, om, gm = _setup(rb, "mag", 8, m=253, T=4)
    with pytest.raises(rb.RbslamError):
        rb.Context(gm, 8, 4, kalman_variant=SYM, information_form=True)
    with rb.Context(gm, 8, 4, rng_mode=1, seed=1, kalman_variant=SYM) as ctx:
        with pytest.raises(rb.RbslamError):
            ctx.smoother_run(*_args(pr), pr["dt"], 2)
