"""Edge cases of the filter entry point through the C ABI: degenerate sizes, argument forms the
reference accepts (src/particleFilter.m:60-64,75-82), and argument errors."""
import numpy as np
import pytest

import oracle
from conftest import assert_close_norm
from test_gpu_filter import _setup, _args, _run_oracle, _compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fam,kw", [("radio", {"m": 40}), ("mag", {"m": 64, "T": 6}), ("mag", {"m": 253, "T": 5})])
def test_single_particle(rbslam_lib, fam, kw):
    """N_P = 1: every draw returns particle 1, the weight is 1, the filter is a plain EKF-like
    recursion along one sampled path."""
    rb = rbslam_lib
    pr, om, gm = _setup(rb, fam, 1, **kw)
    T = pr["y"].shape[0]
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(11), 1, T, 1, om.nz)
    ref, taps = _run_oracle(om, pr, 1, st)
    with rb.Context(gm, 1, T, rng_mode=0) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    assert np.all(o["ancestors"][:, 1:] == 0) and np.all(o["w_hist"] == 1.0)
    _compare(o, ref, taps, T)


@pytest.mark.parametrize("fam,kw", [("radio", {"m": 40}), ("sparse", {"T": 5})])
def test_single_time_step(rbslam_lib, fam, kw):
    """N_T = 1: no resampling, no propagation (src/particleFilter.m:103), one measurement update."""
    rb = rbslam_lib
    N = 12
    pr, om, gm = _setup(rb, fam, N, **kw)
    pr["y"] = pr["y"][:1].copy()
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(12), 1, 1, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    with rb.Context(gm, N, 1, rng_mode=0) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    _compare(o, ref, taps, 1)


def test_dt_vector_and_per_particle_x0_lin(rbslam_lib):
    """dt as a (T-1)-vector (src/particleFilter.m:80-82) and x0_lin as [M x N_P] (:60-64)."""
    rb = rbslam_lib
    N = 16
    pr, om, gm = _setup(rb, "mag", N, m=64, T=8)
    T = pr["y"].shape[0]
    rng = np.random.default_rng(13)
    pr["dt"] = pr["dt"] * (0.5 + rng.random(T - 1))
    pr["x0_lin"] = np.repeat(np.asarray(pr["x0_lin"]).reshape(-1, 1), N, axis=1) + 0.01 * rng.standard_normal((gm.M, N))
    st = oracle.Streams.from_numpy_rng(np.random.default_rng(14), 1, T, N, om.nz)
    ref, taps = _run_oracle(om, pr, N, st)
    with rb.Context(gm, N, T, rng_mode=0) as ctx:
        o = ctx.filter_run(*_args(pr), pr["dt"], streams=st, taps=True)
    _compare(o, ref, taps, T)


def test_argument_errors(rbslam_lib):
    rb = rbslam_lib
    N = 8
    pr, om, gm = _setup(rb, "radio", N, m=40)
    T = pr["y"].shape[0]
    with pytest.raises(rb.RbslamError):            # more steps than the context was created for
        with rb.Context(gm, N, T - 1, rng_mode=1, seed=1) as ctx:
            ctx.filter_run(*_args(pr), pr["dt"])
    with pytest.raises((rb.RbslamError, ValueError)):   # P0_lin of the wrong size
        with rb.Context(gm, N, T, rng_mode=1, seed=1) as ctx:
            a = list(_args(pr))
            a[4] = np.eye(gm.M + 1)
            ctx.filter_run(*a, pr["dt"])
    with pytest.raises((rb.RbslamError, ValueError)):   # injected-stream mode without streams
        with rb.Context(gm, N, T, rng_mode=0) as ctx:
            ctx.filter_run(*_args(pr), pr["dt"])
    with pytest.raises(rb.RbslamError):            # zero particles
        rb.Context(gm, 0, T)
