"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200")]
import numpy as np
import rbslam

what = sys.argv[1] if len(sys.argv) > 1 else "all"
pr = rbslam.synth.dense_mag_problem(N_T=4, m=125, seed=3, m_sim=200)     # M = 128
gm = rbslam.models.from_problem(pr)
a = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
if what in ("all", "packed"):
    with rbslam.Context(gm, 24, 4, rng_mode=1, seed=1, kalman_variant=7) as ctx:
        o = ctx.filter_run(*a, pr["dt"])
    print("packed ok", float(o["xl_mean"][0]))
if what in ("all", "full"):
    with rbslam.Context(gm, 24, 4, rng_mode=1, seed=1, kalman_variant=2) as ctx:
        o = ctx.filter_run(*a, pr["dt"])
    print("full ok", float(o["xl_mean"][0]))
if what in ("all", "group"):
    for v in (2, 7):
        with rbslam.Context(gm, 24, 4, rng_mode=1, seed=1, kalman_variant=v, devices=[0, 0]) as ctx:
            o = ctx.filter_run(*a, pr["dt"])
        print("group ok", v, float(o["xl_mean"][0]))
if what in ("all", "smoother"):
    for form in (0, 1):
        with rbslam.Context(gm, 8, 4, rng_mode=1, seed=1, information_form=(form == 1)) as ctx:
            o = ctx.smoother_run(*a, pr["dt"], 2, form)
        print("smoother ok", form, float(o["XLK"][0, -1]))
if what in ("all", "ekf"):
    M = gm.M
    x0 = np.concatenate([pr["x0_nonLin"][:3], np.zeros(3), pr["x0_lin"].reshape(-1)])
    P0 = np.zeros((M + 6, M + 6)); P0[6:, 6:] = pr["P0_lin"]
    xf, qn, Pl = rbslam.ekf_dense(gm, pr["odometry"], pr["y"], x0, pr["x0_nonLin"][3:], P0, pr["Q"], pr["R"], pr["dt"])
    print("ekf ok", float(xf[0, -1]))
if what in ("all", "loc"):
    rng = np.random.default_rng(0)
    tm, tmean = rbslam.particleFilterLocalization(gm.dynModel, gm.measModel, pr["odometry"], pr["y"], pr["x0_nonLin"], pr["Q"],
                                                  pr["R"], 40, pr["dt"], map_mean=rng.standard_normal(gm.M),
                                                  var_rows=0.5 + rng.random((40, 3)), sigma2=0.1, rng=2)
    print("loc ok", float(tmean[0, -1]))
