#!/usr/bin/env python
"""Smoother seconds per run (second half of BASELINE.json's metric) at the reference's
example scale C1: dense-mag, N_P = 100, m = 512 (M = 515), T = 192 (run_dense3D_magfield.m:85,134),
covariance form and information form, plus the CPU oracle on a bounded sample.

    python tools/bench_smoother.py [--NK 3] [--N 100] [--m 512] [--T 192] [--cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--NK", type=int, default=3)
    ap.add_argument("--N", type=int, default=100)
    ap.add_argument("--m", type=int, default=512)
    ap.add_argument("--T", type=int, default=192)
    ap.add_argument("--cpu", action="store_true", help="also time the oracle on a bounded sample")
    ap.add_argument("--forms", default="0,1", help="comma list: 0 covariance form, 1 information form")
    ap.add_argument("--laps", type=int, default=3)
    a = ap.parse_args()
    import rbslam
    pr = rbslam.synth.dense_mag_problem(N_T=a.T, m=a.m, seed=1, n_laps=a.laps, m_sim=2000)
    gm = rbslam.models.from_problem(pr)
    args = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
    out = {"config": {"workload": "C1 dense-mag smoother", "N_P": a.N, "M": gm.M, "T": a.T, "N_K": a.NK}}
    forms = [int(f) for f in a.forms.split(",")]
    for form, name in ((0, "covariance_form"), (1, "information_form")):
        if form not in forms:
            continue
        with rbslam.Context(gm, a.N, a.T, rng_mode=1, seed=1, information_form=(form == 1)) as ctx:
            ctx.smoother_run(*args, pr["dt"], 1, form)          # warm-up (one plain-filter sweep)
            ctx.phase_timing(True)
            t0 = time.perf_counter()
            o = ctx.smoother_run(*args, pr["dt"], a.NK, form)
            secs = time.perf_counter() - t0
            ph = ctx.phase_times()
        rmse = float(np.sqrt(np.mean((o["XNK"][:3, :, -1] - pr["truth"]["pos"]) ** 2)))
        out[name] = {"seconds_per_run": secs, "seconds_per_sweep": secs / a.NK,
                     "phases_ms": {k: v for k, v in ph.items() if v > 0},
                     "rmse_pos_last_sweep": rmse}
    if a.cpu:
        import oracle
        Nc, Tc = 8, 24
        om = oracle.DenseMag3D(pr["NN"], pr["L"])
        st = oracle.Streams.from_numpy_rng(np.random.default_rng(0), 2, Tc, Nc, om.nz)
        cargs = (pr["odometry"][:Tc], pr["y"][:Tc], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
        t0 = time.perf_counter()
        oracle.particleSmoother(om, *cargs, Nc, 2, pr["dt"], st)
        out["cpu_oracle_covariance_form"] = {"seconds": time.perf_counter() - t0,
                                             "sample": "N_P=%d, T=%d, N_K=2, M=%d, single thread" % (Nc, Tc, gm.M)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
