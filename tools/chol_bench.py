"""K7 alone: batched 515 x 515 Cholesky + solve through rbslam_smoother_run's ancestor phase
(information form, C5 shape N = 4096, T-slice), TFLOP/s against the measured fp64 peak."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200")]
import numpy as np
import rbslam

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Ts = int(sys.argv[2]) if len(sys.argv) > 2 else 12
pr = rbslam.synth.dense_mag_problem(N_T=Ts, m=512, seed=1, n_laps=3, m_sim=2000)
gm = rbslam.models.from_problem(pr)
a = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
with rbslam.Context(gm, N, Ts, rng_mode=1, seed=1, information_form=True) as ctx:
    ctx.smoother_run(*a, pr["dt"], 2, 1)
    ctx.phase_timing(True)
    t0 = time.perf_counter()
    ctx.smoother_run(*a, pr["dt"], 2, 1)
    wall = time.perf_counter() - t0
    ph = ctx.phase_times()
M = gm.M
anc = ph["ancestor"] / (Ts - 1)
fl = N * (M ** 3 / 3.0 + 4.0 * M * M)
print(json.dumps({"N": N, "M": M, "kernel": os.environ.get("RBSLAM_CHOL_KERNEL", "inv"), "threads": os.environ.get("RBSLAM_CHOL_THREADS", "128"), "ancestor_ms_per_step": anc,
                  "tflops": fl / anc / 1e9, "frac_of_37.09": fl / anc / 1e9 / 37.09, "wall_s": wall,
                  "phases_ms": {k: round(v, 2) for k, v in ph.items() if v > 0}}))
