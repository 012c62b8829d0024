"""K6 (covariance-form ancestor weights) at the C1 shape: wall time of a smoother run and the ancestor phase.
Under `ncu --metrics gpu__time_duration.sum -k regex:k_dgemm` the launch list gives TFLOP/s per launch
(tools/gemm_tflops.py)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200")]
import rbslam

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T = int(sys.argv[2]) if len(sys.argv) > 2 else 192
NK = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pr = rbslam.synth.dense_mag_problem(N_T=T, m=512, seed=1, n_laps=3, m_sim=2000)
gm = rbslam.models.from_problem(pr)
a = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
with rbslam.Context(gm, N, T, rng_mode=1, seed=1) as ctx:
    ctx.smoother_run(*a, pr["dt"], 2, 0)
    ctx.phase_timing(True)
    t0 = time.perf_counter()
    ctx.smoother_run(*a, pr["dt"], NK, 0)
    wall = time.perf_counter() - t0
    ph = ctx.phase_times()
M, d = gm.M, 3
fl = sum(N * (2.0 * M * M * 3 * (T - t) + 2.0 * M * (3 * (T - t)) ** 2) for t in range(1, T)) * (NK - 1)
print(json.dumps({"N": N, "M": M, "T": T, "N_K": NK, "gemm": "sync" if os.environ.get("RBSLAM_GEMM_SYNC") else "pipe",
                  "wall_s": round(wall, 4), "ancestor_ms_total": round(ph["ancestor"], 1),
                  "gemm_flops_T": round(fl / 1e12, 3),
                  "ancestor_phase_tflops_lower_bound": round(fl / ph["ancestor"] / 1e9, 2),
                  "phases_ms": {k: round(v, 1) for k, v in ph.items() if v > 0}}))
