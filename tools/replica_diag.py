"""Where does a sweep of the replicated smoother spend its time?  C5 slice on 1 / 2 / 4 / ... GPUs:
wall time per step of sweep 2 without phase timing, then the leader's phase times."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200")]
import rbslam

worlds = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2]
Ts = int(sys.argv[2]) if len(sys.argv) > 2 else 24
N5 = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
pr = rbslam.synth.dense_mag_problem(N_T=Ts, m=512, seed=1, n_laps=3, m_sim=2000)
gm = rbslam.models.from_problem(pr)
a = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
for world in worlds:
    kw = dict(replicas=True, devices=list(range(world))) if world > 1 else {}
    with rbslam.Context(gm, N5, Ts, rng_mode=1, seed=1, information_form=True, **kw) as ctx:
        ctx.smoother_run(*a, pr["dt"], 2, 1)
        t = []
        for nk in (1, 2, 3):
            t0 = time.perf_counter()
            ctx.smoother_run(*a, pr["dt"], nk, 1)
            t.append(time.perf_counter() - t0)
        ctx.phase_timing(True)
        t0 = time.perf_counter()
        ctx.smoother_run(*a, pr["dt"], 2, 1)
        tp = time.perf_counter() - t0
        ph = ctx.phase_times()
    print(json.dumps({"world": world, "T": Ts, "N": N5, "wall_s_nk123": [round(x, 4) for x in t],
                      "ms_per_step_sweep2": round(1e3 * (t[1] - t[0]) / Ts, 3),
                      "ms_per_step_sweep3": round(1e3 * (t[2] - t[1]) / Ts, 3),
                      "wall_s_nk2_with_phase_timing": round(tp, 4),
                      "leader_phases_ms_total": {k: round(v, 2) for k, v in ph.items() if v > 0}}), flush=True)
