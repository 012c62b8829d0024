"""Where does rbslam_filter_run spend its time?  begin / steps / end timed separately."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200")]
import numpy as np
import rbslam
from rbslam import _capi
from bench import make_problem

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
T = int(sys.argv[2]) if len(sys.argv) > 2 else 24
pr, T = make_problem(1024, T)
gm = rbslam.models.from_problem(pr)
fargs = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
t0 = time.perf_counter()
ctx = rbslam.Context(gm, 10000, T, rng_mode=_capi.RNG_PHILOX, seed=1, keep_history=True, kalman_variant=variant)
print("create %.3f s" % (time.perf_counter() - t0))
for rep in range(2):
    t0 = time.perf_counter(); ctx.filter_begin(*fargs, pr["dt"]); t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
    for _ in range(T): ctx.filter_step()
    t3 = time.perf_counter(); ctx.sync(); t4 = time.perf_counter()
    o = ctx.filter_end(T=None); t5 = time.perf_counter()
    print("rep %d: begin %.3f (+sync %.3f)  steps enqueue %.3f (+sync %.3f)  end %.3f  total %.3f" %
          (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0))
    t0 = time.perf_counter(); ctx.filter_run(*fargs, pr["dt"], want_xn_traj=False); print("filter_run %.3f" % (time.perf_counter() - t0))
