"""Kernel times of the two resampling paths at N weights (run under ncu --metrics gpu__time_duration.sum)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200")]
import numpy as np
import rbslam
N = int(sys.argv[1]) if len(sys.argv) > 1 else 80000
pr = rbslam.synth.dense_radio_problem(m=16, seed=1)
gm = rbslam.models.from_problem(pr)
rng = np.random.default_rng(0)
w = rng.random(N) ** 4; w /= w.sum()
with rbslam.Context(gm, 8, 4) as ctx:
    for rep in range(6):
        ctx.op_resample(w, rng.random(N))
    print(ctx.status_counters())
