import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/rao-blackwellized-slam-smoothing_b200')
import numpy as np, rbslam, oracle
pr = rbslam.synth.dense_mag_problem(N_T=24, m=64, seed=3, m_sim=300)
gm = rbslam.models.from_problem(pr); om = oracle.DenseMag3D(pr["NN"], pr["L"])
N=4
xn = np.repeat(pr["x0_nonLin"][:, None], N, axis=1)
xk = pr["x0_nonLin"].copy()
Q = pr["Q"]; dx = np.array([0,0,0,1.0,0,0,0])
with rbslam.Context(gm, N, 4) as ctx:
    print('same', ctx.op_dyn_logweight(xk, xn, dx, 0.01, Q))
    xk2 = xk.copy(); xk2[0]+=0.01
    print('pos', ctx.op_dyn_logweight(xk2, xn, dx, 0.01, Q), -0.5*(om.dynResNorm(xk2, xn[:,0], dx, 0.01, Q)**2).sum())
    xk3 = xk.copy(); xk3[4]+=0.001
    print('quat', ctx.op_dyn_logweight(xk3, xn, dx, 0.01, Q), -0.5*(om.dynResNorm(xk3, xn[:,0], dx, 0.01, Q)**2).sum())
    print('default', ctx.op_dyn_logweight(xk3, xn, dx, 0.01, Q, use_default=True))
    print(Q, dx, xk, xn[:,0])
    one = np.array([0,0,0,1.0,0,0,0]); 
    xn1 = np.repeat(one[:,None], N, axis=1)
    print('ident', ctx.op_dyn_logweight(one, xn1, dx, 0.01, Q))
    Q2 = np.diag([0.25,0.25,0.01,3e-8,3e-8,2.7e-5])
    print('ident Q2', ctx.op_dyn_logweight(one, xn1, dx, 0.01, Q2))
