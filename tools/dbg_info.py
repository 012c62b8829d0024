import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/rao-blackwellized-slam-smoothing_b200')
import numpy as np, rbslam as rb, oracle
pr = rb.synth.dense_radio_problem("line_3D", m=40, seed=2, m_sim=400)
om = oracle.DenseRadio2D(pr["NN"], pr["L"]); gm = rb.models.from_problem(pr)
N,K=20,2; T=pr["y"].shape[0]
st = oracle.Streams.from_numpy_rng(np.random.default_rng(9), K, T, N, om.nz)
args=(pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
rec={}; taps={}
oracle.particleSmootherInformationForm(om, *args, N, K, pr["dt"], st, record=rec, tap=lambda k,t,d: taps.__setitem__((k,t), dict(logw=d['logw'].copy(), hld=d['halfLogDetP'].copy(), ivec=d['ivec'].copy(), Imat=np.array(d['Imat']), xl=d['xl'].copy(), P=np.array(d['P']))))
ai=np.zeros((K,T,N),dtype=np.int32)
for (k,t),v in rec['ai'].items(): ai[k,t]=v
ak=np.array([rec['ak'][k] for k in range(K)],dtype=np.int32)
with rb.Context(gm, N, T, rng_mode=0, information_form=True) as ctx:
    def cb(k,t):
        s=ctx.read_particles(); inf=ctx.read_information(); r=taps[(k,t)]
        lw=s['logw']-s['logw'].max(); lr=r['logw']-r['logw'].max()
        print(k,t,'logw %.2e'%np.abs(lw-lr).max(),'hld %.2e'%np.abs(inf['halfLogDetP']-r['hld']).max(),'ivec %.2e'%(np.abs(inf['ivec']-r['ivec']).max()/np.abs(r['ivec']).max()),'Imat %.2e'%(np.abs(inf['Imat']-r['Imat'].transpose(1,2,0)).max()/np.abs(r['Imat']).max()), 'P %.2e'%(np.abs(s['P']-r['P'].transpose(1,2,0)).max()/np.abs(r['P']).max()), 'xl %.2e'%(np.abs(s['xl']-r['xl']).max()))
    ctx.set_step_callback(cb)
    o=ctx.smoother_run(*args, pr["dt"], K, 1, streams=st, forced_ancestors=ai, forced_ak=ak, want_AI=True)
for t in range(1,T):
    p=rec['paNt'][(1,t)]
    print('AI t=%d err %.2e'%(t, np.abs(o['AI'][:,t,1]-p).max()/p.max()))
