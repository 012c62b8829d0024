import os, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/rao-blackwellized-slam-smoothing_b200')
import numpy as np, torch, torch.distributed as dist
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dist.init_process_group('nccl', device_id=torch.device('cuda',rank))
import rbslam
from rbslam.dist import ShardedFilter
pr = rbslam.synth.dense_mag_problem(N_T=2000, m=1024, seed=1, n_laps=10, m_sim=2000)
T=14; pr['y']=pr['y'][:T].copy(); pr['odometry']=pr['odometry'][:T].copy()
gm = rbslam.models.from_problem(pr)
fargs=(pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
ctx = ShardedFilter(gm, 10000*world, T, rank=rank, world=world, device=rank, seed=1)
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier()
    t0=time.perf_counter(); ctx.filter_begin(*fargs, pr["dt"]); ctx.sync(); t1=time.perf_counter()
    for _ in range(T): ctx.filter_step()
    t2=time.perf_counter(); ctx.sync(); t3=time.perf_counter()
    o=ctx.filter_end(); t4=time.perf_counter()
    print(rank, rep, 'begin %.3f enqueue %.3f steps %.3f end %.3f'%(t1-t0, t2-t1, t3-t1, t4-t3), flush=True)
ctx.close(); dist.destroy_process_group()
