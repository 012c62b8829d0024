"""Sharded filter (one process per GPU, peer-memory data path) against the single-GPU
filter and the oracle.  With fewer GPUs than ranks the ranks share devices (rank r on
device r mod n_devices): CUDA IPC, peer loads / stores and the peer-memory barrier work the
same between two processes on one GPU, so a 1-GPU box still exercises the whole sharded path
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu` runs it over NVLink)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, PKG, assert_close_norm

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(rb, m, T):
    return rb.synth.dense_mag_problem(N_T=T, m=m, seed=3, m_sim=300)


def _worker(rank, world, port, N, m, T, seed, q, overlap=False, variant=2, fused=True):
    import sys
    if overlap:   # read by the library when the sharded context is created
        os.environ["RBSLAM_OVERLAP"] = "1"
    else:
        os.environ.pop("RBSLAM_OVERLAP", None)
    os.environ["RBSLAM_FUSED"] = "1" if fused else "0"   # 0: fetch the migrants ahead of the pass (k_peer_fetch)
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rbslam
        from rbslam.dist import ShardedFilter
        pr = _problem(rbslam, m, T)
        gm = rbslam.models.from_problem(pr)
        from rbslam import _capi as _c
        ndev = _c.lib().rbslam_device_count()
        with ShardedFilter(gm, N, T, rank=rank, world=world, device=rank % ndev, seed=seed,
                           kalman_variant=variant) as ctx:
            o = ctx.filter_run(pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"],
                               pr["Q"], pr["R"], pr["dt"], want_xn_traj=True)
            dist.barrier()
        if rank == 0:
            q.put(("ok", {k: np.array(v) for k, v in o.items()}))
        else:
            q.put(("ok", None))
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N,m,T,overlap,variant,fused", [
    # fused migration (default): migrants are read in place from their exporter by the Kalman pass
    (2, 32, 64, 12, False, 2, True), (2, 64, 253, 8, False, 2, True), (4, 64, 64, 8, False, 2, True),
    (2, 32, 64, 12, False, 7, True), (2, 64, 253, 8, False, 7, True), (4, 64, 64, 8, False, 7, True),
    # fetch-then-pass (RBSLAM_FUSED=0) and its overlapped variant
    (2, 32, 64, 12, False, 2, False), (4, 64, 64, 8, False, 2, False), (2, 64, 253, 8, False, 7, False),
    (2, 32, 64, 12, True, 2, False), (2, 64, 253, 8, True, 2, False), (4, 64, 64, 8, True, 2, False),
])
def test_sharded_filter_matches_single_gpu_and_oracle(rbslam_lib, world, N, m, T, overlap, variant, fused):
    rb = rbslam_lib
    import torch.multiprocessing as mp
    import oracle
    seed = 77
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, N, m, T, seed, q, overlap, variant, fused)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" for r in res), res
    sh = [r[1] for r in res if r[1] is not None][0]
    pr = _problem(rb, m, T)
    gm = rb.models.from_problem(pr)
    args = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
    with rb.Context(gm, N, T, rng_mode=1, seed=seed, kalman_variant=variant) as ctx:
        single = ctx.filter_run(*args, pr["dt"], want_xn_traj=True)
    # G-invariance: same ancestors, same weights -> (near) identical outputs for every GPU count
    for k in ["traj_max", "traj_mean", "xl_max", "xl_mean", "P_max", "P_mean", "traj_sample_iwmax", "xn_traj"]:
        assert_close_norm(sh[k], single[k], 1e-12, "sharded vs single: " + k)
    om = oracle.DenseMag3D(pr["NN"], pr["L"])
    st = oracle.Streams.from_philox(seed, 1, T, N, om.nz)
    ref = oracle.particleFilter(om, *args, N, pr["dt"], st)
    for k, r in zip(["traj_max", "traj_mean", "xl_max", "xl_mean", "P_max", "P_mean"], ref):
        assert_close_norm(sh[k], r, 1e-7, "sharded vs oracle: " + k)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_device_planner_equals_host_planner(rbslam_lib, world):
    """k_plan_shard (device) must reproduce rbslam_plan_shard (host reference) exactly, and its
    work lists must cover every local slab once with the in-place / copy / fetch roles right."""
    import ctypes as C
    from rbslam import _capi
    L = _capi.lib()
    rng = np.random.default_rng(world)
    # the largest size gives every thread of the 8-CTA planning cluster several particles
    for N, steps in ((64 * world, 5), (1000 * world, 5), (3000 * world, 2)):
        Nloc = N // world
        owner = (np.arange(N) // Nloc).astype(np.int32)
        lslot = (np.arange(N) % Nloc).astype(np.int32)
        for step in range(steps):
            w = rng.random(N) ** (1 + 5 * (step % 3))
            ai = rng.choice(N, size=N, p=w / w.sum()).astype(np.int32)
            ho, hl = np.zeros(N, np.int32), np.zeros(N, np.int32)
            nm = C.c_int32()
            assert L.rbslam_plan_shard(N, world, _capi.iptr(ai), _capi.iptr(owner), _capi.iptr(lslot),
                                       _capi.iptr(ho), _capi.iptr(hl), C.byref(nm)) == 0
            n_child = np.bincount(ai, minlength=N)
            for rank in range(world):
                do, dl = np.zeros(N, np.int32), np.zeros(N, np.int32)
                src, glob = np.zeros(Nloc, np.int32), np.zeros(Nloc, np.int32)
                lA, lB = np.zeros(Nloc, np.int32), np.zeros(Nloc, np.int32)
                fetch, cnt = np.zeros(4 * Nloc, np.int32), np.zeros(8, np.int32)
                rc = L.rbslam_op_plan_shard(0, N, world, rank, *[_capi.iptr(x) for x in
                                            (ai, owner, lslot, do, dl, src, glob, lA, lB, fetch, cnt)])
                assert rc == 0
                assert np.array_equal(do, ho) and np.array_equal(dl, hl)
                nA0, nB0, nA1, nB1, nF, nmig = cnt[:6]
                assert nmig == nm.value
                assert nA0 + nA1 + nB0 + nB1 == Nloc
                items = np.concatenate([lA[:nA0 + nA1], lB[:nB0 + nB1]])
                assert sorted(items) == list(range(Nloc))                    # every local slab once
                mine = np.flatnonzero(do == rank)
                assert np.array_equal(np.sort(glob), np.sort(mine))
                assert np.array_equal(dl[glob], np.arange(Nloc))             # glob[j] lives in slot j
                migr = {int(f) for f in fetch[:4 * nF:4]}
                assert nF == np.count_nonzero(owner[ai[mine]] != rank)
                exported = np.zeros(N, bool)
                exported[ai[do != owner[ai]]] = True
                for grp, (la, lb) in enumerate([(lA[:nA0], lB[:nB0]), (lA[nA0:nA0 + nA1], lB[nB0:nB0 + nB1])]):
                    for j in la:                                             # copies: src != dst, local source
                        a = ai[glob[j]]
                        assert owner[a] == rank and src[j] == lslot[a] and src[j] != j
                        assert grp == 1 or not exported[a]
                    for j in lb:
                        a = ai[glob[j]]
                        if j in migr:
                            assert grp == 1 and owner[a] != rank and src[j] == j
                        else:
                            assert owner[a] == rank and src[j] == lslot[a] == j
                            assert grp == 1 or not exported[a]
                for f in range(nF):                                          # migrants land in dead slabs
                    j, sr, ss = fetch[4 * f:4 * f + 3]
                    a = ai[glob[j]]
                    assert owner[a] == sr and lslot[a] == ss
                    old = np.flatnonzero((owner == rank) & (lslot == j))[0]
                    assert n_child[old] == 0
            owner, lslot = ho.copy(), hl.copy()


@pytest.mark.parametrize("world,N,m,T,variant", [(2, 32, 64, 10, 2), (2, 64, 253, 8, 7), (4, 64, 64, 8, 2),
                                                  (2, 4096, 64, 4, 2)])   # N >= 4096: the parallel resampling path
def test_group_single_process_matches_single_gpu(rbslam_lib, world, N, m, T, variant):
    """rbslam_create_group: the same sharded filter driven from ONE process (what a MATLAB caller
    has).  Shards sit on devices 0..n-1 cyclically (all on device 0 on a 1-GPU box); outputs must
    equal the unsharded run to rounding, for every shard count."""
    rb = rbslam_lib
    from rbslam import _capi
    ndev = _capi.lib().rbslam_device_count()
    pr = _problem(rb, m, T)
    gm = rb.models.from_problem(pr)
    args = (pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"])
    seed = 31
    with rb.Context(gm, N, T, rng_mode=1, seed=seed, kalman_variant=variant,
                    devices=[r % ndev for r in range(world)]) as ctx:
        grp = ctx.filter_run(*args, pr["dt"], want_xn_traj=True)
        n_launch = ctx.counters()["kernel_launches"]
        grp2 = ctx.filter_run(*args, pr["dt"], want_xn_traj=False)      # a context is reusable
    with rb.Context(gm, N, T, rng_mode=1, seed=seed, kalman_variant=variant) as ctx:
        single = ctx.filter_run(*args, pr["dt"], want_xn_traj=True)
    assert n_launch > 0
    for k in ["traj_max", "traj_mean", "xl_max", "xl_mean", "P_max", "P_mean", "traj_sample_iwmax", "xn_traj"]:
        assert_close_norm(grp[k], single[k], 1e-12, "group vs single: " + k)
    for k in ["traj_max", "xl_mean", "P_max"]:
        assert np.array_equal(grp2[k], grp[k]), k


def test_group_rejects_what_it_cannot_do(rbslam_lib):
    rb = rbslam_lib
    pr = _problem(rb, 64, 6)
    gm = rb.models.from_problem(pr)
    with pytest.raises(rb.RbslamError):
        rb.Context(gm, 33, 6, rng_mode=1, devices=[0, 0])              # N not divisible by the shard count
    with pytest.raises(rb.RbslamError):
        rb.Context(gm, 32, 6, rng_mode=0, devices=[0, 0])              # injected streams are single-GPU only
    with rb.Context(gm, 32, 6, rng_mode=1, kalman_variant=2, devices=[0, 0]) as ctx:
        with pytest.raises(rb.RbslamError):
            ctx.smoother_run(pr["odometry"], pr["y"], pr["x0_nonLin"], pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"],
                             pr["dt"], 2)


def test_group_on_one_gpu_as_the_first_thing_a_process_does():
    """Cold start: a fresh process whose FIRST device work is a single-process group with both shards on
    one GPU.  Shard 0's peer barrier spins until shard 1's step is enqueued; with lazy module loading the
    first launch of one of shard 1's kernels waited for that spinning kernel (a 20 s barrier time-out:
    'a peer rank did not reach the barrier').  The library requests eager module loading when it is loaded."""
    import subprocess
    import sys
    code = (
        "import sys, time\n"
        "sys.path[:0] = [%r, %r]\n"
        "import rbslam\n"
        "pr = rbslam.synth.dense_mag_problem(N_T=6, m=64, seed=3, m_sim=300)\n"
        "gm = rbslam.models.from_problem(pr)\n"
        "a = (pr['odometry'], pr['y'], pr['x0_nonLin'], pr['x0_lin'], pr['P0_lin'], pr['Q'], pr['R'])\n"
        "t0 = time.time()\n"
        "with rbslam.Context(gm, 32, 6, rng_mode=1, seed=5, kalman_variant=2, devices=[0, 0]) as ctx:\n"
        "    o = ctx.filter_run(*a, pr['dt'])\n"
        "print('GROUP_OK', float(o['xl_mean'][0]), round(time.time() - t0, 2))\n" % (ROOT, PKG))
    env = {k: v for k, v in os.environ.items() if k != "CUDA_MODULE_LOADING"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "GROUP_OK" in r.stdout, (r.stdout[-400:], r.stderr[-1200:])
