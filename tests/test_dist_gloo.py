"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: the migration
planner every rank runs redundantly after resampling, and the data plane it implies
(emulated with gloo send/recv on CPU tensors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, steps, q):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rbslam
        rng = np.random.default_rng(123)          # same stream on every rank (replicated decisions)
        owner = np.sort(np.arange(N) % world).astype(np.int32)
        # "slab" of particle i: a small vector tagged with its genealogy hash
        data = {i: np.full(4, float(i)) for i in range(N) if owner[i] == rank}
        truth = np.arange(N, dtype=np.float64)    # replicated ground truth of every particle's tag
        total_mig = 0
        for step in range(steps):
            w = rng.random(N) ** (1 + 3 * (step % 3))
            ai = rng.choice(N, size=N, p=w / w.sum()).astype(np.int32)
            new_owner, nmig = rbslam.plan_migration(ai, owner, world)
            # every rank must derive the identical plan
            gathered = [torch.zeros(N, dtype=torch.int32) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(new_owner.copy()))
            for g in gathered:
                assert torch.equal(g, gathered[0])
            total_mig += nmig
            # data plane: ship ancestor payloads whose offspring live elsewhere
            new_data = {}
            sends, recvs = [], []
            for i in range(N):
                src, dst = owner[ai[i]], new_owner[i]
                if src == rank and dst == rank:
                    new_data[i] = data[ai[i]].copy()
                elif src == rank:
                    sends.append((i, dst))
                elif dst == rank:
                    recvs.append((i, src))
            # deterministic order; rank 0 sends first to avoid a deadlock with blocking calls
            for phase in range(2):
                if (phase == 0) == (rank == 0):
                    for i, dst in sends:
                        dist.send(torch.from_numpy(data[ai[i]].copy()), dst=int(dst), tag=int(i))
                else:
                    for i, src in recvs:
                        buf = torch.zeros(4, dtype=torch.float64)
                        dist.recv(buf, src=int(src), tag=int(i))
                        new_data[i] = buf.numpy().copy()
            truth = truth[ai]
            owner, data = new_owner, new_data
            assert len(data) == np.count_nonzero(owner == rank)
            for i, v in data.items():
                assert np.all(v == truth[i]), (rank, step, i)
        q.put((rank, total_mig, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, -1, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [16, 200])
def test_migration_plan_two_ranks_gloo(N):
    import __graft_entry__ as g
    g.build()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[2] == "ok" for r in res), res
    assert res[0][1] == res[1][1]          # both ranks counted the same migrations
