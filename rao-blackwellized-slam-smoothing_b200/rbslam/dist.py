"""Sharding one particle filter over the GPUs of a node, one process per GPU.

``torch.distributed`` is plumbing only: it exchanges the 64-byte CUDA-IPC handles once
(and lets benchmarks barrier / max-reduce their timings).  The per-step data path --
peer reads of migrating slabs, all-gather of the log-weights by peer stores, barriers --
runs inside librbslam over NVLink peer memory (csrc/sharded.cu).
"""
import os

from . import _capi
from .api import Context


def torch_all_gather_object(obj):
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


class ShardedFilter(Context):
    """A Context whose N particles are split over ``world`` ranks (N % world == 0).

    Every rank calls the same methods in the same order; only rank 0 receives outputs
    that need the whole population (xl_mean, P_max, xn_traj, ...)."""

    def __init__(self, model, N, T, rank=None, world=None, device=None, seed=0, kalman_variant=0,
                 all_gather_object=torch_all_gather_object, **kw):
        rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
        device = int(os.environ.get("LOCAL_RANK", str(rank))) if device is None else device
        super().__init__(model, N, T, device=device, rng_mode=_capi.RNG_PHILOX, seed=seed,
                         keep_history=True, kalman_variant=kalman_variant, rank=rank, world=world, **kw)
        if world > 1:
            self.connect_peers(all_gather_object)
