"""torch.distributed glue for the benches / tools / tests (NOT part of the product package): hands the NCCL
unique id from rank 0 to every rank and gathers numpy results for the CPU (gloo) tests."""
import numpy as np

from lance_b200 import parallel


def init_comm(dist):
    """dist = torch.distributed (already initialised): lb2_comm_init on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [parallel.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    parallel.comm_init(box[0], rank, world)
    return rank, world


def gather_merge_topk(dist, ids, dists, k):
    """all-gather the per-rank candidate lists (CPU tensors / numpy) and merge on every rank."""
    import torch
    world = dist.get_world_size()
    ti, td = torch.from_numpy(ids.view(np.int64)), torch.from_numpy(dists)
    gi = [torch.empty_like(ti) for _ in range(world)]
    gd = [torch.empty_like(td) for _ in range(world)]
    dist.all_gather(gi, ti)
    dist.all_gather(gd, td)
    return parallel.merge_topk([g.numpy().view(np.uint64) for g in gi], [g.numpy() for g in gd], k)
