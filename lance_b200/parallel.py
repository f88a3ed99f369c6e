"""Multi-GPU host plumbing: one process per GPU (torch.distributed for rendezvous only).

* `init_comm(dist)`  -- rank 0 creates the NCCL unique id (lb2_comm_unique_id), it is broadcast with
  torch.distributed (gloo or nccl object broadcast), every rank calls lb2_comm_init: from then on
  lb2_kmeans_train / lb2_pq_train / lb2_ivfpq_build all-reduce the per-cluster sums over NVLink
   and every rank ends with the SAME centroids / codebook and an index over its
  own row shard.
* `shard_rows` / `merge_topk` -- how a query is answered by a row-sharded index: every rank searches
  its shard, the per-rank (row id, distance) lists are gathered and merged by (distance, row id),
  exactly the ordering of the reference's final SortExec (rust/lance/src/dataset/scanner.rs:3450-3466).
"""
import ctypes as C

import numpy as np

from ._lib import check, lib


def shard_rows(n, rank, world):
    """contiguous row range [lo, hi) of `rank` (contiguous row ranges per GPU)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def unique_id():
    buf = (C.c_uint8 * 128)()
    check(lib().lb2_comm_unique_id(buf))
    return bytes(buf)


def comm_init(uid, rank, world):
    buf = (C.c_uint8 * 128).from_buffer_copy(uid)
    check(lib().lb2_comm_init(buf, C.c_int(rank), C.c_int(world)))


def comm_destroy():
    check(lib().lb2_comm_destroy())


def init_comm(dist):
    """dist = torch.distributed (already initialised)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm_init(box[0], rank, world)
    return rank, world


def merge_topk(ids_list, dists_list, k):
    """merge per-shard results [(nq, k_i)] into the global top-k by (distance, row id)."""
    ids = np.concatenate(ids_list, axis=1)
    dists = np.concatenate(dists_list, axis=1)
    nq = ids.shape[0]
    out_i = np.full((nq, k), np.iinfo(np.uint64).max, np.uint64)
    out_d = np.full((nq, k), np.inf, np.float32)
    for q in range(nq):
        order = np.lexsort((ids[q], dists[q]))[:k]
        keep = order[np.isfinite(dists[q][order]) | (ids[q][order] != np.iinfo(np.uint64).max)]
        out_i[q, :len(keep)] = ids[q][keep]
        out_d[q, :len(keep)] = dists[q][keep]
    return out_i, out_d


def gather_merge_topk(dist, ids, dists, k):
    """all-gather the per-rank candidate lists (CPU tensors / numpy) and merge on every rank."""
    import torch
    world = dist.get_world_size()
    ti, td = torch.from_numpy(ids.view(np.int64)), torch.from_numpy(dists)
    gi = [torch.empty_like(ti) for _ in range(world)]
    gd = [torch.empty_like(td) for _ in range(world)]
    dist.all_gather(gi, ti)
    dist.all_gather(gd, td)
    return merge_topk([g.numpy().view(np.uint64) for g in gi], [g.numpy() for g in gd], k)
