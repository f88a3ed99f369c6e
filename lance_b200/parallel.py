"""Multi-GPU host plumbing: one process per GPU, no framework dependency.

* `unique_id()` / `comm_init(uid, rank, world)` -- rank 0 creates the NCCL unique id (lb2_comm_unique_id),
  the HOST RUNTIME (the Rust side, MPI, torch.distributed -- see tools/dist_util.py for the latter) hands
  it to every rank, every rank calls lb2_comm_init: from then on lb2_kmeans_train / lb2_pq_train /
  lb2_ivfpq_build / lb2_ivfflat_build take this rank's ROW SHARD, exchange the per-cluster partial results
  once per Lloyd iteration over NVLink and every rank ends with the SAME centroids /
  codebook and an index over its own rows; `IvfPqIndex.search_sharded` merges the per-rank results in the
  library (lb2_index_search_sharded).
* `shard_rows` -- contiguous row ranges per GPU; `merge_topk` -- the host-side statement of the merge rule
  (ascending (distance, row id), the reference's final SortExec, rust/lance/src/dataset/scanner.rs:3450-3466),
  used by the CPU tests.
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


def comm_info():
    r, n = C.c_int(0), C.c_int(1)
    check(lib().lb2_comm_info(C.byref(r), C.byref(n)))
    return r.value, n.value


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
