"""world_size-2 host-side logic on CPU (gloo): row sharding, unique-id exchange plumbing and the
cross-shard top-k merge that answers a query from a row-sharded index.  The compute inside each
rank is the CPU oracle here (no GPU in this tier); the same helpers drive the NCCL path on GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lance_b200 import parallel, synth
        from oracle import binding as ob
        n, d, K, M, k, nprobes = 6000, 32, 16, 4, 10, 4
        data = synth.gaussian_mixture(n, d, n_components=K, seed=1)
        queries = synth.gaussian_mixture(20, d, n_components=K, seed=2)
        # a global model (what the NCCL-all-reduced training produces on every rank)
        cent, _, _ = ob.kmeans_train(data[:4096], K, max_iters=8, seed=3)
        part, _, _ = ob.compute_membership(cent, data)
        res = ob.compute_residual(cent, data, part)
        cb, _ = ob.pq_train(res[:4096], M, max_iters=4, seed=4)
        # object broadcast used for the NCCL unique id
        box = [b"x" * 128 if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert box[0] == b"x" * 128
        # this rank's shard -> its own CSR index
        lo, hi = parallel.shard_rows(n, rank, world)
        p_s, codes_s = part[lo:hi], ob.pq_encode(cb, res[lo:hi])
        order = np.argsort(p_s, kind="stable")
        off = np.zeros(K + 1, np.uint64)
        off[1:] = np.cumsum(np.bincount(p_s, minlength=K))
        rid = (order + lo).astype(np.uint64)
        ids, dd, _ = ob.ivfpq_search(cent, cb, off, codes_s[order], rid, queries, k, nprobes)
        from tools import dist_util
        gi, gd = dist_util.gather_merge_topk(dist, ids, dd, k)
        if rank == 0:
            # unsharded reference
            codes = ob.pq_encode(cb, res)
            order = np.argsort(part, kind="stable")
            off = np.zeros(K + 1, np.uint64)
            off[1:] = np.cumsum(np.bincount(part, minlength=K))
            ri, rd, _ = ob.ivfpq_search(cent, cb, off, codes[order], order.astype(np.uint64), queries, k, nprobes)
            ok = bool(np.array_equal(np.sort(gd, axis=1), np.sort(rd, axis=1)))
            cover = parallel.shard_rows(n, 0, world)[0] == 0 and parallel.shard_rows(n, world - 1, world)[1] == n
            q.put((ok, cover))
    finally:
        dist.destroy_process_group()


def test_sharded_search_merge_equals_unsharded_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, cover = q.get(timeout=5)
    assert ok and cover


def test_shard_rows_partition():
    from lance_b200 import parallel
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
