"""torchrun --nproc-per-node N tools/nccl_check.py : the sharded paths over NCCL.
  1. flat k-means / PQ training on row shards: every rank ends with bit-identical models; the loss matches a
     single-GPU run on the concatenated sample to ~1e-6 (the exchange changes only the f32 summation order).
  2. hierarchical k-means (K > 256) on row shards: bit-identical centroids on every rank, K distinct clusters,
     loss within a few % of the single-GPU tree.
  3. sharded index build + lb2_index_search_sharded == one index over all rows with the same model.
  4. lb2_index_repartition (device all-to-all): rank g ends with exactly the partitions p % world == g of the
     one-index-over-all-rows (codes and row ids in its storage order); search_sharded on it is unchanged.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import lance_b200 as lb
from lance_b200 import parallel, synth
from tools import dist_util

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
lb.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))


def same_everywhere(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, 0)
    f = torch.tensor([int(torch.equal(t, ref))], device="cuda")
    dist.all_reduce(f, op=dist.ReduceOp.MIN)
    return bool(f.item())


n, d, K = 16384, 128, 128  # n * world == sample_rate * K: both runs use exactly the same rows
full = synth.sift_like(n * world, d, seed=11)
init = full[np.random.default_rng(0).choice(n * world, K, replace=False)].copy()
single = lb.train_kmeans(full, d, K, max_iters=10, centroids=init, balance_factor=1.0) if rank == 0 else None
single_rand = lb.train_kmeans(full, d, K, max_iters=10, balance_factor=1.0, seed=9) if rank == 0 else None
big_n, big_k = 40000 * world, 600
hdata = synth.gaussian_mixture(big_n, 64, n_components=200, seed=21)
single_h = lb.train_kmeans(hdata, 64, big_k, max_iters=8, seed=4, sample_rate=10**6) if rank == 0 else None
dist_util.init_comm(dist)
lo, hi = parallel.shard_rows(n * world, rank, world)
t0 = time.perf_counter()
km = lb.train_kmeans(full[lo:hi], d, K, max_iters=10, centroids=init, balance_factor=1.0)
t_flat = time.perf_counter() - t0
ok1 = same_everywhere(km.centroids)
# random init: the picks range over the global rows, so the sharded run starts from the same rows
km_r = lb.train_kmeans(full[lo:hi], d, K, max_iters=10, balance_factor=1.0, seed=9)
ok1r = same_everywhere(km_r.centroids)
res = synth.gaussian_mixture(n * world, 64, 300, seed=3)
pq = lb.PQBuildParams(8, 8, max_iters=6, seed=5).build(res[lo:hi])
ok2 = same_everywhere(pq.codebook)
# hierarchical
hlo, hhi = parallel.shard_rows(big_n, rank, world)
t0 = time.perf_counter()
kh = lb.train_kmeans(hdata[hlo:hhi], 64, big_k, max_iters=8, seed=4, sample_rate=10**6)
t_h = time.perf_counter() - t0
ok3 = same_everywhere(kh.centroids) and len(np.unique(kh.centroids, axis=0)) == big_k
# sharded build + sharded search
nb, db, Kb, Mb = 30000 * world, 64, 32, 8
bdata = synth.gaussian_mixture(nb, db, n_components=Kb, seed=31)
q = synth.gaussian_mixture(64, db, n_components=Kb, seed=32)
blo, bhi = parallel.shard_rows(nb, rank, world)
ix = lb.IvfPqIndex.build(bdata[blo:bhi], "l2", lb.IvfBuildParams(num_partitions=Kb, num_sub_vectors=Mb, max_iters=8, pq_max_iters=6),
                         row_ids=np.arange(blo, bhi, dtype=np.uint64))
parts = ix.export()
ok4 = same_everywhere(parts["centroids"]) and same_everywhere(parts["codebook"])
ids, dd = ix.search_sharded(q, k=10, nprobes=6)
ok5 = same_everywhere(ids) and same_everywhere(dd)
# partition ownership: every rank derives the whole index itself (same model, deterministic transform) and compares
own = ix.repartition()
op = own.export()
wpart, wcodes, _ = lb.ivfpq_transform(parts["centroids"], parts["codebook"], bdata)
wref = lb.IvfPqIndex.from_parts(parts["centroids"], parts["codebook"], wpart, wcodes).export()
sizes_w, sizes_o = np.diff(wref["part_offsets"]), np.diff(op["part_offsets"])
ok7 = True
for p_ in range(Kb):
    if p_ % world == rank:
        a0, a1 = int(wref["part_offsets"][p_]), int(wref["part_offsets"][p_ + 1])
        b0, b1 = int(op["part_offsets"][p_]), int(op["part_offsets"][p_ + 1])
        ok7 &= (a1 - a0 == b1 - b0) and np.array_equal(wref["row_ids"][a0:a1], op["row_ids"][b0:b1]) \
            and np.array_equal(wref["codes"][a0:a1], op["codes"][b0:b1])
    else:
        ok7 &= sizes_o[p_] == 0
ri, rd = own.search_sharded(q, k=10, nprobes=6)
ok7 = bool(ok7) and bool(np.array_equal(ri, ids) and np.array_equal(rd, dd))
f7 = torch.tensor([int(ok7)], device="cuda")
dist.all_reduce(f7, op=dist.ReduceOp.MIN)
ok7 = bool(f7.item())
if rank == 0:
    rel = abs(km.loss - single.loss) / single.loss
    relr = abs(km_r.loss - single_rand.loss) / single_rand.loss
    # one index over ALL rows with the sharded model (single-rank code path: the communicator only matters in builds)
    part, codes, _ = lb.ivfpq_transform(parts["centroids"], parts["codebook"], bdata)
    whole = lb.IvfPqIndex.from_parts(parts["centroids"], parts["codebook"], part, codes)
    wi, wd = whole.search(q, k=10, nprobes=6)
    ok6 = bool(np.array_equal(wi, ids) and np.array_equal(wd, dd))
    def loss_of(c):
        _, dists, _ = lb.compute_partitions(c, hdata)
        return float(dists.astype(np.float64).sum())
    lh, ls = loss_of(kh.centroids), loss_of(single_h.centroids)
    print(f"[nccl_check world={world}] flat: identical={ok1} (random init {ok1r}), pq identical={ok2}, loss rel {rel:.2e} "
          f"(random init {relr:.2e}), iters {km.iters} vs {single.iters}, {t_flat*1e3:.1f} ms | hierarchical K={big_k}: "
          f"identical+distinct={ok3}, loss sharded/single {lh/ls:.4f}, {t_h*1e3:.0f} ms | sharded build identical={ok4}, "
          f"search_sharded identical on ranks={ok5}, equals whole-index search={ok6}, repartition (all-to-all) == whole index "
          f"per owned partition + same search={ok7}")
    assert ok1 and ok1r and ok2 and ok3 and ok4 and ok5 and ok6 and ok7 and rel < 1e-5 and relr < 1e-5 and 0.9 < lh / ls < 1.1
parallel.comm_destroy()
dist.destroy_process_group()
