"""torchrun --nproc-per-node 2 tools/nccl_check.py : sharded k-means / PQ training over NCCL.
Every rank must end with bit-identical centroids; the loss must match a single-GPU run on the
concatenated sample to ~1e-6 (the all-reduce changes only the f32 summation order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import lance_b200 as lb
from lance_b200 import parallel, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
lb.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n, d, K = 16384, 128, 128  # n * world == sample_rate * K: both runs use exactly the same rows
full = synth.sift_like(n * world, d, seed=11)
init = full[np.random.default_rng(0).choice(n * world, K, replace=False)].copy()
single = lb.train_kmeans(full, d, K, max_iters=10, centroids=init, balance_factor=1.0) if rank == 0 else None
parallel.init_comm(dist)
lo, hi = parallel.shard_rows(n * world, rank, world)
km = lb.train_kmeans(full[lo:hi], d, K, max_iters=10, centroids=init, balance_factor=1.0)
t = torch.from_numpy(km.centroids).cuda()
ref = t.clone()
dist.broadcast(ref, 0)
same = bool(torch.equal(t, ref))
res = synth.gaussian_mixture(n * world, 64, 300, seed=3)
pq = lb.PQBuildParams(8, 8, max_iters=6, seed=5).build(res[lo:hi])
tp = torch.from_numpy(pq.codebook).cuda()
refp = tp.clone()
dist.broadcast(refp, 0)
samep = bool(torch.equal(tp, refp))
flags = torch.tensor([int(same), int(samep)], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    rel = abs(km.loss - single.loss) / single.loss
    print(f"identical centroids on all ranks: {bool(flags[0])}, identical PQ codebooks: {bool(flags[1])}; "
          f"sharded loss {km.loss:.6e} vs single-GPU {single.loss:.6e} (rel {rel:.2e}), iters {km.iters} vs {single.iters}")
    assert bool(flags[0]) and bool(flags[1]) and rel < 1e-5
parallel.comm_destroy()
dist.destroy_process_group()
