"""Profiling targets for ncu (`--profile-from-start off`: only the region between cudaProfilerStart/Stop is
captured).  usage: ncu --set full --profile-from-start off ... python tools/ncu_targets.py <mode>
modes: c1_train, c1_transform, c1_query, c2_transform, c4_assign"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lance_b200 as lb
from lance_b200 import synth
from bench import wrap_tensor

mode = sys.argv[1]
prof = torch.cuda.profiler
dev = torch.device("cuda", 0)


def sift(n, seed):
    return torch.from_numpy(synth.sift_like(n, 128, seed=seed)).to(dev)


if mode.startswith("c1"):
    data_t = sift(1_000_000, 1)
    data = wrap_tensor(lb, data_t, np.float32)
    if mode == "c1_train":
        sample = wrap_tensor(lb, data_t[:65536].contiguous(), np.float32)
        lb.train_kmeans(sample, 128, 256, max_iters=3, balance_factor=1.0, seed=1)          # warm-up
        lb.PQBuildParams(16, 8, max_iters=3, seed=2).build(sample)
        prof.start()
        lb.train_kmeans(sample, 128, 256, max_iters=3, balance_factor=1.0, seed=1)
        lb.PQBuildParams(16, 8, max_iters=3, seed=2).build(sample)
        prof.stop()
    else:
        ix = lb.IvfPqIndex.build(data, "l2", lb.IvfBuildParams(num_partitions=256, num_sub_vectors=16, seed=7))
        parts = ix.export()
        if mode == "c1_transform":
            lb.ivfpq_transform(parts["centroids"], parts["codebook"], data)
            prof.start()
            lb.ivfpq_transform(parts["centroids"], parts["codebook"], data)
            prof.stop()
        else:
            q = wrap_tensor(lb, sift(10000, 9), np.float32)
            ix.search(q, 10, 10)
            prof.start()
            ix.search(q, 10, 10)
            ix.search_refine(data, q, 10, 10, 10)
            prof.stop()
elif mode == "c2_transform":
    g = torch.Generator(device="cuda").manual_seed(3)
    n, d, K, M = 1_000_000, 768, 4096, 96
    cen = torch.randn(3000, d, device=dev, generator=g) * 2
    data_t = cen[torch.randint(0, 3000, (n,), device=dev, generator=g)] + torch.randn(n, d, device=dev, generator=g)
    data = wrap_tensor(lb, data_t, np.float32)
    cent = data_t[torch.randperm(n, device=dev, generator=g)[:K]].cpu().numpy()
    part, _, _ = lb.compute_partitions(cent, wrap_tensor(lb, data_t[:65536].contiguous(), np.float32))
    res = (data_t[:65536].cpu().numpy() - cent[part])
    pq = lb.PQBuildParams(M, 8, max_iters=2, seed=2).build(res)
    lb.ivfpq_transform(cent, pq.codebook, data)
    prof.start()
    lb.ivfpq_transform(cent, pq.codebook, data)
    prof.stop()
elif mode == "c4_assign":
    g = torch.Generator(device="cuda").manual_seed(4)
    n, d, K = 1_000_000, 1536, 4096
    cen = torch.randn(3000, d, device=dev, generator=g)
    data_t = (cen[torch.randint(0, 3000, (n,), device=dev, generator=g)] + 0.5 * torch.randn(n, d, device=dev, generator=g)).to(torch.bfloat16)
    data = wrap_tensor(lb, data_t.view(torch.uint16), np.uint16)
    cent = data_t[torch.randperm(n, device=dev, generator=g)[:K]].view(torch.uint16).cpu().numpy()
    lb.compute_partitions(cent, data, bf16=True)
    prof.start()
    lb.compute_partitions(cent, data, bf16=True)
    prof.stop()
torch.cuda.synchronize()
print("done", mode)
