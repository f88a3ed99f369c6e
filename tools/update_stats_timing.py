"""Diagnostics: per-iteration time of the two halves of kmeans_update_stats and the cluster-size skew
(run on the GPU box: LB2_SPLIT_UPDATE_STATS=1 python tools/update_stats_timing.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lance_b200 as lb
import bench

lb.set_device(0)
data_t, _ = bench.device_dataset(torch, 200000, 16, 1000, torch.device("cuda"))
data = data_t.cpu().numpy()
for iters in (1, 3, 10, 32):
    lb.profile.reset(); lb.profile.enable(True)
    km = lb.train_kmeans(data, 128, 256, max_iters=iters, balance_factor=1.0, seed=7)
    lb.profile.enable(False)
    p, _, _ = lb.compute_partitions(km.centroids, data[:65536])
    sz = np.bincount(p, minlength=256)
    print("iters", km.iters, "update", lb.profile.get("kmeans_update_only"), "stats", lb.profile.get("kmeans_stats_only"),
          "sizes max/mean/min", sz.max(), sz.mean(), sz.min(), flush=True)
# PQ sub-space training on residuals
part, _, _ = lb.compute_partitions(km.centroids, data)
res = lb.compute_residual(km.centroids, data, part)
lb.profile.reset(); lb.profile.enable(True)
pq = lb.PQBuildParams(16, 8, max_iters=10).build(res)
lb.profile.enable(False)
print("pq update", lb.profile.get("kmeans_update_only"), "stats", lb.profile.get("kmeans_stats_only"))
codes = pq.quantize(res[:65536])
mx = [np.bincount(codes[:, m], minlength=256).max() for m in range(16)]
print("pq max cluster sizes per sub-space", mx)
