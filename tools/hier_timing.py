"""Hierarchical k-means (K > 256) timing vs the number of concurrent split trainings (LB2_SPLIT_THREADS is read once
per process: run one process per setting).  usage: python tools/hier_timing.py [rows] [d] [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lance_b200 as lb
from bench import wrap_tensor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
g = torch.Generator(device="cuda").manual_seed(5)
cent = torch.randn(2000, d, device="cuda", generator=g) * 4
x = cent[torch.randint(0, 2000, (n,), device="cuda", generator=g)] + torch.randn(n, d, device="cuda", generator=g)
xd = wrap_tensor(lb, x, np.float32)
lb.train_kmeans(xd, d, K, max_iters=50, seed=3, sample_rate=10**9)
torch.cuda.synchronize()
t0 = time.perf_counter()
km = lb.train_kmeans(xd, d, K, max_iters=50, seed=3, sample_rate=10**9)
dt = time.perf_counter() - t0
import hashlib
print(f"[threads={os.environ.get('LB2_SPLIT_THREADS', 'default')}] n={n} d={d} K={K}: {dt*1e3:.1f} ms, model sha {hashlib.sha1(km.centroids.tobytes()).hexdigest()[:12]}", flush=True)
