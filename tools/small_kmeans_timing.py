"""Times lb2_kmeans_train on the small problems hierarchical k-means produces (k' <= 16, a few hundred .. a few
thousand rows), fused single-launch path vs the multi-kernel path (LB2_NO_SMALL_KMEANS=1 in a second process)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lance_b200 as lb

lb.set_device(0)
rng = np.random.default_rng(0)
tag = "generic" if os.environ.get("LB2_NO_SMALL_KMEANS") else "fused"
for n, k, d in ((512, 2, 128), (2048, 8, 128), (4096, 16, 128), (8192, 16, 128), (16384, 8, 128), (1024, 16, 768)):
    x = (rng.standard_normal((n, d)) + rng.integers(0, k, n)[:, None] * 0.7).astype(np.float32)
    dx = lb.DeviceArray.from_numpy(x)
    lb.train_kmeans(dx, d, k, max_iters=50, seed=1)
    lb.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for r in range(reps):
        km = lb.train_kmeans(dx, d, k, max_iters=50, seed=1 + r)
    lb.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"[{tag}] n={n:6d} k={k:3d} d={d:4d}: {dt*1e3:8.3f} ms per train_kmeans call, last run {km.iters} iterations "
          f"-> {dt*1e6/max(km.iters,1):7.1f} us / iteration (incl. call overhead)", flush=True)
