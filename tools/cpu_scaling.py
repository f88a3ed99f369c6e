"""How many host threads actually help the CPU arm on this box?  Prints the cgroup CPU limit, the affinity mask and
the oracle's k-means time per iteration (65 536 x 8 rows, 256 centroids: one PQ sub-space) at several thread counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as ob

for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "loadavg", os.getloadavg())
rng = np.random.default_rng(0)
x = rng.standard_normal((65536, 8)).astype(np.float32)
init = x[:256].copy()
for nt in (1, 4, 8, 16, 32, 64, 96, 128):
    if nt > os.cpu_count():
        break
    ob.kmeans_train(x, 256, max_iters=2, init_centroids=init, nthreads=nt)
    t = time.perf_counter()
    _, _, it = ob.kmeans_train(x, 256, max_iters=10, init_centroids=init, nthreads=nt)
    print(f"threads {nt:4d}: {(time.perf_counter() - t) / it * 1e3:8.2f} ms / Lloyd iteration", flush=True)
