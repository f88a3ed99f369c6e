"""Where does a Lloyd iteration's time go?  graph replay vs eager launches vs sum of kernel times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lance_b200 as lb
from lance_b200 import synth

lb.set_device(0)
data = synth.sift_like(65536, 128, seed=5)
dd = lb.DeviceArray.from_numpy(data)
init = data[np.random.default_rng(0).choice(65536, 256, replace=False)].copy()
res = synth.gaussian_mixture(65536, 128, 300, seed=3)
rd = lb.DeviceArray.from_numpy(res)
cb0 = np.stack([res[np.random.default_rng(1).choice(65536, 256, replace=False)][:, m * 8:(m + 1) * 8] for m in range(16)])
ITERS = 24


def run_ivf():
    return lb.train_kmeans(dd, 128, 256, max_iters=ITERS, centroids=init, balance_factor=1.0, tolerance=0.0)


def run_pq():
    return lb.PQBuildParams(16, 8, max_iters=ITERS, codebook=cb0).build(rd)


for name, fn in (("ivf", run_ivf), ("pq", run_pq)):
    for mode in ("graph", "eager"):
        if mode == "eager":
            os.environ["LB2_NO_GRAPH"] = "1"
        else:
            os.environ.pop("LB2_NO_GRAPH", None)
        fn()
        lb.synchronize()
        t0 = time.perf_counter()
        lb.timer_start()
        for _ in range(3):
            fn()
        ms = lb.timer_stop() / 3
        wall = (time.perf_counter() - t0) * 1e3 / 3
        print(f"{name} {mode}: {ms:.3f} ms device ({wall:.3f} ms wall) for {ITERS} iterations -> {ms / ITERS * 1e3:.1f} us/iter")
    os.environ.pop("LB2_NO_GRAPH", None)
    lb.profile.reset(); lb.profile.enable(True)
    fn()
    lb.profile.enable(False)
    tot = 0.0
    for k, (c, ms) in sorted(lb.profile.dump().items()):
        print(f"   {k:32s} {c:4d} launches {ms / max(c,1) * 1e3:8.1f} us each")
        tot += ms
    print(f"   sum of kernel times {tot:.3f} ms -> {tot / ITERS * 1e3:.1f} us/iter")
