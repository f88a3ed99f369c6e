"""One small call of the streamed-tile filter per process (debug aid): python tools/tc_general_debug.py n d K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lance_b200 as lb

n, d, K = (int(a) for a in sys.argv[1:4])
os.environ["LB2_TC_STATS"] = "1"
rng = np.random.default_rng(1)
cent = (rng.standard_normal((K, d)) * 3).astype(np.float32)
data = (cent[rng.integers(0, K, n)] + rng.standard_normal((n, d))).astype(np.float32)
print("calling tc path", n, d, K, flush=True)
p1, d1, v1 = lb.compute_partitions(cent, data)
print("tc path returned", flush=True)
os.environ["LB2_DISABLE_TC"] = "1"
p2, d2, v2 = lb.compute_partitions(cent, data)
print("equal:", np.array_equal(p1, p2), np.array_equal(d1, d2), "mismatches", int((p1 != p2).sum()), flush=True)
