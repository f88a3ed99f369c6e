"""Where does the end-to-end build's staging time go on this box?  LB2_TRACE_BUILD stamps + per-step wall time."""
import os, sys, time
mode = sys.argv[1] if len(sys.argv) > 1 else "trace"
if mode == "trace":
    os.environ["LB2_TRACE_BUILD"] = "1"
if mode == "nozc":
    os.environ["LB2_NO_ZERO_COPY"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lance_b200 as lb
from lance_b200 import synth
lb.set_device(0)
n, d = 1_000_000, 128
pin = lb.PinnedArray((n, d), np.float32)
pin.array[:] = synth.sift_like(n, d)
params = lb.IvfBuildParams(num_partitions=256, num_sub_vectors=16, seed=7)
for i in range(10):
    lb.synchronize()
    t0 = time.perf_counter(); ix = lb.IvfPqIndex.build(pin, "l2", params); t1 = time.perf_counter()
    st = ix.stats
    ix.close()
    print(f"[{mode}] step {i}: build {1e3*(t1-t0):.2f} ms (ivf {st.ms_ivf_train:.2f} pq {st.ms_pq_train:.2f} transform {st.ms_transform:.2f})", file=sys.stderr, flush=True)
