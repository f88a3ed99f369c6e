"""Per-step wall time of the end-to-end build (pinned host -> index -> pinned host) + plain copy bandwidths:
separates a slow host<->device link on a given box from anything the library does."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import lance_b200 as lb
from lance_b200 import synth
lb.set_device(0)
n, d = 1_000_000, 128
pin = lb.PinnedArray((n, d), np.float32)
pin.array[:] = synth.sift_like(n, d)
dev = lb.DeviceArray.from_numpy(pin.array)
params = lb.IvfBuildParams(num_partitions=256, num_sub_vectors=16, seed=7)
pins = {"centroids": lb.PinnedArray((256, d), np.float32), "codebook": lb.PinnedArray((16, 256, 8), np.float32),
        "part_offsets": lb.PinnedArray((257,), np.uint64), "codes": lb.PinnedArray((n, 16), np.uint8),
        "row_ids": lb.PinnedArray((n,), np.uint64)}
host_out = {k: v.array for k, v in pins.items()}
for i in range(4):
    t0 = time.perf_counter(); lb.lib().lb2_memcpy(C.c_void_p(dev.ptr), C.c_void_p(pin.ptr), C.c_size_t(n * d * 4)); t1 = time.perf_counter()
    print(f"plain H2D of 512 MB from pinned: {1e3*(t1-t0):.2f} ms -> {0.512/(t1-t0):.1f} GB/s")
for i in range(14):
    lb.synchronize()
    t0 = time.perf_counter(); ix = lb.IvfPqIndex.build(pin, "l2", params); t1 = time.perf_counter()
    ix.export(out=host_out); t2 = time.perf_counter(); ix.close(); t3 = time.perf_counter()
    st = ix.stats
    print(f"step {i}: build {1e3*(t1-t0):.2f} ms (device total {st.ms_total:.2f}: ivf {st.ms_ivf_train:.2f} pq {st.ms_pq_train:.2f} "
          f"transform {st.ms_transform:.2f} group {st.ms_group:.2f}) export {1e3*(t2-t1):.2f} close {1e3*(t3-t2):.2f}", flush=True)
try:
    print(open("/proc/self/status").read().split("Cpus_allowed_list:")[1].split("\n")[0].strip(), "cpus allowed;",
          os.popen("nvidia-smi topo -m 2>/dev/null | head -4").read())
except Exception as e:
    print("topo:", e)
