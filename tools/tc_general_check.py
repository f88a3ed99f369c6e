"""Times the streamed-tile tensor-core filter against the exact kernels on a few large shapes
(run on the GPU box: python tools/tc_general_check.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lance_b200 as lb

os.environ["LB2_TC_STATS"] = "1"
lb.set_device(0)
NAMES = ("tc_filter", "tc_filter_general", "tc_filter_general16", "tc_rerank", "tc_refine_gather", "tc_refine_filter",
         "tc_refine_rerank", "tc_candidates", "tc_candidates_exact",
         "assign_exact_fallback", "tc_row_norms", "tc_prep_centroids", "assign_exact", "assign_exact_generic",
         "transpose_centroids")
for n, d, K in ((200000, 768, 1024), (500000, 128, 4096), (500000, 256, 256), (100000, 1536, 512), (500000, 64, 1024)):
    g = torch.Generator(device="cuda").manual_seed(n + d + K)
    lat = torch.randn(n, 24, device="cuda", generator=g)
    proj = torch.randn(24, d, device="cuda", generator=g)
    x = (lat @ proj + 0.3 * torch.randn(n, d, device="cuda", generator=g)).contiguous()
    cent = x[torch.randperm(n, device="cuda", generator=g)[:K]].cpu().numpy()
    dx = lb.DeviceArray.__new__(lb.DeviceArray)
    dx.shape, dx.dtype, dx.ptr, dx.nbytes = tuple(x.shape), np.dtype(np.float32), x.data_ptr(), x.numel() * 4
    dx.free = lambda: None
    torch.cuda.synchronize()
    os.environ.pop("LB2_DISABLE_TC", None)
    print("shape", n, d, K, "tc warm-up call", flush=True)
    lb.compute_partitions(cent, dx)
    print("  done", flush=True)
    lb.profile.reset(); lb.profile.enable(True)
    p1, d1, v1 = lb.compute_partitions(cent, dx)
    lb.profile.enable(False)
    tc = {k: lb.profile.get(k) for k in NAMES if lb.profile.get(k)[0]}
    os.environ["LB2_DISABLE_TC"] = "1"
    lb.compute_partitions(cent, dx)
    lb.profile.reset(); lb.profile.enable(True)
    p2, d2, v2 = lb.compute_partitions(cent, dx)
    lb.profile.enable(False)
    ex = {k: lb.profile.get(k) for k in NAMES if lb.profile.get(k)[0]}
    print(f"n={n} d={d} K={K}: equal={np.array_equal(p1, p2) and np.array_equal(d1, d2)}")
    print("   tc   :", tc)
    print("   exact:", ex, flush=True)
