"""Scan-kernel variants on the C1 index (one process, one GPU): classic vs skew (2 teams) vs skew4 (4 teams), at
nprobes 1 / 10 / 50 and batch 10 000 / 64 / 1.  Prints per-variant search ms (CUDA events) and the scan kernel's own ms."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lance_b200 as lb
from lance_b200 import synth
from bench import wrap_tensor

dev = torch.device("cuda", 0)
data_t = torch.from_numpy(synth.sift_like(1_000_000, 128, seed=1)).to(dev)
data = wrap_tensor(lb, data_t, np.float32)
ix = lb.IvfPqIndex.build(data, "l2", lb.IvfBuildParams(num_partitions=256, num_sub_vectors=16, seed=7))
q_t = torch.from_numpy(synth.sift_like_queries(10000, 128, seed=9)).to(dev)
out = {}
for nq in (10000, 64, 1):
    q = wrap_tensor(lb, q_t[:nq].contiguous(), np.float32)
    ids_t = torch.empty((nq, 10), dtype=torch.int64, device=dev)
    d_t = torch.empty((nq, 10), dtype=torch.float32, device=dev)
    o = (wrap_tensor(lb, ids_t, np.uint64), wrap_tensor(lb, d_t, np.float32))
    for nprobes in (1, 10, 50):
        ref = None
        for mode in ("classic", "skew", "skew4"):
            os.environ["LB2_SCAN"] = "classic" if mode == "classic" else "skew"
            os.environ["LB2_SCAN_TEAMS"] = "4" if mode == "skew4" else "2"
            for _ in range(3):
                ix.search(q, 10, nprobes, out=o)
            reps = 20 if nq == 10000 else 200
            lb.profile.reset(); lb.profile.enable(True)
            lb.timer_start()
            for _ in range(reps):
                ix.search(q, 10, nprobes, out=o)
            ms = lb.timer_stop() / reps
            lb.profile.enable(False)
            prof = lb.profile.dump()
            scan = sum(v[1] for k, v in prof.items() if "pq_scan" in k and "replay" not in k) / reps
            res = (ids_t.clone(), d_t.clone())
            if ref is None:
                ref = res
            same = bool(torch.equal(ref[0], res[0]) and torch.equal(ref[1], res[1]))
            out[f"nq{nq}_np{nprobes}_{mode}"] = {"search_ms": round(ms, 4), "scan_ms": round(scan, 4), "qps": round(nq / ms * 1e3), "same_as_classic": same}
            print(f"nq={nq} nprobes={nprobes} {mode}: search {ms:.4f} ms scan {scan:.4f} ms same={same}", flush=True)
# the refine operating point: k * refine_factor = 100 candidates per (query, partition), exact re-rank on the raw column
q = wrap_tensor(lb, q_t, np.float32)
ids_t = torch.empty((10000, 10), dtype=torch.int64, device=dev)
d_t = torch.empty((10000, 10), dtype=torch.float32, device=dev)
o = (wrap_tensor(lb, ids_t, np.uint64), wrap_tensor(lb, d_t, np.float32))
ref = None
for mode in ("classic", "skew", "skew4"):
    os.environ["LB2_SCAN"] = "classic" if mode == "classic" else "skew"
    os.environ["LB2_SCAN_TEAMS"] = "4" if mode == "skew4" else "2"
    for _ in range(3):
        ix.search_refine(data, q, 10, 10, 10, out=o)
    lb.profile.reset(); lb.profile.enable(True)
    lb.timer_start()
    for _ in range(10):
        ix.search_refine(data, q, 10, 10, 10, out=o)
    ms = lb.timer_stop() / 10
    lb.profile.enable(False)
    prof = {k: round(v[1] / 10, 4) for k, v in lb.profile.dump().items() if v[1] / 10 > 0.02}
    res = (ids_t.clone(), d_t.clone())
    if ref is None:
        ref = res
    same = bool(torch.equal(ref[0], res[0]) and torch.equal(ref[1], res[1]))
    out[f"refine10_nq10000_np10_{mode}"] = {"search_ms": round(ms, 4), "qps": round(10000 / ms * 1e3), "kernels_ms": prof, "same_as_classic": same}
    print(f"refine x10 nq=10000 nprobes=10 {mode}: {ms:.4f} ms same={same} {prof}", flush=True)
json.dump(out, open("gpurun_out/scan_timing.json", "w"), indent=1)
