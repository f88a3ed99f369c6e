"""Synthetic workloads shaped like BASELINE.json's configs (no network -> no real SIFT).

C1 "SIFT-shaped": d=128 f32 holding small non-negative integers (SIFT descriptors are u8-valued,
||x|| ~ 512), drawn from a seeded mixture of Gaussians, clipped to [0, 255] and rounded.
"""
import numpy as np


def sift_like(n, d=128, n_components=1024, seed=1234, out=None, chunk=1 << 18):
    rng = np.random.default_rng(seed)
    # component means: sparse-ish non-negative, like gradient histograms
    means = rng.gamma(shape=0.6, scale=40.0, size=(n_components, d)).astype(np.float32)
    scales = rng.uniform(6.0, 22.0, size=(n_components, 1)).astype(np.float32)
    if out is None:
        out = np.empty((n, d), np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        comp = rng.integers(0, n_components, size=e - s)
        x = means[comp] + rng.standard_normal((e - s, d), dtype=np.float32) * scales[comp]
        np.clip(np.rint(x, out=x), 0.0, 255.0, out=out[s:e])
    return out


def sift_like_queries(nq, d=128, n_components=1024, seed=1234, qseed=4321):
    # same mixture (same means), independent draws
    rng = np.random.default_rng(seed)
    means = rng.gamma(shape=0.6, scale=40.0, size=(n_components, d)).astype(np.float32)
    scales = rng.uniform(6.0, 22.0, size=(n_components, 1)).astype(np.float32)
    r2 = np.random.default_rng(qseed)
    comp = r2.integers(0, n_components, size=nq)
    x = means[comp] + r2.standard_normal((nq, d), dtype=np.float32) * scales[comp]
    return np.clip(np.rint(x), 0.0, 255.0).astype(np.float32)


def gaussian_mixture(n, d, n_components=64, seed=0, spread=4.0):
    rng = np.random.default_rng(seed)
    means = rng.standard_normal((n_components, d)).astype(np.float32) * spread
    comp = rng.integers(0, n_components, size=n)
    return (means[comp] + rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
