"""Synthetic workloads shaped like BASELINE.json's configs (no network -> no real SIFT).

C1 "SIFT-shaped": d=128 f32 holding small non-negative integers (SIFT descriptors are u8-valued),
generated from a low-dimensional latent mixture so that nearest neighbours are meaningful:
x = clip(round(relu(12 * (z W) + 20) + noise), 0, 255), z = component mean + N(0, I_latent).
With latent=24 / noise=3 an IVF_PQ(.,16) index reaches recall@10 ~ 0.63 without refine, the same
PQ-only ceiling the reference publishes for real SIFT-1M (0.63-0.68).
The same recipe is re-implemented with torch on the device in bench.py (other RNG, same law).
"""
import numpy as np


def sift_model(d=128, latent=24, n_components=1024, seed=1234):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((latent, d)).astype(np.float32)
    cm = (rng.standard_normal((n_components, latent)) * 1.5).astype(np.float32)
    return W, cm


def sift_like(n, d=128, n_components=1024, seed=1234, latent=24, noise=3.0, draw_seed=None, chunk=1 << 16):
    W, cm = sift_model(d, latent, n_components, seed)
    rng = np.random.default_rng(seed + 1 if draw_seed is None else draw_seed)
    out = np.empty((n, d), np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        comp = rng.integers(0, n_components, size=e - s)
        z = cm[comp] + rng.standard_normal((e - s, latent), dtype=np.float32)
        x = z @ W
        np.maximum(x * 12.0 + 20.0, 0.0, out=x)
        x += rng.standard_normal((e - s, d), dtype=np.float32) * noise
        np.clip(np.rint(x, out=x), 0.0, 255.0, out=out[s:e])
    return out


def sift_like_queries(nq, d=128, n_components=1024, seed=1234, qseed=4321, latent=24, noise=3.0):
    """held-out draws from the same law as sift_like(seed=seed)"""
    return sift_like(nq, d, n_components, seed, latent, noise, draw_seed=qseed)


def gaussian_mixture(n, d, n_components=64, seed=0, spread=4.0):
    rng = np.random.default_rng(seed)
    means = rng.standard_normal((n_components, d)).astype(np.float32) * spread
    comp = rng.integers(0, n_components, size=n)
    return (means[comp] + rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
