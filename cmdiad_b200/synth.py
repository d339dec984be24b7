"""Synthetic patch-feature generators.

MVTec 3D-AD and the DINO / Point-MAE checkpoints are not available offline, so banks and test images are synthetic
float32 patch features with the shapes the reference produces (features.py:160-184: RGB [784,768], XYZ [3136,768];
BASELINE.json also names 1152-d and 1920-d).  numpy's PCG64 is used so the streams are identical on every platform.
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.PCG64(int(seed)))


def centroids(dim, k=2048, seed=0):
    return _rng(10_000 + seed).standard_normal((k, dim), dtype=np.float32)


def patches(n_rows, dim, seed, dist="C", k=2048, anomalous_frac=0.0, cent=None):
    """n_rows x dim float32 patch features.

    dist "G": iid N(0,1).  dist "C": clustered -- row = centroid[j] + 0.35*N(0,1), which mimics the redundancy of
    real patch features (nearest-neighbour distances << norms, so the ||a||^2+||b||^2-2ab form cancels).
    anomalous_frac of the rows get 3*N(0,1) noise instead (test images)."""
    g = _rng(seed)
    if dist == "G":
        return g.standard_normal((n_rows, dim), dtype=np.float32)
    if cent is None:
        cent = centroids(dim, k)
    j = g.integers(0, cent.shape[0], size=n_rows)
    noise = g.standard_normal((n_rows, dim), dtype=np.float32)
    scale = np.full((n_rows, 1), 0.35, dtype=np.float32)
    if anomalous_frac > 0:
        n_anom = max(1, int(round(anomalous_frac * n_rows)))
        scale[g.choice(n_rows, size=n_anom, replace=False)] = 3.0
    return (cent[j] + scale * noise).astype(np.float32)


def image_bank(n_images, patches_per_image, dim, seed, dist="C", k=2048):
    """List of per-image [P,D] patch tensors, as add_sample_to_mem_bank would append them."""
    cent = centroids(dim, k) if dist == "C" else None
    return [patches(patches_per_image, dim, seed * 1000 + i, dist, k, cent=cent) for i in range(n_images)]
