"""Seeded inputs shared by the CPU (oracle vs golden) and GPU (CUDA vs oracle / golden) tests.  The constants mirror
oracle/make_golden.py, which froze the reference's outputs for exactly these inputs."""
import numpy as np
import torch

from cmdiad_b200 import synth

RGB_CASE = dict(n_train=10, P=784, D=768, seed=11, n_test=2, fmap=28)
DUAL_CASE = dict(n_train=3, P_xyz=3136, P_rgb=784, D=768, seed=23, fmap_xyz=56, fmap_rgb=28)


def rgb_train():
    c = RGB_CASE
    return synth.image_bank(c["n_train"], c["P"], c["D"], c["seed"])


def rgb_test_patch(t):
    c = RGB_CASE
    return synth.patches(c["P"], c["D"], c["seed"] * 1000 + 500 + t, anomalous_frac=0.01, cent=synth.centroids(c["D"]))


def rgb_normalised_lib(golden):
    """the normalised training library exactly as the reference computed it (torch float32 ops, golden mean/std)"""
    g = golden["rgb_case"]
    cat = torch.cat([torch.from_numpy(x) for x in rgb_train()], 0)
    return ((cat - torch.tensor(g["rgb_mean"])) / torch.tensor(g["rgb_std"])).numpy()


def dual_train():
    c = DUAL_CASE
    return (synth.image_bank(c["n_train"], c["P_xyz"], c["D"], c["seed"]),
            synth.image_bank(c["n_train"], c["P_rgb"], c["D"], c["seed"] + 1))


def dual_test():
    c = DUAL_CASE
    cent = synth.centroids(c["D"])
    return (synth.patches(c["P_xyz"], c["D"], c["seed"] * 1000 + 700, anomalous_frac=0.01, cent=cent),
            synth.patches(c["P_rgb"], c["D"], c["seed"] * 1000 + 701, anomalous_frac=0.01, cent=cent))


def tie_aware_idx_ok(idx, ref_idx, ref_dist_rows, rtol=2e-6):
    """argmin parity: equal, or the reference's own distances of the two candidates are within float32 mm-form noise"""
    idx, ref_idx = np.asarray(idx), np.asarray(ref_idx)
    bad = np.nonzero(idx != ref_idx)[0]
    for p in bad:
        a, b = ref_dist_rows[p, idx[p]], ref_dist_rows[p, ref_idx[p]]
        if abs(a - b) > rtol * max(abs(a), abs(b)):
            return False, len(bad)
    return True, len(bad)
