"""Which association does torch-CUDA use for linalg.norm on float64?  Exact-rational emulation of a few hypotheses."""
from fractions import Fraction as Fr
import numpy as np
import torch

def rn(x):  # correctly rounded Fraction -> float64
    return float(x)

def fma(a, b, c):
    return rn(Fr(a) * Fr(b) + Fr(c))

def tree(v):
    v = list(v)
    off = 1
    while off < len(v):
        for l in range(0, len(v) - off, 2 * off):
            v[l] = v[l] + v[l + off]
        off *= 2
    return v[0]

def hyp(e, name):
    d = len(e)
    W = 32
    if name == "canon_fma":       # acc_k = fma(e,e,acc_k), ((a0+a1)+a2)+a3, tree
        lanes = []
        for x in range(W):
            acc = [0.0] * 4
            for k, idx in enumerate(range(x, d, W)):
                acc[k % 4] = fma(e[idx], e[idx], acc[k % 4])
            lanes.append(((acc[0] + acc[1]) + acc[2]) + acc[3])
        return np.sqrt(tree(lanes))
    if name == "canon_mul":       # separate multiply and add
        lanes = []
        for x in range(W):
            acc = [0.0] * 4
            for k, idx in enumerate(range(x, d, W)):
                acc[k % 4] = acc[k % 4] + e[idx] * e[idx]
            lanes.append(((acc[0] + acc[1]) + acc[2]) + acc[3])
        return np.sqrt(tree(lanes))
    if name == "contract_combine":  # a0 + e1*e1 fused when acc1 starts from zero
        lanes = []
        for x in range(W):
            idxs = list(range(x, d, W))
            a = 0.0
            for idx in idxs:
                a = fma(e[idx], e[idx], a)
            lanes.append(a)
        return np.sqrt(tree(lanes))
    if name == "seq_lane16":
        W2 = 16
        lanes = []
        for x in range(W2):
            acc = [0.0] * 4
            for k, idx in enumerate(range(x, d, W2)):
                acc[k % 4] = fma(e[idx], e[idx], acc[k % 4])
            lanes.append(((acc[0] + acc[1]) + acc[2]) + acc[3])
        return np.sqrt(tree(lanes))
    if name == "abs_then":        # |x| via sqrt(x*x)?? no-op for real; placeholder for pow path: sum(pow(|x|,2))
        return None

g = np.random.Generator(np.random.PCG64(64))
for d in (64, 96):
    z = g.standard_normal((40, d))
    zt = torch.from_numpy(z).cuda()
    E = (zt - zt[5:6])
    ref = torch.linalg.norm(E, dim=1, keepdims=True).cpu().numpy()[:, 0]
    ref_sq = (torch.linalg.norm(E, dim=1) ** 2).cpu().numpy()
    Ec = E.cpu().numpy()
    assert (Ec == (z - z[5:6])).all()
    for name in ("canon_fma", "canon_mul", "contract_combine", "seq_lane16"):
        got = np.array([hyp(list(Ec[i]), name) for i in range(40)])
        print(f"d={d} {name}: mismatches {(got != ref).sum()} / 40")
    # is it a different op altogether?
    alt = torch.sqrt((E * E).sum(1)).cpu().numpy()
    print(f"d={d} sqrt(sum(E*E)) vs norm mismatches {(alt != ref).sum()}")
    alt2 = torch.linalg.vector_norm(E, 2, dim=1).cpu().numpy()
    print(f"d={d} vector_norm vs norm mismatches {(alt2 != ref).sum()}")
    Ef = E.float()
    print("float32 path check: norm(E32) vs canon ...", end=" ")
    r32 = torch.linalg.norm(Ef, dim=1).cpu().numpy()
    def canon32(e):
        lanes = []
        for x in range(32):
            acc = [np.float32(0)] * 4
            for k, idx in enumerate(range(x, len(e), 32)):
                acc[k % 4] = np.float32(Fr(float(e[idx])) * Fr(float(e[idx])) + Fr(float(acc[k % 4])))
            lanes.append(np.float32(np.float32(np.float32(acc[0] + acc[1]) + acc[2]) + acc[3]))
        v = lanes
        off = 1
        while off < 32:
            for l in range(0, 32 - off, 2 * off):
                v[l] = np.float32(v[l] + v[l + off])
            off *= 2
        return np.sqrt(np.float32(v[0]))
    e32 = Ef.cpu().numpy()
    got32 = np.array([canon32(e32[i]) for i in range(40)], dtype=np.float32)
    print("mismatches", (got32 != r32).sum(), "/ 40")
