"""diagnostic: ONE fresh process scores the two golden RGB test images right after the bank is built (first calls of the
process, one per lane) and compares min_val / min_idx with the exact device scan; prints the differing queries."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmdiad_b200 import Bank  # noqa: E402
from tests import cases  # noqa: E402

golden = {"rgb_case": np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "rgb_case.npz"))}
g = golden["rgb_case"]
lib = cases.rgb_normalised_lib(golden)
rows = lib[g["coreset_idx_TF32"]]
patches = [((torch.from_numpy(cases.rgb_test_patch(t)) - torch.tensor(g["rgb_mean"])) / torch.tensor(g["rgb_std"])).numpy()
           for t in range(2)]
b = Bank(rows.shape[1], rows.shape[0])
b.append(rows)
b.finalize()
from oracle import restate as O  # noqa: E402
res = []
refs = []
for t in range(2):
    r = b.score(patches[t], (28, 28), 224, full=True)
    res.append((r.min_val.copy(), r.min_idx.copy(), b.score_stats()))
    refs.append(O.score_restated(patches[t], rows, (28, 28), 224))   # CPU load between the calls, as in the test
bad = 0
back = b.read().numpy()
if not np.array_equal(back, rows):
    bad += 1
    w = np.nonzero((back != rows).any(1))[0]
    print(f"BANK ROWS DIFFER on the device: {len(w)} rows, first {w[:10]}")
for t in range(2):
    mv, mi, st = res[t]
    d64 = np.sqrt(((patches[t].astype(np.float64)[:, None, :] - rows.astype(np.float64)[None]) ** 2).sum(-1))
    tv, ti = d64.min(1), d64.argmin(1)
    rel = np.abs(mv - tv) / tv
    rel_o = np.abs(refs[t]["min_val"] - tv) / tv
    w = np.nonzero((rel > 2e-5) | (rel_o > 2e-5))[0]
    if len(w):
        bad += 1
        print(f"image {t}: {len(w)} queries off the float64 brute force; ours max rel {rel.max():.3e}, torch oracle max rel {rel_o.max():.3e}")
        for i in w[:20]:
            print(f"   q {i}: ours {mv[i]!r} row {mi[i]} | torch oracle {refs[t]['min_val'][i]!r} row {refs[t]['min_idx'][i]} | float64 {tv[i]!r} row {ti[i]}")
    else:
        print(f"image {t}: ours / torch oracle within {rel.max():.2e} / {rel_o.max():.2e} of the float64 brute force")
for t in range(2):
    P = 784
    ex_val, ex_idx = np.empty(P, np.float32), np.empty(P, np.int64)
    q = np.ascontiguousarray(patches[t])
    b._lib.cmdb_debug_exact_min(b._h, q.ctypes.data, P, ex_val.ctypes.data, ex_idx.ctypes.data)
    mv, mi, st = res[t]
    d = np.nonzero((mv != ex_val) | (mi != ex_idx))[0]
    if len(d):
        bad += 1
        print(f"image {t}: {len(d)} queries differ from the exact scan; stats {st}")
        for i in d[:30]:
            print(f"   q {i} (tile {i // 128}): ours {mv[i]!r} row {mi[i]} | exact {ex_val[i]!r} row {ex_idx[i]}")
    else:
        print(f"image {t}: equal to the exact scan; stats {st}")
b.close()
sys.exit(1 if bad else 0)
