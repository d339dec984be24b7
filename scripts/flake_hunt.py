"""diagnostic: repeat the scoring of the golden RGB case (tiny 784-row bank: almost every query takes the rescan / fallback
tiers) and of a mid-size bank many times on both lanes and report every call whose result differs from the first one."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmdiad_b200 import Bank, synth  # noqa: E402
from tests import cases  # noqa: E402

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 200
golden = {k: np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", k + ".npz")) for k in ("rgb_case",)}
g = golden["rgb_case"]
lib = cases.rgb_normalised_lib(golden)
bank_rows = lib[g["coreset_idx_TF32"]]


def hunt(name, rows, patches, build_knn):
    b = Bank(rows.shape[1], rows.shape[0])
    b.append(rows)
    b.finalize()
    if build_knn:
        b.build_knn()
    first = {}
    bad = 0
    for it in range(n_rep):
        for t, patch in enumerate(patches):
            r = b.score(patch, (28, 28), 224, full=True)
            st = b.score_stats()
            cur = dict(min_val=r.min_val.copy(), min_idx=r.min_idx.copy(), s=np.array(r.s), s_map=r.s_map.copy(),
                       nn_idx=np.array(r.nn_idx))
            if t not in first:
                first[t] = (cur, st)
                continue
            ref, st0 = first[t]
            diffs = {k: int((cur[k] != ref[k]).sum()) for k in cur if not np.array_equal(cur[k], ref[k])}
            if diffs:
                bad += 1
                q = np.nonzero(cur["min_val"] != ref["min_val"])[0]
                print(f"[{name}] iteration {it} image {t}: differs {diffs}; stats now {st} first {st0}")
                for i in q[:6]:
                    print(f"    query {i}: min_val {cur['min_val'][i]!r} idx {cur['min_idx'][i]} | first {ref['min_val'][i]!r} idx {ref['min_idx'][i]}")
    print(f"[{name}] {bad} differing calls of {n_rep * len(patches) - len(patches)}")
    b.close()
    return bad


patches = [((torch.from_numpy(cases.rgb_test_patch(t)) - torch.tensor(g["rgb_mean"])) / torch.tensor(g["rgb_std"])).numpy()
           for t in range(cases.RGB_CASE["n_test"])]
total = hunt("golden 784 rows", bank_rows, patches, False)
cent = synth.centroids(768, 256)
rows = synth.patches(5000, 768, seed=5, cent=cent)
qs = [synth.patches(784, 768, seed=6 + i, anomalous_frac=0.01, cent=cent) for i in range(3)]
total += hunt("5000 rows", rows, qs, False)
total += hunt("5000 rows + table", rows, qs, True)
print("TOTAL differing calls:", total)
