"""diagnostic: batch-32 vs batch-16 results on the 200k bank, both against the exact scan (cmdb_debug_exact_min)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.cuda.set_device(0)
bank = bench.build_bank(0, bench.BANK_ROWS, 0)
bank.finalize()
bank.build_knn()
imgs = bench.test_patches(32)
x32 = torch.stack(imgs).pin_memory()
r32 = bank.score_batch(x32, (28, 28), 224, full=True)
r16a = bank.score_batch(x32[:16].contiguous(), (28, 28), 224, full=True)
r32d = bank.score_batch(x32.cuda(), (28, 28), 224, full=True)
for name, a, b, n in (("host32 vs host16", r32, r16a, 16), ("host32 vs dev32", r32, r32d, 32)):
    bad = [(i, int((a[i].min_val != b[i].min_val).sum()), int((a[i].min_idx != b[i].min_idx).sum())) for i in range(n)
           if not ((a[i].min_val == b[i].min_val).all() and (a[i].min_idx == b[i].min_idx).all())]
    print(name, "differing images:", bad)
for i in (0, 1, 17):
    P = 784
    ex_val, ex_idx = np.empty(P, np.float32), np.empty(P, np.int64)
    q = np.ascontiguousarray(imgs[i].numpy())
    rc = bank._lib.cmdb_debug_exact_min(bank._h, q.ctypes.data, P, ex_val.ctypes.data, ex_idx.ctypes.data)
    for name, r in (("host32", r32), ("dev32", r32d)) + ((("host16", r16a),) if i < 16 else ()):
        dv = np.abs(r[i].min_val - ex_val) / ex_val
        print(f"image {i} {name}: idx mismatches vs exact scan {int((r[i].min_idx != ex_idx).sum())}, max rel val diff {dv.max():.3e} at patch {int(dv.argmax())}"
              f" ours {r[i].min_val[dv.argmax()]:.6f} idx {r[i].min_idx[dv.argmax()]} exact {ex_val[dv.argmax()]:.6f} idx {ex_idx[dv.argmax()]}")
print("stats", bank.score_stats())
bank.close()
