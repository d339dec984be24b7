"""Short workload for `ncu --set full` captures: a few scoring calls on the 200k x 768 bank and a short coreset run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmdiad_b200 import _lib as L  # noqa: E402
import bench  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "score"
torch.cuda.set_device(0)
bank = bench.build_bank(0, bench.BANK_ROWS, 0)
if what == "score":
    bank.finalize()
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    patches = torch.stack(bench.test_patches(B)).cuda()
    bank.build_knn()
    bank.score_batch(patches, (28, 28), 224)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()   # ncu --profile-from-start off: only the calls below are visible to the profiler
    for i in range(3):
        bank.score_batch(patches, (28, 28), 224)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    from sklearn import random_projection
    tr = random_projection.SparseRandomProjection(eps=0.9, random_state=0)
    tr.fit(np.broadcast_to(np.zeros((1, 1)), (bench.BANK_ROWS, bench.DIM)))
    c = tr.components_
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    idx = bank.coreset_select(n, (c.indptr, c.indices, c.data, c.shape[0]), L.CORESET_FP16)
    print("picked", len(idx))
bank.close()
