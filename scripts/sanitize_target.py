"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel on tiny shapes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmdiad_b200 import Bank, synth, upsample_blur  # noqa: E402
from cmdiad_b200 import _lib as L  # noqa: E402
from oracle import restate as O  # noqa: E402

D = 256
lib = synth.patches(1200, D, seed=1, k=32)
b = Bank(D, 1200)
b.append(lib)
m, s, _, _ = b.stats()
b.normalize(m, s)
csr = O.sparse_components(1200, D, 0.9, 0)
for mode in (L.CORESET_FP16, L.CORESET_FP64):
    idx = b.coreset_select(40, csr, mode)
print("coreset ok", idx[:5])
b.gather(np.arange(0, 1200, 2))
b.finalize()
p = np.stack([synth.patches(196, D, seed=5 + i, k=32) for i in range(3)])
for terms in (0, 3, 1):
    b.set_prefilter_terms(terms)
    r = b.score_batch(p, (14, 14), 64, full=True)
b.set_prefilter_terms(0)
# near-duplicate rows: the certificate fails for many queries -> exact rescans (and the GEMM fallback chain runs empty)
t0 = b.score_batch_async(p, (14, 14), 64)
t1 = b.score_batch_async(torch.from_numpy(p).cuda(), (14, 14), 64, full=True)
ra, rb = t0.wait(), t1.wait()
print("async ok", ra[0].s, rb[0].s, b.score_stats())
b.build_knn()   # neighbour table: reweight_cert_kernel in table mode, then the lookup path
rk = b.score_batch(p, (14, 14), 64, full=True)
assert all((rk[i].nn_idx == rb[i].nn_idx).all() and rk[i].s[0] == rb[i].s[0] for i in range(3))
print("knn table ok", rk[0].nn_idx)
b.set_score_impl(L.SCORE_SIMT)
r2 = b.score(p[0], (14, 14), 64)
print("score ok", r[0].s, r2.s)
out, pre, u8 = upsample_blur(np.abs(p[0][:, :1].reshape(14, 14)) + 1, 64)
print("blur ok", out.shape)
b.close()
