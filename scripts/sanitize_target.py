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
# round 2: query normalisation + late-fusion head + device-side result store + pixel metrics (radix sort) on two banks
from cmdiad_b200 import metrics  # noqa: E402
from cmdiad_b200.fusion import LateFusion  # noqa: E402
b.set_score_impl(L.SCORE_TCGEN05)
b2 = Bank(D, 600)
b2.append(lib[:600] * 1.5)
b2.finalize()
b2.build_knn()
b.set_query_norm(0.1, 1.3, True)
b2.set_query_norm(-0.2, 0.9, True)
fus = LateFusion([b, b2], [1.0, 0.1], [1.0, 0.1], [0.8, 0.9], [2.0], [0.7, 1.1], [3.0])
fus.eval_reserve(6, 64)
for _ in range(2):
    rf = fus.score_batch([p, p * 0.5], [(14, 14), (14, 14)], 64, keep_on_device=True, want_patch=True)
gts = [np.zeros((64, 64), np.float32) for _ in range(6)]
gts[1][10:20, 10:20] = 1
gts[4][40:44, 30:50] = 1
ev = metrics.device_pixel_metrics(fus, gts)
print("fused + eval ok", rf.s, ev["pixel_rocauc"], ev["au_pro"])
b.close()
b2.close()
