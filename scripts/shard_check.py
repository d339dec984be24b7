"""torchrun --nproc-per-node N scripts/shard_check.py : row-sharded scoring == single-GPU scoring, bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmdiad_b200 import Bank, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R, D, P = 30_011, 768, 784
cent = synth.centroids(D)
lib = synth.patches(R, D, seed=1, cent=cent)
lo, hi = R * rank // world, R * (rank + 1) // world
shard = Bank(D, hi - lo, device=local, row_offset=lo)
shard.append(lib[lo:hi])
shard.finalize()
full = Bank(D, R, device=local)
full.append(lib)
full.finalize()
bad = 0
# batch form: 5 images per round of collectives
pb = np.stack([synth.patches(P, D, seed=80 + t, anomalous_frac=0.01, cent=cent) for t in range(5)])
ab = shard.score_sharded_batch(pb, (28, 28), 224, full=True)
bb = full.score_batch(pb, (28, 28), 224, full=True)
for t in range(5):
    same = all((getattr(ab[t], n) == getattr(bb[t], n)).all() for n in ("min_idx", "min_val", "s", "s_idx", "nn_idx", "s_map", "w"))
    bad += int(not same)
    if rank == 0:
        print(f"batch image {t}: sharded == single-GPU: {same}", flush=True)
# distributed finish: rank r returns only images r, r + world, ...
dd = shard.score_sharded_batch(pb, (28, 28), 224, full=True, distribute=True)
for t in range(5):
    mine = t % world == rank
    ok = (dd[t] is not None) == mine and (not mine or all((getattr(dd[t], n) == getattr(bb[t], n)).all() for n in ("min_idx", "s", "s_map", "nn_idx")))
    bad += int(not ok)
for t in range(3):
    patch = synth.patches(P, D, seed=50 + t, anomalous_frac=0.01, cent=cent)
    a = shard.score_sharded(patch, (28, 28), 224, full=True)
    b = full.score(patch, (28, 28), 224, full=True)
    same = ((a.min_idx == b.min_idx).all() and (a.min_val == b.min_val).all() and a.s[0] == b.s[0]
            and int(a.s_idx[0]) == int(b.s_idx[0]) and (a.nn_idx == b.nn_idx).all() and (a.s_map == b.s_map).all())
    bad += int(not same)
    if rank == 0:
        print(f"image {t}: sharded == single-GPU: {same}; s={a.s[0]:.6f}/{b.s[0]:.6f} nn={a.nn_idx}/{b.nn_idx}", flush=True)
t = torch.tensor([bad], device="cuda")
dist.all_reduce(t)
if rank == 0:
    print("SHARD CHECK", "OK" if int(t) == 0 else f"FAILED on {int(t)} rank-images", flush=True)
dist.destroy_process_group()
