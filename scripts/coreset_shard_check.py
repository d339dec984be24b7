"""torchrun --nproc-per-node N scripts/coreset_shard_check.py : row-sharded coreset == single-GPU coreset, both dtype modes.

Prints one line per case and CORESET SHARD CHECK OK / FAILED (rank 0); exit code 1 on failure."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmdiad_b200 import Bank, Comm, synth  # noqa: E402
from cmdiad_b200 import _lib as L  # noqa: E402
from cmdiad_b200.sharding import shard_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = Comm(local, d_proj_max=512)
ok_all = True
for (N, D, n, seed, mode) in ((7841, 768, 500, 3, L.CORESET_FP16), (7841, 768, 500, 3, L.CORESET_FP64),
                              (30011, 768, 300, 4, L.CORESET_FP16), (30011, 1152, 300, 6, L.CORESET_FP64),
                              (200_000, 768, 2000, 5, L.CORESET_FP16), (200_000, 768, 500, 5, L.CORESET_FP64)):
    from sklearn import random_projection
    tr = random_projection.SparseRandomProjection(eps=0.9, random_state=0)
    tr.fit(np.broadcast_to(np.zeros((1, 1)), (N, D)))
    c = tr.components_
    csr = (c.indptr, c.indices, c.data, c.shape[0])
    cent = synth.centroids(D)
    lib = np.concatenate([synth.patches(min(25000, N - o), D, seed=seed * 100 + o // 25000, cent=cent) for o in range(0, N, 25000)], 0)
    lib = ((lib - lib.mean()) / lib.std()).astype(np.float32)
    lo, hi = shard_range(N, rank, world)
    shard = Bank(D, hi - lo, device=local, row_offset=lo)
    shard.append(lib[lo:hi])
    shard.coreset_select_sharded(comm, N, 16, csr, mode)  # warm-up: module load, peer-buffer growth
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    idx_sh = shard.coreset_select_sharded(comm, N, n, csr, mode)
    t_sh = time.perf_counter() - t0
    full = Bank(D, N, device=local)
    full.append(lib)
    full.coreset_select(16, csr, mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx_1 = full.coreset_select(n, csr, mode)
    t_1 = time.perf_counter() - t0
    same = bool((idx_sh == idx_1).all())
    t = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_all &= bool(int(t))
    if rank == 0:
        first = np.nonzero(idx_sh != idx_1)[0][:1]
        print(f"N={N} D={D} d'={csr[3]} n={n} {'FP16' if mode == L.CORESET_FP16 else 'FP64'}: sharded({world}) == single: "
              f"{bool(int(t))} (first diff {first}); {t_sh * 1e6 / n:.1f} us/pick sharded vs {t_1 * 1e6 / n:.1f} us/pick single",
              flush=True)
    shard.close()
    full.close()
if rank == 0:
    print("CORESET SHARD CHECK", "OK" if ok_all else "FAILED", f"(world {world})", flush=True)
comm.close()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
