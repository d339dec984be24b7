"""Times the five phases of the row-sharded scoring protocol with CUDA events (torchrun --nproc-per-node N; N = 1 works)."""
import ctypes
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cmdiad_b200 import _lib as L  # noqa: E402
from cmdiad_b200.bank import Bank, _ptr  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29555")
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
bank = bench.build_bank(rank, world)
bank.finalize()
B, P, D = 16, bench.P, bench.DIM
x = torch.stack(bench.test_patches(B)).cuda()
dev = torch.device("cuda", local)
st = bank.stream()
lib = bank._lib


def one_round(timed):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    k = 0

    def mark():
        nonlocal k
        if timed:
            ev[k].record(st)
        k += 1
    with torch.cuda.stream(st):
        mark()
        keys = torch.empty(B * P, dtype=torch.int64, device=dev)
        L.check(lib.cmdb_score_shard_min(bank._h, _ptr(x), B, P, 1, 224, _ptr(keys)))
        mark()
        dist.all_reduce(keys, op=dist.ReduceOp.MIN)
        mark()
        m_star = torch.empty(B * D, dtype=torch.float32, device=dev)
        L.check(lib.cmdb_score_shard_select(bank._h, _ptr(keys), B, P, _ptr(m_star)))
        mark()
        dist.all_reduce(m_star, op=dist.ReduceOp.SUM)
        mark()
        top = torch.empty(B * 3, dtype=torch.int64, device=dev)
        L.check(lib.cmdb_score_shard_topk(bank._h, _ptr(m_star), B, P, _ptr(top)))
        mark()
        gathered = torch.empty(world * B * 3, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, top)
        mark()
        nn_rows = torch.empty(B * 3 * D, dtype=torch.float32, device=dev)
        L.check(lib.cmdb_score_shard_nn(bank._h, _ptr(gathered), world, B, _ptr(nn_rows)))
        mark()
        dist.all_reduce(nn_rows, op=dist.ReduceOp.SUM)
        mark()
        res, outs, _ = Bank._alloc_out(B, P, 224, False)
        L.check(lib.cmdb_score_shard_finish(bank._h, _ptr(nn_rows), B, P, 28, 28, 224, rank % world, world, outs))
        mark()
    torch.cuda.synchronize()
    if timed:
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(9)]


for _ in range(5):
    one_round(False)
ts = np.array([one_round(True) for _ in range(10)])
names = ["shard_min", "allreduce MIN keys", "select", "allreduce SUM m_star", "topk (re-weighting)", "all_gather keys",
         "merge + contrib", "allreduce SUM rows", "finish (final, blur, D2H)"]
if rank == 0:
    med = np.median(ts, 0)
    for n, t in zip(names, med):
        print(f"{n:28s} {t:7.3f} ms")
    print(f"{'sum':28s} {med.sum():7.3f} ms")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        bank.score_sharded_batch(x, (28, 28), 224, distribute=True)
    torch.cuda.synchronize()
    print(f"score_sharded_batch wall: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/step")
else:
    for _ in range(20):
        bank.score_sharded_batch(x, (28, 28), 224, distribute=True)
torch.cuda.synchronize()
del x
bank.close()
dist.barrier()
dist.destroy_process_group()
