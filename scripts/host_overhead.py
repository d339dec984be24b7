"""diagnostic: host-side cost of one pipelined scoring call.  With a tiny bank the GPU work per call is far below the
host's enqueue + result handling time, so the steady-state period of the submit / wait loop IS the host cost."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.cuda.set_device(0)
bank = bench.build_bank(0, rows, 0)
bank.finalize()
bank.build_knn()
for B in (16, 1):
    imgs = torch.stack(bench.test_patches(B)).cuda()
    t_sub = t_wait = 0.0

    depth = int(os.environ.get("DEPTH", "3"))

    def run(k):
        global t_sub, t_wait
        pending = []
        for _ in range(k):
            t0 = time.perf_counter()
            pending.append(bank.score_batch_async(imgs, (28, 28), 224))
            t1 = time.perf_counter()
            if len(pending) == depth:
                pending.pop(0).wait()
            t2 = time.perf_counter()
            t_sub += t1 - t0
            t_wait += t2 - t1
        while pending:
            pending.pop(0).wait()

    run(20)
    torch.cuda.synchronize()
    t_sub = t_wait = 0.0
    n = 200
    t0 = time.perf_counter()
    run(n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"rows {rows} B {B} depth {depth}: period {dt / n * 1e3:.3f} ms per call; submit {t_sub / n * 1e3:.3f} ms, wait (incl. blocking) {t_wait / n * 1e3:.3f} ms")
bank.close()
