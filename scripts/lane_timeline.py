"""diagnostic: where the stages of two consecutive pipelined batches sit on a common time base (do the tail kernels of batch
k really overlap the distance GEMM of batch k + 1 on the other lane?).  CMDB_OPT_TIMING = 2 + cmdb_debug_lane_timeline."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cmdiad_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rows = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BANK_ROWS
torch.cuda.set_device(0)
bank = bench.build_bank(0, rows, 0)
bank.finalize()
bank.build_knn()
imgs = torch.stack(bench.test_patches(B)).cuda()
lib = bank._lib
lib.cmdb_debug_lane_timeline.restype = ctypes.c_int
lib.cmdb_debug_lane_timeline.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]


def run(k):
    pending = None
    for _ in range(k):
        tk = bank.score_batch_async(imgs, (28, 28), 224)
        if pending is not None:
            pending.wait()
        pending = tk
    pending.wait()


run(6)
torch.cuda.synchronize()
L.check(lib.cmdb_bank_set_option(bank._h, L.OPT_TIMING, 2))
n = 8
run(n)
names = list(L.T_STAGES) + ["done", "cert0", "cert1", "rescan1", "tier2_1"]
arr = (ctypes.c_float * (2 * len(names)))()
L.check(lib.cmdb_debug_lane_timeline(bank._h, arr))
t = [[arr[l * len(names) + i] for i in range(len(names))] for l in range(2)]
first = 0 if t[0][0] < t[1][0] else 1   # lane of batch n - 2
for l in (first, 1 - first):
    print(f"lane {l} (batch {'n-2' if l == first else 'n-1'}):", "  ".join(f"{nm}@{x:.3f}" for nm, x in zip(names, t[l])))
a, b = t[first], t[1 - first]
print(f"batch n-2: gemm {a[2] - a[1]:.3f} ms, refine {a[3] - a[2]:.3f}, map {a[4] - a[3]:.3f}, reweight {a[5] - a[4]:.3f}, out {a[6] - a[5]:.3f}")
print(f"batch n-1: gemm {b[2] - b[1]:.3f} ms, refine {b[3] - b[2]:.3f}, map {b[4] - b[3]:.3f}, reweight {b[5] - b[4]:.3f}, out {b[6] - b[5]:.3f}")
print(f"GEMM(n-1) starts {b[1] - a[2]:+.3f} ms after GEMM(n-2) ends; tail(n-2) ends {a[5] - b[1]:+.3f} ms after GEMM(n-1) starts "
      f"and {a[5] - b[2]:+.3f} ms relative to its end; period (end of refine to end of refine) {b[3] - a[3]:.3f} ms")
for nm, x in (("n-2", a), ("n-1", b)):
    print(f"batch {nm} refine stage: memsets {x[7] - x[2]:.3f}, certificate {x[8] - x[7]:.3f}, rescan {x[9] - x[8]:.3f}, "
          f"counters copy + tier-2 launches {x[10] - x[9]:.3f} ms")
bank.close()
