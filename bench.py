#!/usr/bin/env python
"""Headline benchmark of the memory-bank anomaly-scoring hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" scores one batch (--batch, default 16) of synthetic 784-patch images (DINO ViT-B/8 shaped, 768-d) against a
200 000 x 768 float32 bank through the full path: distance GEMM + min/argmin, s*/m*/top-3 re-weighting, bilinear
upsample and Gaussian blur, one result set per image (cmdb_score_batch; --batch 1 is the reference's image-at-a-time
call pattern).
  value   patch-NN scores/s with the patch already resident in HBM (device pointer through the C ABI)
  e2e     the same through the C ABI with HOST buffers (pinned patch in, results out), copies inside the timed region
  coreset_select_s   projection + greedy selection of 10 % of the same bank (measured once, outside the K steps)
N > 1 (torchrun, one rank per GPU): the bank is row-sharded, every step runs the five-phase sharded scoring with NCCL
collectives in between (strong scaling: the job is still one 200k bank).
`--impl reference` times the CPU restatement of the reference path (oracle/, the one place it may be executed from
here) on the host cores with a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BANK_ROWS, DIM, P, FMAP, OUT_HW = 200_000, 768, 784, 28, 224
METRIC = "patch-NN scores/sec at 200k x 768 bank"
# dram__bytes_read.sum + dram__bytes_write.sum of one score_gemm_kernel<1> launch at batch 16 (ncu --set full capture,
# profiles/r01_prof_gemm.txt): fp16 bank once (307 MB) + queries + the per-tile spill of the candidate lists through L2
GEMM1_DRAM_BYTES_B16 = 620.37e6 + 340.58e6
WORKLOAD = ("cfg5 headline: score 784-patch images (28x28x768) against an un-subsampled 200000x768 fp32 bank "
            "(min/argmin + s*/m*/top-3 reweight + bilinear 224^2 + blur); coreset 10% of the same bank reported beside")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_bank(rank, world):
    from cmdiad_b200 import Bank, synth
    lo = BANK_ROWS * rank // world
    hi = BANK_ROWS * (rank + 1) // world
    bank = Bank(DIM, hi - lo, device=torch.cuda.current_device(), row_offset=lo)
    cent = synth.centroids(DIM)
    chunk = 25_000
    for c0 in range(0, BANK_ROWS, chunk):  # same global rows whatever the world size
        a, b = max(lo, c0), min(hi, c0 + chunk)
        if a < b:
            rows = synth.patches(chunk, DIM, seed=5000 + c0 // chunk, cent=cent)
            bank.append(rows[a - c0:b - c0])
    return bank


def test_patches(n):
    from cmdiad_b200 import synth
    cent = synth.centroids(DIM)
    return [torch.from_numpy(synth.patches(P, DIM, seed=7000 + i, anomalous_frac=0.01, cent=cent)) for i in range(n)]


def cpu_reference_leg(steps, warmup, bank_rows=BANK_ROWS):
    """reference path restated on the CPU (oracle/restate.py: torch.cdist + min + topk + interpolate + PIL-exact blur),
    all host threads; one step = one image against the full bank"""
    from cmdiad_b200 import synth
    from oracle import restate as O
    cent = synth.centroids(DIM)
    lib = torch.from_numpy(np.concatenate(
        [synth.patches(25_000, DIM, seed=5000 + i, cent=cent) for i in range(bank_rows // 25_000)], 0))
    patches = test_patches(max(1, min(4, steps + warmup)))
    for i in range(warmup):
        O.score_restated(patches[i % len(patches)], lib, (FMAP, FMAP), OUT_HW)
    t0 = time.perf_counter()
    for i in range(steps):
        O.score_restated(patches[i % len(patches)], lib, (FMAP, FMAP), OUT_HW)
    dt = time.perf_counter() - t0
    return P * steps / dt, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one step = one 784-patch image on the host cores (0.3 s on the GPU box's 16 cores): bounded so the arm ends in < 1 min
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 3))
    val, ms = cpu_reference_leg(steps, warm)
    cores = torch.get_num_threads()
    line = {"metric": METRIC, "value": val, "unit": "patch-NN scores/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "bank_rows": BANK_ROWS, "dim": DIM, "patches_per_image": P},
            "cpu_baseline": {"value": val, "unit": "patch-NN scores/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} images x {P} patches against the full {BANK_ROWS}x{DIM} bank"},
            "e2e": {"value": val, "unit": "patch-NN scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    from cmdiad_b200 import _lib as L
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pk = peaks()
    bank = build_bank(rank, world)
    bank.finalize()
    knn_build_s = None
    if world == 1:  # neighbour table of the bank (one-off, like finalize); the re-weighting is then a lookup
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bank.build_knn()
        torch.cuda.synchronize()
        knn_build_s = time.perf_counter() - t0
    bank.set_timing(world == 1)
    st = bank.stream()
    B = args.batch
    imgs = test_patches(max(B, 16))
    n_img = 3  # distinct batches cycled through
    host = [torch.stack([imgs[(k * 5 + i) % len(imgs)] for i in range(B)]).pin_memory() for k in range(n_img)]
    dev = [p.cuda() for p in host]
    dims = (FMAP, FMAP)

    def step(patches):
        if world == 1:
            return bank.score_batch(patches, dims, OUT_HW)
        return bank.score_sharded_batch(patches, dims, OUT_HW, distribute=True)

    def timed(patches, steps, collect_stage=False, pipelined=False):
        """K steps between barriers; device time from CUDA events on the bank's stream, max over ranks.
        pipelined (N = 1): the submit / wait pair with two batches in flight -- every step still copies its inputs from
        the host block and its results back inside the timed region, the copies just overlap the other batch's kernels."""
        stage_ms = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        w0 = time.perf_counter()
        pending = None
        for i in range(steps):
            if pipelined:
                t = bank.score_batch_async(patches[i % len(patches)], dims, OUT_HW)
                if pending is not None:
                    pending.wait()
                pending = t
                continue
            step(patches[i % len(patches)])
            if collect_stage:
                stage_ms.append(bank.timings())
        if pending is not None:
            pending.wait()
        e1.record(st)
        e1.synchronize()
        wall = time.perf_counter() - w0
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), stage_ms

    for i in range(args.warmup):
        step(dev[i % n_img])
        step(host[i % n_img])
    if world == 1:  # single-image latency of the reference's call pattern, reported beside the batch throughput
        one = dev[0][0].contiguous()
        for _ in range(3):
            bank.score(one, dims, OUT_HW)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            bank.score(one, dims, OUT_HW)
        single_ms = (time.perf_counter() - t0) / 20 * 1e3
        single_stage = bank.timings()
    pipe = world == 1 and not args.no_pipeline
    sync_call = None
    if pipe:  # the plain synchronous call, for the per-stage times and as a reference point beside the pipelined numbers
        ms_s, wall_s, stages = timed(dev, max(5, args.steps // 2), collect_stage=True)
        ms_h, wall_h, _ = timed(host, max(5, args.steps // 2))
        n_s = max(5, args.steps // 2)
        sync_call = {"value": B * P * n_s / (max(ms_s, wall_s) * 1e-3), "e2e": B * P * n_s / (max(ms_h, wall_h) * 1e-3),
                     "unit": "patch-NN scores/s", "note": "cmdb_score_batch, one batch at a time (host waits for every batch)"}
        for i in range(3):
            timed(host, 2, pipelined=True)
    with ClockSampler(local) as clk:
        ms_dev, wall_dev, st_dev = timed(dev, args.steps, collect_stage=(world == 1 and not pipe), pipelined=pipe)
        ms_e2e, wall_e2e, _ = timed(host, args.steps, pipelined=pipe)
    if not pipe:
        stages = st_dev
    # results are host-visible when the loop ends, so the event span equals the wall span; report the larger (safer) one
    t_dev, t_e2e = max(ms_dev, wall_dev), max(ms_e2e, wall_e2e)
    value = B * P * args.steps / (t_dev * 1e-3)
    e2e = B * P * args.steps / (t_e2e * 1e-3)

    line = {"metric": METRIC, "value": value, "unit": "patch-NN scores/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32 (fp16 tensor-core pre-filter with an error-bound certificate, exact fp32 re-check, FP32-equivalent "
                     "fp16 hi/lo fallback)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "bank_rows": BANK_ROWS, "dim": DIM, "patches_per_image": P,
                       "images_per_step": B,
                       "call": ("cmdb_score_batch_submit / _wait, two batches in flight" if pipe else
                                "cmdb_score_batch" if world == 1 else "cmdb_score_shard_* phases"),
                       "sharding": "single GPU" if world == 1 else f"bank row-sharded over {world} GPUs, 4 NCCL collectives (MIN/SUM/all-gather) "
                                                                    f"per step, map + device->host of image i on rank i % {world}; host queries: every rank "
                                                                    f"stages 1/{world} of the rows over PCIe, NVLink all-gather",
                       "l2": "inputs larger than L2: the bank streams 0.9 GB (fp16 rows for the GEMM, fp32 rows for the re-weighting) per step vs 126 MB of L2"},
            "e2e": {"value": e2e, "unit": "patch-NN scores/s", "h2d_bytes_per_step": B * P * DIM * 4,
                    "d2h_bytes_per_step": B * (OUT_HW * OUT_HW * 4 + P * 12 + 64)},
            # per step: q_split, GEMM, certified refine, decide, rescan, rescan-finish, GEMM-fallback chain (q_split, GEMM,
            # refine: sized on the device, empty unless many certificates fail), select + neighbour-table lookup (N = 1;
            # re-weighting GEMM chain when sharded), 2 blur kernels (+ pack/unpack/
            # select/merge/final/contrib in the sharded protocol); two timed loops (device-resident and host inputs)
            "gpu_launches": args.steps * 2 * (13 if world == 1 else 18),
            "clocks": clk.summary()}
    if sync_call:
        line["sync_call"] = sync_call
    if world == 1:
        gemm_ms = float(np.mean([s["gemm"] for s in stages]))
        flop = 2.0 * B * P * BANK_ROWS * DIM
        achieved = flop / (gemm_ms * 1e-3) / 1e12
        stats = bank.score_stats()
        # DRAM bytes of one launch from the ncu --set full capture of this workload (profiles/r01_prof_gemm.txt:
        # dram__bytes_read.sum + dram__bytes_write.sum at batch 16); algorithmic bytes = fp16 bank once = 307 MB
        traffic = GEMM1_DRAM_BYTES_B16 if B == 16 else None
        line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": achieved / pk["bf16_tflops"], "traffic": traffic, "kernel": "score_gemm_kernel<1>",
                            "kernel_ms": gemm_ms,
                            "frac_of_sustained": achieved / pk["bf16_sustained"] if pk.get("bf16_sustained") else None,
                            "note": f"algorithmic 2*P*R*D FLOP per launch / CUDA-event time of the kernel on its stream; the "
                                    f"certified pre-filter issues exactly these FLOPs as fp16 tcgen05 MMAs, so peak = dense "
                                    f"16-bit tensor throughput of {pk['source']} (burst, kernel timed alone)"}
        line["prefilter"] = {"mode": stats["mode"], "queries_per_step": stats["queries"],
                             "uncertified_queries_last_step": stats["fallback_queries"],
                             "rescan_pairs_last_step": stats["rescan_pairs"], "gemm_fallback_last_step": stats["gemm_fallback"],
                             "note": "mode 0 = certified hi.hi pre-filter; where the error-bound certificate fails the rows it could not "
                                     "exclude are rescanned exactly (or, for many failures, the queries are redone with the "
                                     "FP32-equivalent 3-term GEMM) inside the same call; results identical to mode 3"}
        line["stage_ms"] = {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}
        line["single_image"] = {"ms_per_image": single_ms, "value": P / (single_ms * 1e-3), "unit": "patch-NN scores/s",
                                "stage_ms": single_stage}
        # re-weighting: lookup in the bank's neighbour table (built once after finalize with the certified GEMM of the bank
        # against itself: 2*R*R*D FLOP)
        line["knn_table"] = {"build_s": knn_build_s, "tflops": 2.0 * BANK_ROWS * BANK_ROWS * DIM / knn_build_s / 1e12,
                             "note": "cmdb_bank_build_knn: exact 3 nearest rows of every bank row (pre-filter GEMM + certificate + "
                                     "exact re-check), one-off per bank; without it the per-batch re-weighting costs 0.11 ms"}
    # coreset selection of 10 % of the same bank (BASELINE.json: "coreset-select seconds"); N > 1: row-sharded loop with
    # the in-kernel NVLink mailbox exchange
    if not args.skip_coreset:
        from sklearn import random_projection
        tr = random_projection.SparseRandomProjection(eps=0.9, random_state=0)
        tr.fit(np.broadcast_to(np.zeros((1, 1)), (BANK_ROWS, DIM)))
        c = tr.components_
        csr = (c.indptr, c.indices, c.data, c.shape[0])
        n_sel = BANK_ROWS // 10
        if world == 1:
            run = lambda n: bank.coreset_select(n, csr, L.CORESET_FP16)
        else:
            from cmdiad_b200 import Comm
            comm = Comm(local, d_proj_max=512)
            run = lambda n: bank.coreset_select_sharded(comm, BANK_ROWS, n, csr, L.CORESET_FP16)
        run(64)  # warm-up (module load, allocations)
        barrier()
        t0 = time.perf_counter()
        idx = run(n_sel)
        torch.cuda.synchronize()
        cs = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(cs, op=dist.ReduceOp.MAX)
            comm.close()
        cs = float(cs)
        d_proj = c.shape[0]
        byts = (n_sel - 1) * BANK_ROWS * d_proj * 2.0
        line["coreset_select_s"] = cs
        line["coreset_roofline"] = {"bound": "hbm", "achieved": byts / cs / 1e9, "peak": pk["hbm_gbs"] * world, "unit": "GB/s",
                                    "frac": byts / cs / 1e9 / (pk["hbm_gbs"] * world), "kernel": "coreset_kernel<__half,3>",
                                    "note": f"(n-1)*N*d'*2 B with N={BANK_ROWS}, d'={d_proj}, n={n_sel}; wall time of the whole "
                                            f"call incl. projection (max over ranks); peak = {world} x HBM; the projected bank "
                                            f"({BANK_ROWS * d_proj * 2 / 1e6:.0f} MB) is pinned in L2 as far as it fits, so "
                                            f"achieved/HBM-peak may exceed 1", "unique": int(len(set(idx.tolist())))}
    if world == 1:
        # the FP32-equivalent 3-term GEMM for every query (CMDB_OPT_PREFILTER_TERMS=3), same results, for comparison
        bank.set_prefilter_terms(3)
        for i in range(3):
            step(dev[i % n_img])
        n_full = max(5, args.steps // 2)
        ms_full, wall_full, st_full = timed(dev, n_full, collect_stage=True)
        bank.set_prefilter_terms(0)
        g3 = float(np.mean([x["gemm"] for x in st_full]))
        line["fp32_equivalent_3term_mode"] = {"value": B * P * n_full / (max(ms_full, wall_full) * 1e-3), "unit": "patch-NN scores/s",
                                              "gemm_ms": g3, "gemm_tflops_algorithmic": flop / (g3 * 1e-3) / 1e12,
                                              "frac_of_tf32_equivalent_peak": flop / (g3 * 1e-3) / 1e12 / (pk["bf16_tflops"] / 2.0),
                                              "note": "CMDB_OPT_PREFILTER_TERMS=3: hi.hi + hi.lo + lo.hi for every query; 3x the "
                                                      "tensor work; identical outputs"}
    if world == 1 and rank == 0 and not args.skip_cpu and not args.skip_coreset:
        # the reference's own coreset path on the same projected bank: its torch loop on this GPU ("reference
        # GPU" line, features.py:401-420 as shipped) and on the host cores, both timed on a bounded number of picks and
        # extrapolated linearly (per-pick cost is constant); sklearn's projection timed on a row sample
        from oracle import restate as O
        z = torch.from_numpy(bank.project(csr))
        picks_gpu, picks_cpu = 200, 12
        O.coreset_torch_literal(z[:4096], 8, "FP16", device="cuda")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref_idx = O.coreset_torch_literal(z, picks_gpu + 1, "FP16", device="cuda")
        torch.cuda.synchronize()
        t_gpu = (time.perf_counter() - t0) / picks_gpu
        t0 = time.perf_counter()
        O.coreset_torch_literal(z, picks_cpu + 1, "FP16", device="cpu")
        t_cpu = (time.perf_counter() - t0) / picks_cpu
        sample = bank.read(0, 20_000)
        from sklearn import random_projection as _rp
        t0 = time.perf_counter()
        _rp.SparseRandomProjection(n_components=csr[3], random_state=0).fit_transform(sample)
        t_proj = (time.perf_counter() - t0) * (BANK_ROWS / 20_000)
        line["coreset_baselines"] = {
            "reference_torch_cuda_loop_s": t_gpu * (n_sel - 1), "reference_torch_cpu_loop_s": t_cpu * (n_sel - 1),
            "sklearn_projection_s": t_proj, "cores": torch.get_num_threads(),
            "first_picks_equal_ours": bool((ref_idx.numpy() == idx[:picks_gpu + 1]).all()),
            "sample": f"torch loop: {picks_gpu} picks on cuda / {picks_cpu} picks on cpu of the same {BANK_ROWS}x{csr[3]} projected bank, "
                      f"extrapolated to {n_sel - 1}; projection: 20000 rows extrapolated to {BANK_ROWS}"}
        del z
    if world == 1 and rank == 0 and not args.skip_cpu:  # after every GPU measurement: it keeps all host cores busy
        v, ms = cpu_reference_leg(3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "patch-NN scores/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"3 images x {P} patches against the full {BANK_ROWS}x{DIM} bank (oracle/restate.py "
                                          f"score_restated: torch.cdist + min + topk + bilinear + blur), {ms:.0f} ms/image"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    # device / pinned tensors go before the bank (whose stream they were used on), the bank before the process group
    torch.cuda.synchronize()
    del dev, host, imgs
    bank.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per step")
    ap.add_argument("--skip-coreset", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="N=1: time the synchronous call instead of submit/wait")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: cmdiad_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    run_ours(args)


if __name__ == "__main__":
    main()
