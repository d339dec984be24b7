"""Diagnostics run on the GPU box (not a test, not a benchmark): prints mismatch statistics and rough timings for every
kernel so one gpurun call tells as much as possible."""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmdiad_b200 import Bank, coreset_rownorms, synth, upsample_blur  # noqa: E402
from cmdiad_b200 import _lib as L  # noqa: E402
from oracle import restate as O  # noqa: E402


def section(name):
    def deco(fn):
        def run():
            print(f"\n=== {name} ===", flush=True)
            t = time.time()
            try:
                fn()
            except Exception:
                traceback.print_exc()
            print(f"--- {name}: {time.time() - t:.1f}s", flush=True)
        return run
    return deco


def ev_time(stream, fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


@section("rownorms vs oracle vs torch-cuda")
def s_rownorms():
    for d in (64, 100, 128, 129, 130, 131, 221, 301, 375):
        g = np.random.Generator(np.random.PCG64(d))
        z = (g.standard_normal((4099, d)) * 1.7).astype(np.float16)
        got = coreset_rownorms(z, z[17])
        orc = O.rownorms_restated(z, z[17])
        zt = torch.from_numpy(z).cuda()
        ref = torch.linalg.norm(zt - zt[17:18], dim=1, keepdims=True).cpu().numpy()[:, 0]
        z64 = g.standard_normal((4099, d))
        got64 = coreset_rownorms(z64, z64[5])
        zt64 = torch.from_numpy(z64).cuda()
        ref64 = torch.linalg.norm(zt64 - zt64[5:6], dim=1, keepdims=True).cpu().numpy()[:, 0]
        import ctypes
        nf = np.zeros(4099)
        l5 = np.ascontiguousarray(z64[5])
        O.lib().oracle_rownorms_fp64_nofma(O._p(z64), O._p(l5), ctypes.c_int64(4099), ctypes.c_int(d), O._p(nf))
        print(f"   fp64 nofma-oracle != torch: {(nf != ref64).sum()}")
        print(f"d={d}: fp16 kernel!=oracle {(got.view(np.uint16) != orc.view(np.uint16)).sum()}, kernel!=torch "
              f"{(got.view(np.uint16) != ref.view(np.uint16)).sum()}, oracle!=torch {(orc.view(np.uint16) != ref.view(np.uint16)).sum()}"
              f" | fp64 kernel!=oracle {(got64 != O.rownorms_restated(z64, z64[5])).sum()}, kernel!=torch {(got64 != ref64).sum()}")


@section("projection")
def s_proj():
    g = np.random.Generator(np.random.PCG64(1))
    x = g.standard_normal((3000, 768), dtype=np.float32)
    csr = O.sparse_components(3000, 768, 0.9, 0)
    b = Bank(768, 3000)
    b.append(x)
    z = b.project(csr)
    print("mismatch vs oracle:", (z != O.project_restated(x, *csr)).sum(), "of", z.size)
    b.close()


@section("coreset small parity")
def s_coreset_small():
    lib = np.concatenate(synth.image_bank(10, 784, 768, 11), 0)
    lib = (lib - lib.mean()) / lib.std()
    csr = O.sparse_components(lib.shape[0], 768, 0.9, 0)
    z = O.project_restated(lib, *csr)
    b = Bank(768, lib.shape[0])
    b.append(lib)
    n = 784
    for mode, name in ((L.CORESET_FP64, "TF32"), (L.CORESET_FP16, "FP16")):
        t = time.time()
        idx = b.coreset_select(n, csr, mode)
        dt = time.time() - t
        ref = O.coreset_restated(z, n, name)
        lit = O.coreset_torch_literal(torch.from_numpy(z), n, name, device="cuda").numpy()
        ne, nl = np.nonzero(idx != ref)[0], np.nonzero(idx != lit)[0]
        print(f"{name}: {dt*1e3:.1f} ms; vs oracle first-div {ne[:1]}, vs torch-cuda literal first-div {nl[:1]}, "
              f"oracle vs literal first-div {np.nonzero(ref != lit)[0][:1]}")
    b.close()


@section("coreset 200k x 768 -> 301, 10% timing")
def s_coreset_big():
    N, D = 200_000, 768
    cent = synth.centroids(D)
    b = Bank(D, N)
    for i in range(8):
        b.append(synth.patches(N // 8, D, seed=100 + i, cent=cent))
    mean, std, _, _ = b.stats()
    b.normalize(mean, std)
    csr = O.sparse_components(N, D, 0.9, 0)
    for n in (2000, 20000):
        torch.cuda.synchronize()
        t = time.time()
        idx = b.coreset_select(n, csr, L.CORESET_FP16)
        dt = time.time() - t
        print(f"FP16 n={n}: {dt:.3f} s total, {(dt) / (n - 1) * 1e6:.2f} us/pick incl. projection; d'={csr[3]}; "
              f"algorithmic {(n-1)*N*csr[3]*2/dt/1e9:.0f} GB/s; unique {len(set(idx.tolist()))}")
    t = time.time()
    z = b.project(csr, 0, 1000)
    print("projection of 1000 rows (incl. D2H)", time.time() - t)
    t = time.time()
    idx = b.coreset_select(500, csr, L.CORESET_FP64)
    dt = time.time() - t
    print(f"FP64 n=500: {dt:.3f} s")
    b.close()


@section("score small parity (both impls)")
def s_score_small():
    cent = synth.centroids(768, 256)
    lib = synth.patches(5000, 768, seed=1, cent=cent)
    patch = synth.patches(784, 768, seed=2, anomalous_frac=0.01, cent=cent)
    ref = O.score_restated(patch, lib, (28, 28), 224)
    for impl, name in ((L.SCORE_SIMT, "simt"), (L.SCORE_TCGEN05, "tcgen05")):
        try:
            b = Bank(768, 5000)
            b.append(lib)
            b.finalize()
            b.set_score_impl(impl)
            r = b.score(patch, (28, 28), 224, full=True)
            print(f"{name}: argmin mismatches {(r.min_idx != ref['min_idx']).sum()}, min_val max rel err "
                  f"{np.max(np.abs(r.min_val - ref['min_val']) / ref['min_val']):.2e}, s {r.s[0]:.6f} vs {ref['s']:.6f}, "
                  f"s_idx {r.s_idx[0]} vs {ref['s_idx']}, nn {r.nn_idx} vs {ref['nn_idx']}, knn {r.m_star_knn} vs {ref['m_star_knn']}, "
                  f"pre max rel {np.max(np.abs(r.s_map_pre - ref['s_map_pre']) / ref['s_map_pre']):.2e}, "
                  f"u8 mismatches {(r.s_map_u8 != ref['s_map_u8']).sum()}, blur mismatches {(r.s_map != ref['s_map']).sum()}")
            b.close()
        except Exception:
            traceback.print_exc()


@section("upsample+blur")
def s_blur():
    g = np.random.Generator(np.random.PCG64(12))
    m = (np.abs(g.standard_normal((28, 28))) * 7 + 3).astype(np.float32)
    out, pre, u8 = upsample_blur(m, 224)
    ref, ref_u8 = O.knn_blur_restated(pre)
    print("pre mismatches", (pre != O.bilinear_restated(m, 224)).sum(), "u8", (u8 != ref_u8).sum(), "blur", (out != ref).sum())


@section("score 200k x 768 timing")
def s_score_big():
    R, D, P = 200_000, 768, 784
    cent = synth.centroids(D)
    b = Bank(D, R)
    for i in range(8):
        b.append(synth.patches(R // 8, D, seed=300 + i, cent=cent))
    t = time.time()
    b.finalize()
    print("finalize", time.time() - t)
    patch = torch.from_numpy(synth.patches(P, D, seed=400, anomalous_frac=0.01, cent=cent)).cuda()
    st = b.stream()
    for impl, name in ((L.SCORE_TCGEN05, "tcgen05"), (L.SCORE_SIMT, "simt")):
        b.set_score_impl(impl)
        best, med = ev_time(st, lambda: b.score(patch, (28, 28), 224), iters=5, warm=2)
        flop = 2.0 * P * R * D
        print(f"{name}: whole call best {best:.3f} ms median {med:.3f} ms -> {P / med * 1e3:.0f} patches/s, "
              f"{flop / med / 1e9:.1f} algorithmic TFLOP/s")
    b.set_score_impl(L.SCORE_TCGEN05)
    r1 = b.score(patch, (28, 28), 224)
    b.set_score_impl(L.SCORE_SIMT)
    r2 = b.score(patch, (28, 28), 224)
    print("tcgen05 vs simt: idx mismatches", (r1.min_idx != r2.min_idx).sum(), "val mismatches", (r1.min_val != r2.min_val).sum(),
          "s", r1.s[0], r2.s[0])
    b.close()


@section("GEMM mode / batch sweep on the 200k x 768 bank (stage times from the library's CUDA events)")
def s_gemm_sweep():
    import bench
    b = bench.build_bank(0, 1)
    b.finalize()
    b.set_timing(True)
    imgs = torch.stack(bench.test_patches(16)).cuda()
    for mode in (0, 3):
        b.set_prefilter_terms(mode)
        for B in (1, 2, 4, 8, 16):
            x = imgs[:B].contiguous()
            for _ in range(3):
                b.score_batch(x, (28, 28), 224)
            ts = []
            for _ in range(6):
                b.score_batch(x, (28, 28), 224)
                ts.append(b.timings())
            med = {k: float(np.median([t[k] for t in ts])) for k in ts[0]}
            print(f"mode {mode} B={B:2d} gemm {med['gemm']:.3f} refine {med['refine']:.3f} reweight {med['reweight']:.3f} "
                  f"map {med['map']:.3f} out {med['out']:.3f} stage_in {med['stage_in']:.3f}  stats {b.score_stats()}", flush=True)
    b.close()


@section("pipelined submit/wait: device vs pinned-host inputs, raw copy rates")
def s_pipeline():
    import bench
    b = bench.build_bank(0, 1)
    b.finalize()
    imgs = bench.test_patches(16)
    host = [torch.stack([imgs[(k * 5 + i) % 16] for i in range(16)]).pin_memory() for k in range(3)]
    dev = [h.cuda() for h in host]
    # raw copy rates of one batch
    cs = torch.cuda.Stream()
    tgt = torch.empty_like(dev[0])
    back = torch.empty(16, 224, 224, dtype=torch.float32).pin_memory()
    src_maps = torch.empty(16, 224, 224, dtype=torch.float32, device="cuda")
    for name, fn in (("H2D 38.5 MB", lambda: tgt.copy_(host[0], non_blocking=True)),
                     ("D2H 3.2 MB", lambda: back.copy_(src_maps, non_blocking=True))):
        with torch.cuda.stream(cs):
            best, med = ev_time(cs, fn, iters=8, warm=2)
        print(f"{name}: best {best:.3f} ms median {med:.3f} ms")

    def loop(xs, n, pipelined):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pending = None
        for i in range(n):
            if pipelined:
                t = b.score_batch_async(xs[i % 3], (28, 28), 224)
                if pending is not None:
                    pending.wait()
                pending = t
            else:
                b.score_batch(xs[i % 3], (28, 28), 224)
        if pending is not None:
            pending.wait()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    for name, xs in (("device", dev), ("host", host)):
        for pipe in (False, True):
            loop(xs, 5, pipe)
            print(f"{name:6s} pipelined={pipe}: {loop(xs, 30, pipe):.3f} ms/step", flush=True)
    # host-side cost of one submit and one wait while the GPU is idle
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t = b.score_batch_async(dev[0], (28, 28), 224)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    t.wait()
    t3 = time.perf_counter()
    print(f"host time: submit {1e3 * (t1 - t0):.3f} ms, wait (results already on the host) {1e3 * (t3 - t2):.3f} ms")
    b.close()


@section("big banks: 1M x 768 (coreset + scoring) and 627k x 1152 (cfg2 XYZ scoring)")
def s_big():
    for (R, D, P, fm) in ((1_000_000, 768, 784, 28), (627_200, 1152, 3136, 56)):
        cent = synth.centroids(D)
        b = Bank(D, R)
        t = time.time()
        for i in range(R // 50_000 + 1):
            n = min(50_000, R - i * 50_000)
            if n > 0:
                b.append(synth.patches(n, D, seed=800 + i, cent=cent))
        print(f"R={R} D={D}: filled in {time.time() - t:.1f}s")
        mean, std, _, _ = b.stats()
        b.normalize(mean, std)
        if D == 768:
            csr = O.sparse_components(R, D, 0.9, 0)
            b.coreset_select(64, csr, L.CORESET_FP16)
            t = time.time()
            idx = b.coreset_select(2000, csr, L.CORESET_FP16)
            dt = time.time() - t
            print(f"  coreset d'={csr[3]} 2000 picks: {dt:.3f} s ({dt / 1999 * 1e6:.1f} us/pick, {1999 * R * csr[3] * 2 / dt / 1e9:.0f} GB/s algorithmic), unique {len(set(idx.tolist()))}")
        t = time.time()
        b.finalize()
        print(f"  finalize {time.time() - t:.3f}s")
        nb = 4
        patches = torch.from_numpy(np.stack([synth.patches(P, D, seed=900 + i, anomalous_frac=0.01, cent=cent) for i in range(nb)])).cuda()
        patches = (patches - mean) / std
        r = b.score_batch(patches, (fm, fm), 224)
        st = b.stream()
        best, med = ev_time(st, lambda: b.score_batch(patches, (fm, fm), 224), iters=3, warm=1)
        print(f"  score batch of {nb}x{P}: {med:.2f} ms -> {nb * P / med * 1e3:.0f} patches/s, {2.0 * nb * P * R * D / med / 1e9:.0f} algorithmic TFLOP/s")
        # brute-force check of image 0 on the GPU (chunks of bank rows to bound memory)
        q = patches[0]
        bank_dev = None
        best_v = torch.full((P,), float("inf"), device="cuda")
        best_i = torch.zeros(P, dtype=torch.long, device="cuda")
        for c0 in range(0, R, 100_000):
            rows = b.read(c0, min(100_000, R - c0)).cuda()
            dd = torch.cdist(q, rows, compute_mode="donot_use_mm_for_euclid_dist")
            v, i = dd.min(1)
            upd = v < best_v
            best_v = torch.where(upd, v, best_v)
            best_i = torch.where(upd, i + c0, best_i)
        mism = (torch.from_numpy(r[0].min_idx).cuda() != best_i).sum().item()
        rel = (torch.from_numpy(r[0].min_val).cuda() - best_v).abs().max().item() / best_v.max().item()
        print(f"  vs GPU brute force: argmin mismatches {mism}/{P}, max rel err {rel:.2e}")
        b.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.__version__)
    which = sys.argv[1:] or ["rownorms", "proj", "blur", "score_small", "coreset_small", "score_big", "coreset_big"]
    table = dict(big=s_big, rownorms=s_rownorms, proj=s_proj, coreset_small=s_coreset_small, coreset_big=s_coreset_big,
                 score_small=s_score_small, blur=s_blur, score_big=s_score_big, gemm_sweep=s_gemm_sweep, pipeline=s_pipeline)
    for w in which:
        table[w]()
