/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's greedy k-center coreset loop.
 * Nothing in the product path (cmdiad_b200/) may link, import or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Follows /root/reference/feature_extractors/features.py:372-425 (get_coreset_idx_randomp after the projection):
 *   :372-378  pick 0 := row 0, min_distances = ||z - z[0]||_2   (float64, on the host)
 *   :388-391  FP16 mode: z, last_item, min_distances -> .half()  ('TF32' mode: data stays float64)
 *   :401-420  n-1 times: d = ||z - last||_2 ; min_d = minimum(d, min_d) ; sel = argmax(min_d) (ties -> lowest
 *             index) ; last = z[sel] ; min_d[sel] = 0
 *
 * The arithmetic the reference leaves to a third-party dependency is torch (README pins 2.2.0; this image has
 * 2.11.0+cu128).  The loop runs on CUDA in the reference (features.py:397-399), so the summation order restated here
 * is the one ATen's CUDA reduction uses for `torch.linalg.norm(x[N,d], dim=1)` with x contiguous
 * (ATen/native/cuda/Reduce.cuh in the torch 2.11 wheel: setReduceConfig :1033-1179, input_vectorized_thread_reduce_impl
 * :500-559, thread_reduce_impl :561-632, block_x_reduce; NormTwoOps in ATen/native/SharedReduceOps.h:378-405):
 *   - one 32-lane warp per output row (block 32x16), accumulation type float (half input) / double (double input);
 *   - d >= 128: "vectorize along input" with 4-element aligned vectors.  Row i of a contiguous [N,d] tensor starts
 *     shift = (i*d) % 4 elements past a vector boundary.  If shift > 0, lanes shift..3 first take the 4-shift head
 *     elements into accumulator 0; lane x then walks aligned vectors x, x+32, ... putting element j of each vector
 *     into accumulator j; the < 4 tail elements go to lanes 0.. (accumulator 0) afterwards;
 *   - d < 128: lane x walks elements x, x+W, x+2W, ... round-robin into 4 accumulators (W = lanes per row);
 *   - every step is acc = fma(v, v, acc); per lane ((a0+a1)+a2)+a3; then a shfl_down tree with offsets 1,2,4,...;
 *   - sqrt (IEEE), and for half output one round-to-nearest-even float->half conversion.
 * `z - last` is rounded to the storage type before the norm (half: float subtract, then RNE to half).
 * This order is pinned on the GPU box against torch itself by tests/test_coreset_gpu.py (free-running index
 * equality with the literal torch restatement of the reference loop in oracle/restate.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 half_t;

#define VEC 4
#define MAXLANES 32

/* ---- canonical-order plan: for each of the 4 row alignments, which (lane, accumulator) owns element e ---- */
typedef struct {
    int d, lanes, vectorized;
    uint8_t *lane[VEC]; /* [shift][e] */
    uint8_t *acc[VEC];
} plan_t;

static int last_pow2(int n) {
    int p = 1;
    while (p * 2 <= n) p *= 2;
    return p;
}

static void plan_init(plan_t *p, int d) {
    p->d = d;
    p->vectorized = d >= 128; /* Reduce.cuh:1099 */
    int dim0 = p->vectorized ? d / VEC : d;
    int w = last_pow2(dim0);
    p->lanes = w < MAXLANES ? w : MAXLANES; /* Reduce.cuh:100-108 with dim1 >= 512 */
    for (int s = 0; s < VEC; ++s) {
        p->lane[s] = (uint8_t *)malloc((size_t)d);
        p->acc[s] = (uint8_t *)malloc((size_t)d);
        if (!p->vectorized) {
            for (int e = 0; e < d; ++e) {
                p->lane[s][e] = (uint8_t)(e % p->lanes);
                p->acc[s][e] = (uint8_t)((e / p->lanes) % VEC);
            }
            continue;
        }
        int base = 0, end = d;
        if (s > 0) { /* Reduce.cuh:508-517 */
            for (int x = s; x < VEC; ++x) {
                p->lane[s][x - s] = (uint8_t)x;
                p->acc[s][x - s] = 0;
            }
            base = VEC - s;
            end = d + s - VEC;
        }
        int tail_start = end - end % VEC;
        for (int q = 0; q < tail_start; ++q) { /* Reduce.cuh:534-541 */
            int v = q / VEC;
            p->lane[s][base + q] = (uint8_t)(v % p->lanes);
            p->acc[s][base + q] = (uint8_t)(q % VEC);
        }
        for (int q = tail_start; q < end; ++q) { /* Reduce.cuh:544-551 */
            p->lane[s][base + q] = (uint8_t)(q - tail_start);
            p->acc[s][base + q] = 0;
        }
    }
}

static void plan_free(plan_t *p) {
    for (int s = 0; s < VEC; ++s) {
        free(p->lane[s]);
        free(p->acc[s]);
    }
}

/* ||a - b||_2 of one half row in the canonical order; returns the half-rounded norm */
static half_t rownorm_f16(const plan_t *p, const half_t *a, const half_t *b, int shift) {
    float acc[MAXLANES][VEC];
    memset(acc, 0, sizeof(acc));
    const uint8_t *ln = p->lane[shift], *ac = p->acc[shift];
    for (int e = 0; e < p->d; ++e) {
        half_t diff = (half_t)((float)a[e] - (float)b[e]); /* sub in opmath float, stored as half */
        float v = (float)diff;
        acc[ln[e]][ac[e]] = fmaf(v, v, acc[ln[e]][ac[e]]);
    }
    float lanev[MAXLANES];
    for (int l = 0; l < p->lanes; ++l) lanev[l] = ((acc[l][0] + acc[l][1]) + acc[l][2]) + acc[l][3];
    for (int off = 1; off < p->lanes; off <<= 1)
        for (int l = 0; l + off < p->lanes; l += 2 * off) lanev[l] = lanev[l] + lanev[l + off];
    return (half_t)sqrtf(lanev[0]);
}

static double rownorm_f64(const plan_t *p, const double *a, const double *b, int shift) {
    double acc[MAXLANES][VEC];
    memset(acc, 0, sizeof(acc));
    const uint8_t *ln = p->lane[shift], *ac = p->acc[shift];
    for (int e = 0; e < p->d; ++e) {
        double v = a[e] - b[e];
        acc[ln[e]][ac[e]] = fma(v, v, acc[ln[e]][ac[e]]);
    }
    double lanev[MAXLANES];
    for (int l = 0; l < p->lanes; ++l) lanev[l] = ((acc[l][0] + acc[l][1]) + acc[l][2]) + acc[l][3];
    for (int off = 1; off < p->lanes; off <<= 1)
        for (int l = 0; l + off < p->lanes; l += 2 * off) lanev[l] = lanev[l] + lanev[l + off];
    return sqrt(lanev[0]);
}

/* torch's double -> Half conversion goes through float (c10::Half has only a float constructor) */
static half_t f64_to_f16_like_torch(double v) { return (half_t)(float)v; }

/*
 * FP16 mode (reference default, main.py:151).  z: float64 [N,d] projected bank (what sklearn returns).
 * out_idx: int64 [n_select].  out_zh (optional): the half bank, out_min0 (optional): initial half min-distances,
 * out_last_min (optional): final min-distance vector as half bits.  force_idx (optional, teacher forcing): if not
 * NULL, pick k uses force_idx[k] as the selected row instead of the argmax, while out_idx still records the argmax.
 */
int oracle_coreset_fp16(const double *z, int64_t N, int d, int64_t n_select, int64_t *out_idx, uint16_t *out_zh,
                        uint16_t *out_min0, uint16_t *out_last_min, const int64_t *force_idx) {
    if (N <= 0 || d <= 0 || n_select <= 0 || n_select > N) return -1;
    plan_t p;
    plan_init(&p, d);
    half_t *zh = (half_t *)malloc(sizeof(half_t) * (size_t)N * d);
    half_t *mind = (half_t *)malloc(sizeof(half_t) * (size_t)N);
    half_t *last = (half_t *)malloc(sizeof(half_t) * (size_t)d);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        for (int e = 0; e < d; ++e) zh[i * d + e] = f64_to_f16_like_torch(z[i * d + e]);
        /* features.py:378 -- float64 on the host, then .half() (:391) */
        mind[i] = f64_to_f16_like_torch(rownorm_f64(&p, z + i * d, z, (int)((i * d) % VEC)));
    }
    if (out_zh) memcpy(out_zh, zh, sizeof(half_t) * (size_t)N * d);
    if (out_min0) memcpy(out_min0, mind, sizeof(half_t) * (size_t)N);
    int64_t sel = 0;
    out_idx[0] = 0;
    for (int64_t k = 1; k < n_select; ++k) {
        memcpy(last, zh + sel * d, sizeof(half_t) * (size_t)d);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            half_t dist = rownorm_f16(&p, zh + i * d, last, (int)((i * d) % VEC));
            /* torch.minimum propagates NaN; distances here are never NaN for finite input */
            if ((float)dist < (float)mind[i]) mind[i] = dist;
        }
        int64_t best = 0;
        float bv = (float)mind[0];
        for (int64_t i = 1; i < N; ++i) /* argmax, ties -> lowest index (features.py:415) */
            if ((float)mind[i] > bv) {
                bv = (float)mind[i];
                best = i;
            }
        out_idx[k] = best;
        sel = force_idx ? force_idx[k] : best;
        mind[sel] = (half_t)0.0f; /* features.py:419 */
    }
    if (out_last_min) memcpy(out_last_min, mind, sizeof(half_t) * (size_t)N);
    free(zh);
    free(mind);
    free(last);
    plan_free(&p);
    return 0;
}

/* 'TF32' mode of the reference (features.py:392-393): only a matmul flag is set, the data stays float64. */
int oracle_coreset_fp64(const double *z, int64_t N, int d, int64_t n_select, int64_t *out_idx, double *out_min0,
                        double *out_last_min, const int64_t *force_idx) {
    if (N <= 0 || d <= 0 || n_select <= 0 || n_select > N) return -1;
    plan_t p;
    plan_init(&p, d);
    double *mind = (double *)malloc(sizeof(double) * (size_t)N);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) mind[i] = rownorm_f64(&p, z + i * d, z, (int)((i * d) % VEC));
    if (out_min0) memcpy(out_min0, mind, sizeof(double) * (size_t)N);
    int64_t sel = 0;
    out_idx[0] = 0;
    for (int64_t k = 1; k < n_select; ++k) {
        const double *last = z + sel * d;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            double dist = rownorm_f64(&p, z + i * d, last, (int)((i * d) % VEC));
            if (dist < mind[i]) mind[i] = dist;
        }
        int64_t best = 0;
        double bv = mind[0];
        for (int64_t i = 1; i < N; ++i)
            if (mind[i] > bv) {
                bv = mind[i];
                best = i;
            }
        out_idx[k] = best;
        sel = force_idx ? force_idx[k] : best;
        mind[sel] = 0.0;
    }
    if (out_last_min) memcpy(out_last_min, mind, sizeof(double) * (size_t)N);
    free(mind);
    plan_free(&p);
    return 0;
}

/* One canonical-order distance pass (used to pin the order against torch on the GPU box): half in, half out. */
int oracle_rownorms_fp16(const uint16_t *zh_bits, const uint16_t *last_bits, int64_t N, int d, uint16_t *out_bits) {
    plan_t p;
    plan_init(&p, d);
    const half_t *zh = (const half_t *)zh_bits;
    const half_t *last = (const half_t *)last_bits;
    half_t *out = (half_t *)out_bits;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) out[i] = rownorm_f16(&p, zh + i * d, last, (int)((i * d) % VEC));
    plan_free(&p);
    return 0;
}

/* probe variant: multiply and add rounded separately (no fma) -- used once to pin which one torch's CUDA build does */
int oracle_rownorms_fp64_nofma(const double *z, const double *last, int64_t N, int d, double *out) {
    plan_t p;
    plan_init(&p, d);
    for (int64_t i = 0; i < N; ++i) {
        double acc[MAXLANES][VEC];
        memset(acc, 0, sizeof(acc));
        int shift = (int)((i * d) % VEC);
        const uint8_t *ln = p.lane[shift], *ac = p.acc[shift];
        for (int e = 0; e < d; ++e) {
            volatile double v = z[i * d + e] - last[e];
            volatile double sq = v * v;
            acc[ln[e]][ac[e]] = acc[ln[e]][ac[e]] + sq;
        }
        double lanev[MAXLANES];
        for (int l = 0; l < p.lanes; ++l) lanev[l] = ((acc[l][0] + acc[l][1]) + acc[l][2]) + acc[l][3];
        for (int off = 1; off < p.lanes; off <<= 1)
            for (int l = 0; l + off < p.lanes; l += 2 * off) lanev[l] = lanev[l] + lanev[l + off];
        out[i] = sqrt(lanev[0]);
    }
    plan_free(&p);
    return 0;
}

int oracle_rownorms_fp64(const double *z, const double *last, int64_t N, int d, double *out) {
    plan_t p;
    plan_init(&p, d);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) out[i] = rownorm_f64(&p, z + i * d, last, (int)((i * d) % VEC));
    plan_free(&p);
    return 0;
}

/*
 * sklearn SparseRandomProjection.transform restated (features.py:365-366): X is up-cast to float64 and multiplied by
 * components_.T through scipy's csr_matvecs, i.e. for every output (i, j): y = 0; for k in row j of the CSR matrix in
 * STORED order: y += data[k] * x[i, indices[k]] (separate multiply and add, float64).  scipy 1.18 / sklearn 1.9 in
 * this image (requirements.txt pins sklearn 1.4.0; same algorithm).
 */
int oracle_sparse_project(const float *x, int64_t N, int D, const int32_t *indptr, const int32_t *indices,
                          const double *data, int d_proj, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const float *xr = x + i * D;
        for (int j = 0; j < d_proj; ++j) {
            volatile double y = 0.0; /* volatile: forbid fma contraction / reassociation */
            for (int32_t k = indptr[j]; k < indptr[j + 1]; ++k) {
                volatile double prod = data[k] * (double)xr[indices[k]];
                y = y + prod;
            }
            out[i * d_proj + j] = y;
        }
    }
    return 0;
}
