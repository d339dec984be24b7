// Everything of compute_single_s_s_map (reference features.py:225-297) after the distance GEMM:
//   refine      exact float32 re-check of the GEMM epilogue's per-CTA top-2 candidates -> min_val / min_idx (:227),
//               and the packed argmax key of min_val -> s_star / s_idx (:228-231)
//   select      m_test = patch[s_idx], m_star = bank[min_idx[s_idx]] (:235-251)
//   reweight    w_dist = ||m_star - bank_r|| for every bank row, 3 smallest (:239-254; HBM bound, R*D*4 bytes), with
//               the m_star selection as prologue and m_star_knn, w, s (:275-290) as last-block epilogue
//   upsample_blur  bilinear 28^2/56^2 -> 224^2 (:293-294) + KNNGaussianBlur (utils/utils.py:71-83): /max, 8-bit
//               truncation, Pillow's 3+3 pass integer box blur, /255, *max -- one CTA, whole image in shared memory
#include <math.h>

#include "common.cuh"

namespace cmdb {

__device__ __forceinline__ unsigned long long pack_min_key(float v, unsigned int idx) {
    return ((unsigned long long)__float_as_uint(v) << 32) | idx;  // v >= 0: unsigned order == float order
}

// exact ||a - b||^2 in float32, one warp: lanes stride over float4s, xor-shuffle tree
__device__ __forceinline__ float warp_sqdist(const float *__restrict__ a, const float *__restrict__ b, int dim4, int lane) {
    float acc = 0.f;
    for (int c = lane; c < dim4; c += 32) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(a) + c), y = __ldg(reinterpret_cast<const float4 *>(b) + c);
        float d;
        d = x.x - y.x, acc = fmaf(d, d, acc);
        d = x.y - y.y, acc = fmaf(d, d, acc);
        d = x.z - y.z, acc = fmaf(d, d, acc);
        d = x.w - y.w, acc = fmaf(d, d, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// ---------------------------------------------------------------------------------------------------------------
// refine: one warp per query.  cand[q][c] = (val1, idx1, val2, idx2) from n_cand producers (approximate d^2 up to a
// per-query constant; idx < 0 = empty).  Takes the 4 best approximate candidates, recomputes their distance exactly
// and keeps the smallest (ties -> lowest row).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRefineTop = 4;

__global__ void __launch_bounds__(256) refine_kernel(const float4 *__restrict__ cand, int n_cand, int cand_stride,
                                                     const float *__restrict__ q, const float *__restrict__ bank, int dim,
                                                     int P, long long row_offset, float *__restrict__ min_val,
                                                     long long *__restrict__ min_idx, unsigned long long *s_key) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= P) return;
    // each lane keeps its own sorted top-4 (approx value, row) over the candidates it scans
    float tv[kRefineTop];
    int ti[kRefineTop];
#pragma unroll
    for (int k = 0; k < kRefineTop; ++k) tv[k] = INFINITY, ti[k] = -1;
    auto insert = [&](float v, int i) {
        if (i < 0) return;
#pragma unroll
        for (int k = 0; k < kRefineTop; ++k) {
            if (ti[k] == i) return;  // duplicates cannot happen across producers, cheap guard anyway
            if (v < tv[k] || (v == tv[k] && i < ti[k]) || ti[k] < 0) {
                float fv = tv[k];
                int fi = ti[k];
                tv[k] = v, ti[k] = i;
                v = fv, i = fi;
                if (i < 0) return;
            }
        }
    };
    for (int c = lane; c < n_cand; c += 32) {
        const float4 t = cand[(size_t)qi * cand_stride + c];
        insert(t.x, __float_as_int(t.y));
        insert(t.z, __float_as_int(t.w));
    }
    // warp-wide top-4: pop the global minimum 4 times
    int sel_idx[kRefineTop];
#pragma unroll
    for (int r = 0; r < kRefineTop; ++r) {
        float v = tv[0];
        int i = ti[0];
        if (i < 0) v = INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (oi >= 0 && (i < 0 || ov < v || (ov == v && oi < i))) v = ov, i = oi;
        }
        sel_idx[r] = i;
        if (i >= 0 && ti[0] == i) {  // owner pops
#pragma unroll
            for (int k = 0; k + 1 < kRefineTop; ++k) tv[k] = tv[k + 1], ti[k] = ti[k + 1];
            tv[kRefineTop - 1] = INFINITY, ti[kRefineTop - 1] = -1;
        }
    }
    float best = INFINITY;
    int best_i = -1;
#pragma unroll
    for (int r = 0; r < kRefineTop; ++r) {
        const int i = sel_idx[r];
        if (i < 0) continue;
        const float d2 = warp_sqdist(q + (size_t)qi * dim, bank + (size_t)i * dim, dim >> 2, lane);
        if (best_i < 0 || d2 < best || (d2 == best && i < best_i)) best = d2, best_i = i;
    }
    if (lane == 0) {
        const float dv = sqrtf(best);
        min_val[qi] = dv;
        min_idx[qi] = best_i < 0 ? -1 : (long long)best_i + row_offset;
        // argmax over queries, ties -> lowest query: max of (value bits, ~query)
        atomicMax(s_key, ((unsigned long long)__float_as_uint(dv) << 32) | (0xffffffffu - (unsigned int)qi));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// diagnostics scorer: exact direct-form float32 distances on CUDA cores, same candidate format as the GEMM epilogue.
// grid (query tiles of 64, bank slices); block 256 = 16x16 threads, 4x4 outputs each, K chunks of 32 through smem.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simt_min_kernel(const float *__restrict__ q, int P, const float *__restrict__ bank,
                                                       long long R, int dim, float4 *__restrict__ cand, int cand_stride) {
    __shared__ float qs[64][33], bs[64][33];
    __shared__ float rv[64][16][2];
    __shared__ int ri[64][16][2];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int q0 = blockIdx.x * 64;
    const long long rows_per_slice = ((R + gridDim.y - 1) / gridDim.y + 63) / 64 * 64;
    const long long r_begin = (long long)blockIdx.y * rows_per_slice, r_end = min(R, r_begin + rows_per_slice);
    float b1[4], b2[4];
    int i1[4], i2[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) b1[a] = b2[a] = INFINITY, i1[a] = i2[a] = -1;
    for (long long r0 = r_begin; r0 < r_end; r0 += 64) {
        float acc[4][4] = {};
        for (int k0 = 0; k0 < dim; k0 += 32) {
            __syncthreads();
            for (int i = threadIdx.x; i < 64 * 32; i += 256) {
                const int r = i >> 5, c = i & 31;
                qs[r][c] = (q0 + r < P) ? q[(size_t)(q0 + r) * dim + k0 + c] : 0.f;
                bs[r][c] = (r0 + r < r_end) ? bank[(size_t)(r0 + r) * dim + k0 + c] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                float qa[4], ba[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) qa[a] = qs[ty * 4 + a][k], ba[a] = bs[tx * 4 + a][k];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float d = qa[a] - ba[c];
                        acc[a][c] = fmaf(d, d, acc[a][c]);
                    }
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long long r = r0 + tx * 4 + c;
                if (r >= r_end) continue;
                const float v = acc[a][c];
                const int ri_ = (int)r;
                if (v < b1[a] || (v == b1[a] && ri_ < i1[a]) || i1[a] < 0) {
                    b2[a] = b1[a], i2[a] = i1[a], b1[a] = v, i1[a] = ri_;
                } else if (v < b2[a] || (v == b2[a] && ri_ < i2[a]) || i2[a] < 0) {
                    b2[a] = v, i2[a] = ri_;
                }
            }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        rv[ty * 4 + a][tx][0] = b1[a], rv[ty * 4 + a][tx][1] = b2[a];
        ri[ty * 4 + a][tx][0] = i1[a], ri[ty * 4 + a][tx][1] = i2[a];
    }
    __syncthreads();
    if (threadIdx.x < 64 && q0 + threadIdx.x < P) {
        float c1 = INFINITY, c2 = INFINITY;
        int j1 = -1, j2 = -1;
        for (int t = 0; t < 16; ++t)
            for (int u = 0; u < 2; ++u) {
                const float v = rv[threadIdx.x][t][u];
                const int i = ri[threadIdx.x][t][u];
                if (i < 0) continue;
                if (j1 < 0 || v < c1 || (v == c1 && i < j1)) c2 = c1, j2 = j1, c1 = v, j1 = i;
                else if (j2 < 0 || v < c2 || (v == c2 && i < j2)) c2 = v, j2 = i;
            }
        cand[(size_t)(q0 + threadIdx.x) * cand_stride + blockIdx.y] = make_float4(c1, __int_as_float(j1), c2, __int_as_float(j2));
    }
}

int score_simt_candidates(cmdb_bank *b, int P, int *n_cand_out) {
    const int q_tiles = (P + 63) / 64;
    int slices = std::max(1, std::min(b->num_sms, (2 * b->num_sms + q_tiles - 1) / q_tiles));
    slices = (int)std::min<long long>(slices, (b->fin_rows + 63) / 64);
    simt_min_kernel<<<dim3(q_tiles, slices), 256, 0, b->stream>>>(b->ss.q_f32, P, b->data, b->fin_rows, b->dim, b->ss.cand,
                                                                 b->num_sms);
    CMDB_CUDA(cudaGetLastError());
    *n_cand_out = slices;
    return CMDB_OK;
}

int score_refine(cmdb_bank *b, int P, int n_cand) {
    CMDB_CUDA(cudaMemsetAsync(b->ss.s_key, 0, sizeof(unsigned long long), b->stream));
    refine_kernel<<<(P + 7) / 8, 256, 0, b->stream>>>(b->ss.cand, n_cand, b->num_sms, b->ss.q_f32, b->data, b->dim, P,
                                                      b->row_offset, b->ss.min_val, b->ss.min_idx, b->ss.s_key);
    CMDB_CUDA(cudaGetLastError());
    return CMDB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// select: decode the argmax key, stage m_test and (single-GPU) m_star rows
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_kernel(const unsigned long long *s_key, const long long *min_idx,
                                                     const float *__restrict__ q, const float *__restrict__ bank, int dim,
                                                     long long row_offset, long long rows, float *m_test, float *m_star,
                                                     TailResult *res) {
    const unsigned long long key = *s_key;
    const int s_idx = (int)(0xffffffffu - (unsigned int)(key & 0xffffffffu));
    const long long g = min_idx[s_idx];
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        m_test[c] = q[(size_t)s_idx * dim + c];
        const long long l = g - row_offset;
        if (m_star && l >= 0 && l < rows) m_star[c] = bank[(size_t)l * dim + c];
    }
    if (threadIdx.x == 0) {
        res->s_idx = s_idx;
        res->s_star = __uint_as_float((unsigned int)(key >> 32));
        res->m_star_row = g;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// reweight: exact ||m_star - bank_r||^2 for every local bank row, 3 smallest as packed keys
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void top3_insert(unsigned long long (&t)[3], unsigned long long k) {
    if (k < t[2]) {
        if (k < t[1]) {
            t[2] = t[1];
            if (k < t[0]) t[1] = t[0], t[0] = k;
            else t[1] = k;
        } else {
            t[2] = k;
        }
    }
}

// block-wide 3 smallest of the per-thread sorted triples t[]: three rounds of "everyone offers its head, winner pops"
__device__ __forceinline__ void block_top3(unsigned long long (&t)[3], unsigned long long *sh /* [8] */,
                                           unsigned long long (&out)[3]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        unsigned long long v = t[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ov = __shfl_xor_sync(0xffffffffu, v, o);
            v = ov < v ? ov : v;
        }
        __syncthreads();
        if (lane == 0) sh[warp] = v;
        __syncthreads();
        v = sh[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) v = sh[w] < v ? sh[w] : v;
        out[r] = v;
        if (v != ~0ULL && t[0] == v) t[0] = t[1], t[1] = t[2], t[2] = ~0ULL;  // keys are unique (row in the low bits)
    }
}

struct ReweightParams {
    const float *bank;               // [rows, dim] float32 (local shard)
    long long rows, row_offset;
    int dim;
    const float *q;                  // [P, dim] normalised patches
    const unsigned long long *s_key; // packed argmax of min_val
    const long long *min_idx;        // [P] global rows
    const float *m_star_explicit;    // sharded mode: replicated m_star row; NULL = take it from the local bank
    unsigned long long *block_keys;  // [gridDim.x * 3]
    unsigned long long *top3;        // [3] merged result
    unsigned int *done_counter;      // last-block-done counter (self resetting)
    TailResult *res;
    int fuse_final;                  // 1: the last block also computes m_star_knn, w, s (single-GPU path)
};

// w_dist pass (features.py:239-254): exact ||m_star - bank_r||^2 for every local bank row, 3 smallest.
// Prologue = "select" (decode s*, s_idx, locate m_star); epilogue = last-block-done merge (+ final re-weighting), so the
// whole re-weighting stage is one launch.  HBM bound: rows*dim*4 bytes.
__global__ void __launch_bounds__(256) reweight_kernel(ReweightParams p) {
    extern __shared__ __align__(16) float ms[];
    __shared__ unsigned long long wk[8];
    __shared__ bool is_last;
    __shared__ float knn[2];
    const unsigned long long skey = *p.s_key;
    const int s_idx = (int)(0xffffffffu - (unsigned int)(skey & 0xffffffffu));
    const long long g_star = p.min_idx[s_idx];
    const float *m_star = p.m_star_explicit ? p.m_star_explicit : p.bank + (size_t)(g_star - p.row_offset) * p.dim;
    for (int c = threadIdx.x; c < p.dim; c += blockDim.x) ms[c] = m_star[c];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dim = p.dim, dim4 = dim >> 2;
    unsigned long long t[3] = {~0ULL, ~0ULL, ~0ULL};
    const long long warps = (long long)gridDim.x * 8;
    // two rows in flight per warp
    for (long long r = (long long)blockIdx.x * 8 + warp; r < p.rows; r += 2 * warps) {
        const long long r2 = r + warps;
        const bool has2 = r2 < p.rows;
        const float4 *a = reinterpret_cast<const float4 *>(p.bank + (size_t)r * dim);
        const float4 *a2 = reinterpret_cast<const float4 *>(p.bank + (size_t)(has2 ? r2 : r) * dim);
        float acc = 0.f, acc2 = 0.f;
        for (int c = lane; c < dim4; c += 32) {
            const float4 x = __ldg(a + c), x2 = __ldg(a2 + c);
            const float4 y = reinterpret_cast<const float4 *>(ms)[c];
            float d;
            d = x.x - y.x, acc = fmaf(d, d, acc);
            d = x.y - y.y, acc = fmaf(d, d, acc);
            d = x.z - y.z, acc = fmaf(d, d, acc);
            d = x.w - y.w, acc = fmaf(d, d, acc);
            d = x2.x - y.x, acc2 = fmaf(d, d, acc2);
            d = x2.y - y.y, acc2 = fmaf(d, d, acc2);
            d = x2.z - y.z, acc2 = fmaf(d, d, acc2);
            d = x2.w - y.w, acc2 = fmaf(d, d, acc2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
        }
        if (lane == 0) {  // one copy per warp keeps the keys unique for block_top3
            top3_insert(t, pack_min_key(acc, (unsigned int)(r + p.row_offset)));
            if (has2) top3_insert(t, pack_min_key(acc2, (unsigned int)(r2 + p.row_offset)));
        }
    }
    unsigned long long f[3];
    block_top3(t, wk, f);
    if (threadIdx.x == 0) {
        p.block_keys[blockIdx.x * 3 + 0] = f[0];
        p.block_keys[blockIdx.x * 3 + 1] = f[1];
        p.block_keys[blockIdx.x * 3 + 2] = f[2];
        __threadfence();
        is_last = atomicAdd(p.done_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    // ---- last block: merge every block's keys, then (single-GPU path) finish the re-weighting ----
    __threadfence();
    t[0] = t[1] = t[2] = ~0ULL;
    for (int i = threadIdx.x; i < (int)gridDim.x * 3; i += blockDim.x) top3_insert(t, __ldcg(p.block_keys + i));
    block_top3(t, wk, f);
    if (threadIdx.x == 0) {
        p.top3[0] = f[0], p.top3[1] = f[1], p.top3[2] = f[2];
        *p.done_counter = 0;
        p.res->s_idx = s_idx;
        p.res->s_star = __uint_as_float((unsigned int)(skey >> 32));
        p.res->m_star_row = g_star;
        for (int k = 0; k < 3; ++k) p.res->nn_idx[k] = f[k] == ~0ULL ? -1 : (long long)(f[k] & 0xffffffffULL);
    }
    if (!p.fuse_final) return;
    // features.py:275-283: m_star_knn = ||m_test - bank[nn_idx[1:]]||, m_test = patch[s_idx]
    if (warp < 2) {
        const unsigned long long key = f[1 + warp];
        float d2 = 0.f;
        if (key != ~0ULL)
            d2 = warp_sqdist(p.q + (size_t)s_idx * dim, p.bank + (size_t)((long long)(key & 0xffffffffULL) - p.row_offset) * dim,
                             dim4, lane);
        if (lane == 0) knn[warp] = key != ~0ULL ? sqrtf(d2) : NAN;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float Dn = sqrtf((float)dim);  // torch.sqrt(torch.tensor(patch.shape[1]))  (features.py:285)
        const float s_star = __uint_as_float((unsigned int)(skey >> 32));
        const float den = expf(knn[0] / Dn) + expf(knn[1] / Dn);
        const float w = 1.f - expf(s_star / Dn) / den;  // features.py:287
        p.res->w = w;
        p.res->s = w * s_star;                          // features.py:290
        p.res->knn0 = knn[0], p.res->knn1 = knn[1];
    }
}

// merge gathered keys -> 3 smallest (one block; sharded mode, after the all-gather)
__global__ void __launch_bounds__(256) merge_top3_kernel(const unsigned long long *__restrict__ keys, int n_keys,
                                                         unsigned long long *__restrict__ out3) {
    __shared__ unsigned long long wk[8];
    unsigned long long t[3] = {~0ULL, ~0ULL, ~0ULL}, f[3];
    for (int i = threadIdx.x; i < n_keys; i += blockDim.x) top3_insert(t, keys[i]);
    block_top3(t, wk, f);
    if (threadIdx.x == 0) out3[0] = f[0], out3[1] = f[1], out3[2] = f[2];
}

// final: m_star_knn = ||m_test - bank[nn[1:]]|| (features.py:275-283), w and s (:285-290).  nn_rows: optional
// [3][dim] rows supplied by the caller (sharded mode); otherwise rows are read from the local bank.
__global__ void __launch_bounds__(64) final_kernel(const unsigned long long *__restrict__ keys3, const float *__restrict__ m_test,
                                                   const float *__restrict__ bank, const float *__restrict__ nn_rows, int dim,
                                                   long long row_offset, TailResult *res) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 2 warps: one per neighbour
    __shared__ float knn[2];
    const unsigned long long key = keys3[1 + warp];
    const long long g = (long long)(key & 0xffffffffULL);
    const bool valid = key != ~0ULL;
    float d2 = 0.f;
    if (valid) {
        const float *row = nn_rows ? nn_rows + (size_t)(1 + warp) * dim : bank + (size_t)(g - row_offset) * dim;
        d2 = warp_sqdist(m_test, row, dim >> 2, lane);
    }
    if (lane == 0) knn[warp] = valid ? sqrtf(d2) : NAN;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float Dn = sqrtf((float)dim);  // torch.sqrt(torch.tensor(patch.shape[1]))
        const float s_star = res->s_star;
        const float den = expf(knn[0] / Dn) + expf(knn[1] / Dn);
        const float w = 1.f - expf(s_star / Dn) / den;
        res->w = w;
        res->s = w * s_star;
        res->knn0 = knn[0], res->knn1 = knn[1];
        for (int k = 0; k < 3; ++k) res->nn_idx[k] = keys3[k] == ~0ULL ? -1 : (long long)(keys3[k] & 0xffffffffULL);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// upsample + blur in two small multi-CTA kernels (the image is split into 16 row bands, then 16 column bands):
//   K1  every CTA recomputes the global max of the upsampled map (needed before the 8-bit quantisation; 50k pixels, cheaper
//       than a grid barrier), upsamples + quantises its row band and runs the 3 horizontal box passes in shared memory;
//   K2  every CTA loads its column band of K1's result and runs the 3 vertical passes (Pillow transposes instead), then
//       applies ToTensor (/255) and * max.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBlurThreads = 512;
constexpr int kBlurBands = 16;

__device__ __forceinline__ void bilinear_coeff(int dst, float scale, int n_in, int &i0, int &i1, float &w0, float &w1) {
    // ATen area_pixel_compute_source_index (align_corners=False) + HelperInterpLinear weights
    float real = fmaxf(__fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f), 0.f);
    i0 = (int)real;
    i1 = min(i0 + 1, n_in - 1);
    w1 = __fsub_rn(real, (float)i0);
    w0 = __fsub_rn(1.f, w1);
}

__device__ __forceinline__ float bilinear_at(const float *__restrict__ in_s, int fh, int fw_, float sh, float sw, int y, int x) {
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    bilinear_coeff(y, sh, fh, y0, y1, wy0, wy1);
    bilinear_coeff(x, sw, fw_, x0, x1, wx0, wx1);
    // ATen Interpolate<2,...,interp_size 2>: t0*w0 + t1*w1 contracted as fma(t0, w0, t1*w1), W inside H
    const float t0 = __fmaf_rn(in_s[y0 * fw_ + x0], wx0, __fmul_rn(in_s[y0 * fw_ + x1], wx1));
    const float t1 = __fmaf_rn(in_s[y1 * fw_ + x0], wx0, __fmul_rn(in_s[y1 * fw_ + x1], wx1));
    return __fmaf_rn(t0, wy0, __fmul_rn(t1, wy1));
}

// one box pass along the `n`-long axis of an [n_lines][n] (x_stride == 1) or [n][n_lines] (x_stride == n_lines) tile;
// Pillow ImagingLineBoxBlur8 as a clamped 9-tap integer FIR with 32-bit fixed-point weights
__device__ __forceinline__ void box_pass(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, int n_lines,
                                         int n, int line_stride, int x_stride, int radius, unsigned int ww, unsigned int fw) {
    for (int i = threadIdx.x; i < n_lines * n; i += kBlurThreads) {
        // consecutive threads walk the contiguous tile dimension in both orientations (bank-conflict free)
        int line, x;
        if (x_stride == 1) line = i / n, x = i - line * n;
        else x = i / n_lines, line = i - x * n_lines;
        const unsigned char *ln = src + line * line_stride;
        unsigned int acc = 0;
        for (int k = -radius; k <= radius; ++k) acc += ln[min(max(x + k, 0), n - 1) * x_stride];
        const unsigned int far = ln[max(x - radius - 1, 0) * x_stride] + ln[min(x + radius + 1, n - 1) * x_stride];
        const unsigned int bulk = acc * ww + far * fw;
        dst[line * line_stride + x * x_stride] = (unsigned char)((bulk + (1u << 23)) >> 24);
    }
}

__global__ void __launch_bounds__(kBlurThreads) upsample_hblur_kernel(const float *__restrict__ map_in, int fh, int fw_,
                                                                      int out_hw, int band, float *__restrict__ pre,
                                                                      unsigned char *__restrict__ u8_out,
                                                                      unsigned char *__restrict__ tmp, float *__restrict__ mx_out,
                                                                      int radius, unsigned int ww, unsigned int fwt) {
    extern __shared__ __align__(16) unsigned char sm[];
    float *in_s = reinterpret_cast<float *>(sm);                    // [fh*fw]
    unsigned char *A = sm + sizeof(float) * ((fh * fw_ + 3) & ~3);  // [band][out_hw]
    unsigned char *B = A + ((band * out_hw + 15) & ~15);
    __shared__ float red[kBlurThreads / 32];
    for (int i = threadIdx.x; i < fh * fw_; i += kBlurThreads) in_s[i] = map_in[i];
    __syncthreads();
    const float sh = (float)fh / (float)out_hw, sw = (float)fw_ / (float)out_hw;
    const int npix = out_hw * out_hw;
    float mx = -INFINITY;  // map_max = img.max() over the WHOLE upsampled image (utils/utils.py:81)
    for (int i = threadIdx.x; i < npix; i += kBlurThreads) mx = fmaxf(mx, bilinear_at(in_s, fh, fw_, sh, sw, i / out_hw, i % out_hw));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < kBlurThreads / 32; ++w) mx = fmaxf(mx, red[w]);
    if (blockIdx.x == 0 && threadIdx.x == 0) *mx_out = mx;
    const int y0 = blockIdx.x * band, rows = min(band, out_hw - y0);
    if (rows <= 0) return;
    // KNNGaussianBlur: img / max -> ToPILImage: mul(255).byte() (truncation)
    for (int i = threadIdx.x; i < rows * out_hw; i += kBlurThreads) {
        const int y = y0 + i / out_hw, x = i % out_hw;
        const float v = bilinear_at(in_s, fh, fw_, sh, sw, y, x);
        if (pre) pre[y * out_hw + x] = v;
        const unsigned char u = (unsigned char)(int)__fmul_rn(__fdiv_rn(v, mx), 255.f);
        A[i] = u;
        if (u8_out) u8_out[y * out_hw + x] = u;
    }
    __syncthreads();
    box_pass(A, B, rows, out_hw, out_hw, 1, radius, ww, fwt);
    __syncthreads();
    box_pass(B, A, rows, out_hw, out_hw, 1, radius, ww, fwt);
    __syncthreads();
    box_pass(A, B, rows, out_hw, out_hw, 1, radius, ww, fwt);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * out_hw; i += kBlurThreads) tmp[y0 * out_hw + i] = B[i];
}

__global__ void __launch_bounds__(kBlurThreads) vblur_kernel(const unsigned char *__restrict__ tmp, int out_hw, int band,
                                                             const float *__restrict__ mx_in, float *__restrict__ out,
                                                             int radius, unsigned int ww, unsigned int fwt) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int x0 = blockIdx.x * band, cols = min(band, out_hw - x0);
    if (cols <= 0) return;
    unsigned char *A = sm;  // [out_hw][cols]: column band, lines = columns (stride 1), pass axis = rows (stride cols)
    unsigned char *B = A + ((out_hw * band + 15) & ~15);
    for (int i = threadIdx.x; i < out_hw * cols; i += kBlurThreads) {
        const int y = i / cols, c = i - y * cols;
        A[i] = tmp[y * out_hw + x0 + c];
    }
    __syncthreads();
    box_pass(A, B, cols, out_hw, 1, cols, radius, ww, fwt);
    __syncthreads();
    box_pass(B, A, cols, out_hw, 1, cols, radius, ww, fwt);
    __syncthreads();
    box_pass(A, B, cols, out_hw, 1, cols, radius, ww, fwt);
    __syncthreads();
    const float mx = *mx_in;
    // ToTensor (/255) then * map_max
    for (int i = threadIdx.x; i < out_hw * cols; i += kBlurThreads) {
        const int y = i / cols, c = i - y * cols;
        out[y * out_hw + x0 + c] = __fmul_rn(__fdiv_rn((float)B[i], 255.f), mx);
    }
}

int score_select(cmdb_bank *b, bool local_m_star) {
    select_kernel<<<1, 256, 0, b->stream>>>(b->ss.s_key, b->ss.min_idx, b->ss.q_f32, b->data, b->dim, b->row_offset,
                                            b->fin_rows, b->ss.m_test, local_m_star ? b->ss.m_star : nullptr,
                                            reinterpret_cast<TailResult *>(b->ss.tail));
    CMDB_CUDA(cudaGetLastError());
    return CMDB_OK;
}

// fused = single-GPU path (select prologue + merge + final in one launch); otherwise m_star comes from ss.m_star and only
// the merged top-3 keys are produced
int score_reweight(cmdb_bank *b, bool fused) {
    ReweightParams p{};
    p.bank = b->data, p.rows = b->fin_rows, p.row_offset = b->row_offset, p.dim = b->dim;
    p.q = b->ss.q_f32, p.s_key = b->ss.s_key, p.min_idx = b->ss.min_idx;
    p.m_star_explicit = fused ? nullptr : b->ss.m_star;
    p.block_keys = b->ss.topk_keys, p.top3 = b->ss.top3, p.done_counter = b->ss.done_counter;
    p.res = reinterpret_cast<TailResult *>(b->ss.tail);
    p.fuse_final = fused ? 1 : 0;
    reweight_kernel<<<b->ss.n_topk_blocks, 256, sizeof(float) * b->dim, b->stream>>>(p);
    CMDB_CUDA(cudaGetLastError());
    return CMDB_OK;
}

int score_merge_top3(cmdb_bank *b, int n_keys) {
    merge_top3_kernel<<<1, 256, 0, b->stream>>>(b->ss.topk_keys, n_keys, b->ss.top3);
    CMDB_CUDA(cudaGetLastError());
    return CMDB_OK;
}

int score_final(cmdb_bank *b, bool use_nn_rows) {
    final_kernel<<<1, 64, 0, b->stream>>>(b->ss.top3, b->ss.m_test, b->data, use_nn_rows ? b->ss.nn_rows : nullptr, b->dim,
                                          b->row_offset, reinterpret_cast<TailResult *>(b->ss.tail));
    CMDB_CUDA(cudaGetLastError());
    return CMDB_OK;
}

int upsample_blur_launch(cudaStream_t stream, const float *map_dev, int fh, int fw, int out_hw, float *pre_dev,
                         float *out_dev, unsigned char *u8_dev, unsigned char *tmp_dev, float *mx_dev) {
    CMDB_REQUIRE(fh > 0 && fw > 0 && out_hw >= 8 && out_hw <= 256, CMDB_ERR_INVALID,
                 "upsample_blur: need out_hw in [8,256] (got %d) and positive map dims", out_hw);
    // Pillow _gaussian_blur_radius(radius=4, passes=3) evaluated like BoxBlur.c (float sigma2 = 16/3, the rest double)
    const float sigma2 = 4.f * 4.f / 3.f;
    const double L = sqrt(12.0 * (double)sigma2 + 1.0);
    const double l = floor((L - 1.0) / 2.0);
    double a = (2 * l + 1) * (l * (l + 1) - 3 * (double)sigma2);
    a /= 6 * ((double)sigma2 - (l + 1) * (l + 1));
    const float fr = (float)(l + a);
    const int radius = (int)fr;
    const unsigned int ww = (unsigned int)((float)(1 << 24) / (fr * 2 + 1));
    const unsigned int fwt = ((1u << 24) - (unsigned int)(radius * 2 + 1) * ww) / 2;
    CMDB_REQUIRE(out_hw > radius + 1, CMDB_ERR_INVALID, "upsample_blur: image smaller than the blur radius");
    const int band = (out_hw + kBlurBands - 1) / kBlurBands;
    const size_t tile = (size_t)((band * out_hw + 15) & ~15);
    const size_t smem1 = sizeof(float) * ((fh * fw + 3) & ~3) + 2 * tile;
    CMDB_REQUIRE(smem1 <= 200 * 1024, CMDB_ERR_UNSUPPORTED, "upsample_blur: map too large for shared memory");
    CMDB_CUDA(cudaFuncSetAttribute(upsample_hblur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    upsample_hblur_kernel<<<kBlurBands, kBlurThreads, smem1, stream>>>(map_dev, fh, fw, out_hw, band, pre_dev, u8_dev, tmp_dev,
                                                                      mx_dev, radius, ww, fwt);
    vblur_kernel<<<kBlurBands, kBlurThreads, 2 * tile, stream>>>(tmp_dev, out_hw, band, mx_dev, out_dev, radius, ww, fwt);
    CMDB_CUDA(cudaGetLastError());
    return CMDB_OK;
}

}  // namespace cmdb
