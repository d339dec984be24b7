// extern "C" entry points that orchestrate the kernels: coreset selection, projection, scoring (single GPU and the
// row-sharded phases), stand-alone upsample+blur.  Declarations and reference citations: include/cmdiad_b200.h.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace cmdb {

__global__ void __launch_bounds__(512) f32_to_f64_kernel(const float *__restrict__ x, long long n, double *__restrict__ z) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        z[i] = (double)x[i];
}

// keys[p] = (float_bits(min_val[p]) << 32) | global_row ; and the inverse
__global__ void pack_keys_kernel(const float *__restrict__ min_val, const long long *__restrict__ min_idx, int P,
                                 long long *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        const long long g = min_idx[i];
        keys[i] = g < 0 ? 0x7fffffffffffffffLL
                        : (long long)(((unsigned long long)__float_as_uint(min_val[i]) << 32) | (unsigned long long)g);
    }
}
__global__ void unpack_keys_kernel(const long long *__restrict__ keys, int P, int P_img, float *__restrict__ min_val,
                                   long long *__restrict__ min_idx, unsigned long long *s_key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        const unsigned long long k = (unsigned long long)keys[i];
        const float v = __uint_as_float((unsigned int)(k >> 32));
        min_val[i] = v;
        min_idx[i] = (long long)(k & 0xffffffffULL);
        atomicMax(s_key + i / P_img, ((unsigned long long)__float_as_uint(v) << 32) | (0xffffffffu - (unsigned int)(i % P_img)));
    }
}
// out[c] = bank[global_row - offset][c] if this shard owns the row, else 0
__global__ void contrib_rows_kernel(const float *__restrict__ bank, long long rows, long long row_offset, int dim,
                                    const long long *__restrict__ global_rows, const unsigned long long *__restrict__ keys,
                                    int n, float *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * dim; i += gridDim.x * blockDim.x) {
        const int r = i / dim, c = i - r * dim;
        long long g = global_rows ? global_rows[r] : (keys[r] == ~0ULL ? -1 : (long long)(keys[r] & 0xffffffffULL));
        const long long l = g - row_offset;
        out[i] = (g >= 0 && l >= 0 && l < rows) ? bank[(size_t)l * dim + c] : 0.f;
    }
}

void api_prefer_carveout_fuse();
void api_prefer_carveout() {
    api_prefer_carveout_fuse();
    CMDB_PREFER_MAX_SMEM(pack_keys_kernel);
    CMDB_PREFER_MAX_SMEM(unpack_keys_kernel);
    CMDB_PREFER_MAX_SMEM(contrib_rows_kernel);
    (void)cudaGetLastError();
}

static int check_score_args(cmdb_bank *b, const void *patch, int B, int P, const char *fn) {
    CMDB_REQUIRE(b && patch, CMDB_ERR_INVALID, "%s: NULL argument", fn);
    CMDB_REQUIRE(b->finalized, CMDB_ERR_STATE, "%s: call cmdb_bank_finalize first", fn);
    CMDB_REQUIRE(P >= 1 && P <= (1 << 20), CMDB_ERR_INVALID, "%s: P=%d out of range", fn, P);
    CMDB_REQUIRE(B >= 1 && B <= 4096, CMDB_ERR_INVALID, "%s: batch=%d out of range", fn, B);
    return CMDB_OK;
}

static int stage_alloc(cmdb_bank *b, int B, int P, int out_hw) {
    CMDB_CUDA(cudaSetDevice(b->device));
    return score_scratch_alloc(b, B, P, out_hw);
}

// device->host copy of the result block of a sub-batch, then scatter into the caller's buffers.  All images: ONE copy.
// A strided subset (sharded finish: this rank owns images img_first, img_first + img_step, ...): the scalar / per-patch
// prefix in one copy plus one map copy per owned image.
// bytes of the result block that a full-batch copy has to move (want: bit 0 = pre-blur maps, bit 1 = 8-bit maps)
static size_t out_block_extent(const cmdb_bank *b, int B, unsigned want) {
    const ScoreScratch &s = b->ss;
    size_t bytes = s.off_map_out + sizeof(float) * s.map_stride * B;
    if (want & 1u) bytes = s.off_map_pre + sizeof(float) * s.map_stride * B;
    if (want & 2u) bytes = s.off_map_u8 + s.map_stride * B;
    return bytes;
}
static unsigned want_mask_of(const cmdb_score_out *outs, int B, int first = 0, int step = 1) {
    unsigned w = 0;
    for (int i = first; i < B; i += step) w |= (outs[i].s_map_pre ? 1u : 0u) | (outs[i].s_map_u8 ? 2u : 0u);
    return w;
}

// host block of a slot -> the caller's buffers (maps = false: only the scalars and per-patch arrays of image i)
static void scatter_one(const cmdb_bank *b, const unsigned char *h, int i, int P, int out_hw, cmdb_score_out *out, bool maps = true) {
    const ScoreScratch &s = b->ss;
    const size_t npix = (size_t)out_hw * out_hw;
    TailResult tr;
    memcpy(&tr, h + sizeof(TailResult) * i, sizeof(tr));
    if (out->min_val) memcpy(out->min_val, h + s.off_min_val + sizeof(float) * (size_t)i * P, sizeof(float) * P);
    if (out->min_idx) memcpy(out->min_idx, h + s.off_min_idx + sizeof(long long) * (size_t)i * P, sizeof(long long) * P);
    if (maps) {
        if (out->s_map) memcpy(out->s_map, h + s.off_map_out + sizeof(float) * s.map_stride * i, sizeof(float) * npix);
        if (out->s_map_pre) memcpy(out->s_map_pre, h + s.off_map_pre + sizeof(float) * s.map_stride * i, sizeof(float) * npix);
        if (out->s_map_u8) memcpy(out->s_map_u8, h + s.off_map_u8 + s.map_stride * i, npix);
    }
    if (out->s) *out->s = tr.s;
    if (out->s_star) *out->s_star = tr.s_star;
    if (out->s_idx) *out->s_idx = tr.s_idx;
    if (out->w) *out->w = tr.w;
    if (out->m_star_knn) out->m_star_knn[0] = tr.knn0, out->m_star_knn[1] = tr.knn1;
    if (out->nn_idx)
        for (int k = 0; k < 3; ++k) out->nn_idx[k] = tr.nn_idx[k];
}
static void scatter_outputs(const cmdb_bank *b, const unsigned char *h, int B, int P, int out_hw, cmdb_score_out *outs,
                            int img_first = 0, int img_step = 1) {
    for (int i = img_first; i < B; i += img_step) scatter_one(b, h, i, P, out_hw, outs + i);
}

// sharded finish: device->host copy of the result block (all images: ONE copy; a strided subset -- this rank owns images
// img_first, img_first + img_step, ... -- the scalar / per-patch prefix in one copy plus one map copy per owned image),
// then scatter into the caller's buffers
static int copy_outputs(cmdb_bank *b, int B, int P, int out_hw, cmdb_score_out *outs, int img_first, int img_step) {
    ScoreScratch &s = b->ss;
    cudaStream_t st = b->stream;
    const size_t npix = (size_t)out_hw * out_hw;
    const unsigned want = want_mask_of(outs, B, img_first, img_step);
    if (img_step == 1 && img_first == 0) {
        CMDB_CUDA(cudaMemcpyAsync(s.out_block_host, s.out_block, out_block_extent(b, B, want), cudaMemcpyDeviceToHost, st));
    } else {
        CMDB_CUDA(cudaMemcpyAsync(s.out_block_host, s.out_block, s.off_map_out, cudaMemcpyDeviceToHost, st));
        for (int i = img_first; i < B; i += img_step) {
            const size_t o1 = s.off_map_out + sizeof(float) * s.map_stride * i, o2 = s.off_map_pre + sizeof(float) * s.map_stride * i;
            const size_t o3 = s.off_map_u8 + s.map_stride * i;
            CMDB_CUDA(cudaMemcpyAsync(s.out_block_host + o1, s.out_block + o1, sizeof(float) * npix, cudaMemcpyDeviceToHost, st));
            if (want & 1u) CMDB_CUDA(cudaMemcpyAsync(s.out_block_host + o2, s.out_block + o2, sizeof(float) * npix, cudaMemcpyDeviceToHost, st));
            if (want & 2u) CMDB_CUDA(cudaMemcpyAsync(s.out_block_host + o3, s.out_block + o3, npix, cudaMemcpyDeviceToHost, st));
        }
    }
    CMDB_CUDA(cudaStreamSynchronize(st));
    scatter_outputs(b, s.out_block_host, B, P, out_hw, outs, img_first, img_step);
    return CMDB_OK;
}

static int blur_batch(cmdb_bank *b, int B, int fh, int fw, int out_hw, int img_first = 0, int img_step = 1) {
    ScoreScratch &s = b->ss;
    CMDB_REQUIRE((size_t)out_hw * out_hw <= s.map_stride, CMDB_ERR_INVALID, "scoring: out_hw=%d larger than the scratch maps", out_hw);
    const int n_img = img_first < B ? (B - img_first + img_step - 1) / img_step : 0;
    return upsample_blur_launch(b->stream, n_img, img_first, img_step, s.map_stride, s.min_val, fh, fw, out_hw, s.map_pre,
                                s.map_out, s.map_u8, s.map_tmp, s.map_max, s.map_max + s.cap_b);
}

// ---------------------------------------------------------------------------------------------------------------
// late-fusion head (multiple_features.py:986-994): lambda scaling in float32, SGDOneClassSVM.score_samples in float64
//   X @ coef_ on 2 columns: dgemv evaluates fma(x0, c0, x1 * c1) for the [npix, 2] map matrix and the ddot of a single
//   row fma(x1, c1, x0 * c0) (OpenBLAS as shipped with numpy/scipy in this image; pinned by tests against sklearn).
// grid (pixel blocks, images)
// ---------------------------------------------------------------------------------------------------------------
struct FuseParams {
    int n_modal, B, npix;
    const float *maps[3];
    size_t map_stride[3];
    const TailResult *tails[3];
    float s_lambda[3], smap_lambda[3];
    double det_coef[3], det_off, seg_coef[3], seg_off;
    double *out_s;        // [B]
    float *out_s_modal;   // [B][3]
    double *out_map;      // [B][npix] or nullptr
    double *acc_map;      // [B][npix] slice of the device-side result store, or nullptr
    double *acc_s;        // [B] or nullptr
};

__global__ void __launch_bounds__(256) fuse_head_kernel(FuseParams p) {
    const int b = blockIdx.y;
    const float *m0 = p.maps[0] + (size_t)b * p.map_stride[0];
    const float *m1 = p.n_modal > 1 ? p.maps[1] + (size_t)b * p.map_stride[1] : nullptr;
    const float *m2 = p.n_modal > 2 ? p.maps[2] + (size_t)b * p.map_stride[2] : nullptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.npix; i += gridDim.x * blockDim.x) {
        // lambda * s_map in float32 (python float * float32 tensor), then sklearn's float64 copy
        const double x0 = (double)__fmul_rn(p.smap_lambda[0], m0[i]);
        double v;
        if (p.n_modal == 1) {
            v = __dmul_rn(x0, p.seg_coef[0]);
        } else {
            const double x1 = (double)__fmul_rn(p.smap_lambda[1], m1[i]);
            if (p.n_modal == 2) {
                v = __fma_rn(x0, p.seg_coef[0], __dmul_rn(x1, p.seg_coef[1]));
            } else {
                const double x2 = (double)__fmul_rn(p.smap_lambda[2], m2[i]);
                v = __fma_rn(x0, p.seg_coef[0], __fma_rn(x1, p.seg_coef[1], __dmul_rn(x2, p.seg_coef[2])));
            }
        }
        v = __dadd_rn(__dsub_rn(v, p.seg_off), p.seg_off);  // decision_function(X) + offset_
        if (p.out_map) p.out_map[(size_t)b * p.npix + i] = v;
        if (p.acc_map) p.acc_map[(size_t)b * p.npix + i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double x[3] = {0.0, 0.0, 0.0};
        for (int m = 0; m < p.n_modal; ++m) {
            const float sm = __fmul_rn(p.s_lambda[m], p.tails[m][b].s);
            p.out_s_modal[b * 3 + m] = sm;
            x[m] = (double)sm;
        }
        double v = __dmul_rn(x[0], p.det_coef[0]);  // a single row goes through ddot: sequential fma over the columns
        for (int m = 1; m < p.n_modal; ++m) v = __fma_rn(x[m], p.det_coef[m], v);
        v = __dadd_rn(__dsub_rn(v, p.det_off), p.det_off);
        p.out_s[b] = v;
        if (p.acc_s) p.acc_s[b] = v;
    }
}

void api_prefer_carveout_fuse() { CMDB_PREFER_MAX_SMEM(fuse_head_kernel); }

}  // namespace cmdb

using namespace cmdb;

extern "C" {

int cmdb_project(cmdb_bank *b, const int32_t *indptr, const int32_t *indices, const double *data, int d_proj, int64_t row0,
                 int64_t n_rows, double *out_host) {
    CMDB_REQUIRE(b && out_host && row0 >= 0 && n_rows > 0 && row0 + n_rows <= b->rows, CMDB_ERR_INVALID,
                 "cmdb_project: bad row range");
    CMDB_REQUIRE(d_proj > 0, CMDB_ERR_INVALID, "cmdb_project: d_proj must be positive");
    CMDB_CUDA(cudaSetDevice(b->device));
    double *z = nullptr;
    CMDB_CUDA(cudaMalloc(&z, sizeof(double) * (size_t)n_rows * d_proj));
    int rc = project_rows(b, b->data + row0 * b->dim, n_rows, b->dim, indptr, indices, data, d_proj, z);
    cudaError_t e = cudaSuccess;
    if (rc == CMDB_OK) e = cudaMemcpy(out_host, z, sizeof(double) * (size_t)n_rows * d_proj, cudaMemcpyDeviceToHost);
    cudaFree(z);
    if (rc != CMDB_OK) return rc;
    CMDB_CUDA(e);
    return CMDB_OK;
}

static int coreset_impl(cmdb_bank *b, int64_t n_select, const int32_t *indptr, const int32_t *indices, const double *data,
                        int d_proj, int dtype_mode, int64_t *out_idx_host, const int64_t *force_idx, void *out_min_last,
                        ShardCtx *sh = nullptr) {
    CMDB_REQUIRE(b && out_idx_host, CMDB_ERR_INVALID, "cmdb_coreset_select: NULL argument");
    CMDB_REQUIRE(b->rows > 0, CMDB_ERR_STATE, "cmdb_coreset_select: bank is empty");
    const int64_t n_total = sh ? sh->n_total : b->rows;
    CMDB_REQUIRE(n_select >= 1 && n_select <= n_total, CMDB_ERR_INVALID, "cmdb_coreset_select: n_select=%lld not in [1,%lld]",
                 (long long)n_select, (long long)n_total);
    CMDB_REQUIRE(d_proj >= 0 && d_proj <= b->dim, CMDB_ERR_INVALID,
                 "cmdb_coreset_select: d_proj=%d exceeds dim=%d (sklearn raises ValueError here; pass 0 to skip the projection)",
                 d_proj, b->dim);
    CMDB_CUDA(cudaSetDevice(b->device));
    const int d = d_proj > 0 ? d_proj : b->dim;
    // the shard's projected rows start (row_offset*d) & 3 elements past a 32-byte boundary (see coreset_greedy_dev)
    const int pad = (int)((b->row_offset * d) & 3);
    double *z_alloc = nullptr;
    CMDB_CUDA(cudaMalloc(&z_alloc, sizeof(double) * ((size_t)b->rows * d + 4)));
    double *z = z_alloc + pad;
    int rc;
    if (d_proj > 0) {
        rc = project_rows(b, b->data, b->rows, b->dim, indptr, indices, data, d_proj, z);
    } else {
        // features.py:369-370: projection skipped, the float32 bank itself is used
        f32_to_f64_kernel<<<b->num_sms * 4, 512, 0, b->stream>>>(b->data, (long long)b->rows * d, z);
        rc = cudaGetLastError() == cudaSuccess ? CMDB_OK : CMDB_ERR_CUDA;
    }
    if (rc == CMDB_OK) rc = coreset_greedy_dev(b, z, b->rows, d, n_select, dtype_mode, out_idx_host, force_idx, out_min_last, sh);
    cudaFree(z_alloc);
    return rc;
}

int cmdb_coreset_select(cmdb_bank *b, int64_t n_select, const int32_t *indptr, const int32_t *indices, const double *data,
                        int d_proj, int dtype_mode, int64_t *out_idx_host) {
    return coreset_impl(b, n_select, indptr, indices, data, d_proj, dtype_mode, out_idx_host, nullptr, nullptr);
}

// test hook (not in the public header): teacher forcing and the final min-distance vector
int cmdb_coreset_select_debug(cmdb_bank *b, int64_t n_select, const int32_t *indptr, const int32_t *indices,
                              const double *data, int d_proj, int dtype_mode, int64_t *out_idx_host,
                              const int64_t *force_idx_host, void *out_min_last_host) {
    return coreset_impl(b, n_select, indptr, indices, data, d_proj, dtype_mode, out_idx_host, force_idx_host,
                        out_min_last_host);
}

size_t cmdb_coreset_mailbox_bytes(int world, int d_proj_max) {
    (void)world, (void)d_proj_max;
    return kCommHeaderBytes;  // flags + key slots; the replica of the projected bank follows (cmdb_coreset_comm_bytes)
}

size_t cmdb_coreset_comm_bytes(int world, int d_proj, int64_t n_total_rows, int dtype_mode) {
    (void)world;
    const size_t es = dtype_mode == CMDB_CORESET_FP16 ? sizeof(__half) : sizeof(double);
    return kCommHeaderBytes + es * (size_t)std::max<int64_t>(n_total_rows, 1) * (size_t)std::max(d_proj, 1) + 256;
}

int cmdb_coreset_select_sharded(cmdb_bank *b, cmdb_comm *comm, int64_t n_total_rows, int64_t n_select, const int32_t *indptr,
                                const int32_t *indices, const double *data, int d_proj, int dtype_mode, const double *z0_host,
                                int64_t *out_idx_host) {
    CMDB_REQUIRE(b && comm && z0_host && out_idx_host, CMDB_ERR_INVALID, "cmdb_coreset_select_sharded: NULL argument");
    CMDB_REQUIRE(d_proj > 0, CMDB_ERR_UNSUPPORTED, "cmdb_coreset_select_sharded: needs a projection (d_proj > 0)");
    ShardCtx sh{};
    size_t mb_bytes = 0;
    unsigned char *local = nullptr;
    CMDB_CHECK(comm_info(comm, &sh.rank, &sh.world, &local, sh.peers, &mb_bytes));
    sh.row_offset = b->row_offset, sh.n_total = n_total_rows, sh.z0_host = z0_host;
    sh.comm_bytes = mb_bytes;
    CMDB_REQUIRE(mb_bytes >= cmdb_coreset_comm_bytes(sh.world, d_proj, n_total_rows, dtype_mode), CMDB_ERR_INVALID,
                 "cmdb_coreset_select_sharded: the peer buffer has %zu bytes, need %zu (cmdb_coreset_comm_bytes)", mb_bytes,
                 cmdb_coreset_comm_bytes(sh.world, d_proj, n_total_rows, dtype_mode));
    CMDB_REQUIRE(b->row_offset + b->rows <= n_total_rows, CMDB_ERR_INVALID, "cmdb_coreset_select_sharded: shard exceeds n_total_rows");
    CMDB_CUDA(cudaSetDevice(b->device));
    return coreset_impl(b, n_select, indptr, indices, data, d_proj, dtype_mode, out_idx_host, nullptr, nullptr, &sh);
}

int cmdb_coreset_rownorms(int device, const void *z_host, const void *last_host, int64_t n_rows, int d, int dtype_mode,
                          void *out_host) {
    return coreset_rownorms(device, z_host, last_host, n_rows, d, dtype_mode, out_host);
}

// Enqueue one sub-batch (<= score_max_batch images) on buffer slot `slot`: staging, GEMM + certificate, maps, re-weighting
// and the device->host copies of the results (on the d2h stream, the maps as soon as the blur is done).  No host sync.
// A submitted call takes the next result slot (kResultSlots of them) and the next compute lane (two, alternating).
struct SlotPick {
    int lane, rslot;
};
static SlotPick take_slot(cmdb_bank *b) {
    const SlotPick p{b->next_lane, b->next_slot};
    b->next_lane ^= 1;
    b->next_slot = (b->next_slot + 1) % kResultSlots;
    return p;
}

static int submit_sub_batch(cmdb_bank *b, const float *src, int is_device, int bc, int P, int fh, int fw, int out_hw,
                            unsigned want, SlotPick pick, bool host_maps = true) {
    const int slot = pick.rslot, lane = pick.lane;
#define CMDB_MARK(i)                                                   \
    do {                                                               \
        if (b->timing) CMDB_CUDA(cudaEventRecord(b->timing == 2 ? b->ev_tl[lane][i] : b->ev[i], b->stream)); \
    } while (0)
    ScoreScratch &s = b->ss;
    CMDB_CHECK(stage_alloc(b, bc, P, out_hw));
    score_select_slot(b, lane, slot);
    cudaStream_t st = b->stream;  // the lane's stream (only valid after the selection)
    CMDB_MARK(CMDB_T_STAGE_IN);
    // this slot's q_f32 was last read by the batch submitted three calls ago
    CMDB_CHECK(score_local_min(b, src, is_device, bc, P, CMDB_T_GEMM, CMDB_T_REFINE, b->ev_compute[slot]));
    CMDB_MARK(CMDB_T_MAP);
    CMDB_CHECK(blur_batch(b, bc, fh, fw, out_hw));
    // min_val / min_idx / maps do not depend on the re-weighting: their copy overlaps it
    CMDB_CUDA(cudaEventRecord(b->ev_chunk[0], st));
    CMDB_CUDA(cudaStreamWaitEvent(b->d2h_stream, b->ev_chunk[0], 0));
    CMDB_CUDA(cudaMemcpyAsync(s.out_block_host + s.off_min_val, s.out_block + s.off_min_val,
                              (host_maps ? out_block_extent(b, bc, want) : s.off_map_out) - s.off_min_val, cudaMemcpyDeviceToHost,
                              b->d2h_stream));
    CMDB_MARK(CMDB_T_REWEIGHT);
    CMDB_CHECK(score_reweight(b, bc, P, true));
    CMDB_MARK(CMDB_T_OUT);
    CMDB_CUDA(cudaEventRecord(b->ev_compute[slot], st));
    CMDB_CUDA(cudaStreamWaitEvent(b->d2h_stream, b->ev_compute[slot], 0));
    CMDB_CUDA(cudaMemcpyAsync(s.out_block_host, s.out_block, s.off_min_val, cudaMemcpyDeviceToHost, b->d2h_stream));
    CMDB_CUDA(cudaEventRecord(b->ev_done[slot], b->d2h_stream));
    if (b->timing) CMDB_CUDA(cudaEventRecord(b->timing == 2 ? b->ev_tl[lane][CMDB_T_COUNT] : b->ev[CMDB_T_COUNT], b->d2h_stream));
#undef CMDB_MARK
    cmdb_bank::Pending &pd = b->pending[slot];
    pd.active = true, pd.B = bc, pd.P = P, pd.out_hw = out_hw, pd.want = want, pd.host_maps = host_maps;
    return CMDB_OK;
}

static int wait_slot(cmdb_bank *b, int slot, cmdb_score_out *outs) {
    cmdb_bank::Pending &pd = b->pending[slot];
    CMDB_CUDA(cudaEventSynchronize(b->ev_done[slot]));
    pd.active = false;
    scatter_outputs(b, b->ss.out_block_host_buf[slot], pd.B, pd.P, pd.out_hw, outs);
    if (b->timing == 1) b->ev_valid = true;
    return CMDB_OK;
}

static int check_batch_args(cmdb_bank *b, const float *patches, int B, int P, int fh, int fw, int out_hw, const char *fn) {
    CMDB_CHECK(check_score_args(b, patches, B, P, fn));
    CMDB_REQUIRE(fh > 0 && fw > 0 && fh * fw == P, CMDB_ERR_INVALID, "%s: feature_map_dims %dx%d != P=%d", fn, fh, fw, P);
    CMDB_REQUIRE(out_hw >= 8 && out_hw <= 256, CMDB_ERR_INVALID, "%s: out_hw=%d not in [8,256]", fn, out_hw);
    CMDB_CUDA(cudaSetDevice(b->device));
    if (b->ss.map_stride && (size_t)out_hw * out_hw != b->ss.map_stride) {  // map stride is fixed per scratch
        CMDB_REQUIRE(!b->any_pending(), CMDB_ERR_STATE,
                     "%s: out_hw changes while a submitted batch is outstanding; wait for it first", fn);
        score_scratch_free(b);
    }
    return CMDB_OK;
}

int cmdb_score_batch(cmdb_bank *b, const float *patches, int B, int P, int fh, int fw, int out_hw, int patch_is_device,
                     cmdb_score_out *outs) {
    CMDB_REQUIRE(outs, CMDB_ERR_INVALID, "cmdb_score_batch: outs is NULL");
    CMDB_CHECK(check_batch_args(b, patches, B, P, fh, fw, out_hw, "cmdb_score_batch"));
    const int bc_max = score_max_batch(b);
    // sub-batches alternate between the two lanes: sub-batch k + 1 is enqueued before the host waits for sub-batch k
    int prev_slot = -1, prev_b0 = 0;
    for (int b0 = 0; b0 < B; b0 += bc_max) {
        const int bc = std::min(bc_max, B - b0);
        CMDB_REQUIRE(!b->pending[b->next_slot].active, CMDB_ERR_STATE,
                     "cmdb_score_batch: all result slots hold submitted batches; wait for one first");
        const SlotPick pick = take_slot(b);
        const int slot = pick.rslot;
        CMDB_CHECK(submit_sub_batch(b, patches + (size_t)b0 * P * b->dim, patch_is_device, bc, P, fh, fw, out_hw,
                                    want_mask_of(outs + b0, bc), pick));
        if (prev_slot >= 0) CMDB_CHECK(wait_slot(b, prev_slot, outs + prev_b0));
        prev_slot = slot, prev_b0 = b0;
    }
    if (prev_slot >= 0) CMDB_CHECK(wait_slot(b, prev_slot, outs + prev_b0));
    return CMDB_OK;
}

int cmdb_score_batch_submit(cmdb_bank *b, const float *patches, int B, int P, int fh, int fw, int out_hw, int patch_is_device,
                            unsigned want_maps, int64_t *out_ticket) {
    CMDB_REQUIRE(out_ticket, CMDB_ERR_INVALID, "cmdb_score_batch_submit: out_ticket is NULL");
    CMDB_CHECK(check_batch_args(b, patches, B, P, fh, fw, out_hw, "cmdb_score_batch_submit"));
    CMDB_REQUIRE(B <= score_max_batch(b), CMDB_ERR_INVALID, "cmdb_score_batch_submit: batch=%d exceeds the per-call limit %d", B,
                 score_max_batch(b));
    CMDB_REQUIRE(!b->pending[b->next_slot].active, CMDB_ERR_STATE,
                 "cmdb_score_batch_submit: three batches are already outstanding; call cmdb_score_batch_wait first");
    const SlotPick pick = take_slot(b);
    const int slot = pick.rslot;
    CMDB_CHECK(submit_sub_batch(b, patches, patch_is_device, B, P, fh, fw, out_hw, want_maps & 3u, pick));
    b->pending[slot].ticket = ++b->ticket_counter;
    *out_ticket = b->pending[slot].ticket;
    return CMDB_OK;
}

int cmdb_score_batch_wait(cmdb_bank *b, int64_t ticket, cmdb_score_out *outs) {
    CMDB_REQUIRE(b && outs, CMDB_ERR_INVALID, "cmdb_score_batch_wait: NULL argument");
    for (int slot = 0; slot < kResultSlots; ++slot)
        if (b->pending[slot].active && b->pending[slot].ticket == ticket) {
            CMDB_CUDA(cudaSetDevice(b->device));
            return wait_slot(b, slot, outs);
        }
    set_error("cmdb_score_batch_wait: ticket %lld is not outstanding", (long long)ticket);
    return CMDB_ERR_STATE;
}

int cmdb_score(cmdb_bank *b, const float *patch, int P, int fh, int fw, int out_hw, int patch_is_device,
               cmdb_score_out *out) {
    return cmdb_score_batch(b, patch, 1, P, fh, fw, out_hw, patch_is_device, out);
}

static int check_shard_batch(cmdb_bank *b, int B, const char *fn) {
    CMDB_REQUIRE(B >= 1 && B <= score_max_batch(b), CMDB_ERR_INVALID, "%s: batch=%d exceeds the per-call limit %d", fn, B,
                 score_max_batch(b));
    return CMDB_OK;
}

int cmdb_score_shard_min(cmdb_bank *b, const float *patches, int B, int P, int patch_is_device, int out_hw,
                         int64_t *keys_device) {
    CMDB_CHECK(check_score_args(b, patches, B, P, "cmdb_score_shard_min"));
    CMDB_CHECK(check_shard_batch(b, B, "cmdb_score_shard_min"));
    CMDB_REQUIRE(keys_device, CMDB_ERR_INVALID, "cmdb_score_shard_min: keys_device is NULL");
    CMDB_REQUIRE(out_hw >= 8 && out_hw <= 256, CMDB_ERR_INVALID, "cmdb_score_shard_min: out_hw=%d not in [8,256]", out_hw);
    CMDB_CUDA(cudaSetDevice(b->device));
    {
        // the round runs on lane next_lane (what cmdb_bank_stream returns before this call) with result slot next_slot; with
        // the submit / wait finish up to three rounds may be outstanding (three result slots on the two lanes)
        const bool busy = b->any_pending();
        CMDB_REQUIRE(!busy || (size_t)out_hw * out_hw == b->ss.map_stride, CMDB_ERR_STATE,
                     "cmdb_score_shard_min: out_hw changes while a submitted round is outstanding; wait for it first");
        if (!busy && (size_t)out_hw * out_hw != b->ss.map_stride) score_scratch_free(b);
        const int slot = b->next_slot, lane = b->next_lane;   // advanced by the finish call of the round
        CMDB_REQUIRE(!b->pending[slot].active, CMDB_ERR_STATE,
                     "cmdb_score_shard_min: all result slots hold submitted rounds; wait for one first");
        CMDB_CHECK(stage_alloc(b, B, P, out_hw));  // fails with CMDB_ERR_STATE if the scratch would have to grow
        score_select_slot(b, lane, slot);
        b->shard_slot = slot, b->shard_lane = lane;
        CMDB_CHECK(score_local_min(b, patches, patch_is_device, B, P, -1, -1, b->ev_compute[slot]));
    }
    pack_keys_kernel<<<(B * P + 255) / 256, 256, 0, b->stream>>>(b->ss.min_val, b->ss.min_idx, B * P, (long long *)keys_device);
    CMDB_CUDA(cudaGetLastError());
    return CMDB_OK;  // stream-ordered on the handle's stream (cmdb_bank_stream): run the collective there
}

int cmdb_score_shard_select(cmdb_bank *b, const int64_t *reduced_keys_device, int B, int P, float *m_star_contrib_device) {
    CMDB_CHECK(check_score_args(b, reduced_keys_device, B, P, "cmdb_score_shard_select"));
    CMDB_REQUIRE(m_star_contrib_device && B * P <= b->ss.cap_p && B <= b->ss.cap_b, CMDB_ERR_INVALID,
                 "cmdb_score_shard_select: bad arguments (call cmdb_score_shard_min first)");
    CMDB_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = b->stream;
    CMDB_CUDA(cudaMemsetAsync(b->ss.s_key, 0, sizeof(unsigned long long) * B, st));
    unpack_keys_kernel<<<(B * P + 255) / 256, 256, 0, st>>>((const long long *)reduced_keys_device, B * P, P, b->ss.min_val,
                                                            b->ss.min_idx, b->ss.s_key);
    CMDB_CHECK(score_select(b, B, P, true));  // m_star rows this rank owns, zeros elsewhere
    CMDB_CUDA(cudaMemcpyAsync(m_star_contrib_device, b->ss.m_star, sizeof(float) * (size_t)B * b->dim, cudaMemcpyDeviceToDevice, st));
    return CMDB_OK;
}

int cmdb_score_shard_topk(cmdb_bank *b, const float *m_star_device, int B, int P, int64_t *topk_keys_device) {
    CMDB_CHECK(check_score_args(b, m_star_device, B, P, "cmdb_score_shard_topk"));
    CMDB_REQUIRE(topk_keys_device && B <= b->ss.cap_b, CMDB_ERR_INVALID, "cmdb_score_shard_topk: bad arguments");
    CMDB_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = b->stream;
    CMDB_CUDA(cudaMemcpyAsync(b->ss.m_star, m_star_device, sizeof(float) * (size_t)B * b->dim, cudaMemcpyDeviceToDevice, st));
    CMDB_CHECK(score_reweight(b, B, P, false));
    CMDB_CUDA(cudaMemcpyAsync(topk_keys_device, b->ss.top3, sizeof(long long) * 3 * B, cudaMemcpyDeviceToDevice, st));
    return CMDB_OK;
}

int cmdb_score_shard_nn(cmdb_bank *b, const int64_t *gathered_keys_device, int n_ranks, int B, float *nn_rows_contrib_device) {
    CMDB_CHECK(check_score_args(b, gathered_keys_device, B, 1, "cmdb_score_shard_nn"));
    CMDB_REQUIRE(nn_rows_contrib_device && n_ranks >= 1 && n_ranks <= b->ss.n_topk_blocks && B <= b->ss.cap_b, CMDB_ERR_INVALID,
                 "cmdb_score_shard_nn: bad arguments");
    CMDB_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = b->stream;
    CMDB_CUDA(cudaMemcpyAsync(b->ss.topk_keys, gathered_keys_device, sizeof(long long) * 3 * (size_t)n_ranks * B,
                              cudaMemcpyDeviceToDevice, st));
    CMDB_CHECK(score_merge_top3(b, n_ranks, B));
    contrib_rows_kernel<<<8 * B, 256, 0, st>>>(b->data, b->fin_rows, b->row_offset, b->dim, nullptr, b->ss.top3, 3 * B,
                                               nn_rows_contrib_device);
    CMDB_CUDA(cudaGetLastError());
    return CMDB_OK;
}

int cmdb_score_shard_finish(cmdb_bank *b, const float *nn_rows_device, int B, int P, int fh, int fw, int out_hw,
                            int img_first, int img_step, cmdb_score_out *outs) {
    CMDB_CHECK(check_score_args(b, nn_rows_device, B, P, "cmdb_score_shard_finish"));
    CMDB_REQUIRE(outs && fh > 0 && fw > 0 && fh * fw == P && B * P <= b->ss.cap_p && B <= b->ss.cap_b, CMDB_ERR_INVALID,
                 "cmdb_score_shard_finish: bad arguments");
    CMDB_REQUIRE((size_t)out_hw * out_hw == b->ss.map_stride, CMDB_ERR_INVALID,
                 "cmdb_score_shard_finish: out_hw differs from cmdb_score_shard_min");
    CMDB_REQUIRE(img_first >= 0 && img_step >= 1, CMDB_ERR_INVALID, "cmdb_score_shard_finish: bad image subset");
    CMDB_CUDA(cudaSetDevice(b->device));
    CMDB_CUDA(cudaMemcpyAsync(b->ss.nn_rows, nn_rows_device, sizeof(float) * 3 * (size_t)B * b->dim, cudaMemcpyDeviceToDevice,
                              b->stream));
    CMDB_CHECK(score_final(b, B));
    CMDB_CHECK(blur_batch(b, B, fh, fw, out_hw, img_first, img_step));
    return copy_outputs(b, B, P, out_hw, outs, img_first, img_step);
}

// ---- sharded rounds with the replicated neighbour table: min -> [MIN all-reduce] -> lookup -> [SUM all-reduce] -> finish ----

int cmdb_score_shard_lookup(cmdb_bank *b, const int64_t *reduced_keys_device, int B, int P, float *knn_d2_contrib_device) {
    CMDB_CHECK(check_score_args(b, reduced_keys_device, B, P, "cmdb_score_shard_lookup"));
    CMDB_REQUIRE(knn_d2_contrib_device && B * P <= b->ss.cap_p && B <= b->ss.cap_b, CMDB_ERR_INVALID,
                 "cmdb_score_shard_lookup: bad arguments (call cmdb_score_shard_min first)");
    CMDB_REQUIRE(b->knn_table && b->knn_rows >= b->row_offset + b->fin_rows, CMDB_ERR_STATE,
                 "cmdb_score_shard_lookup: install the replicated neighbour table first (cmdb_bank_set_knn_table)");
    CMDB_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = b->stream;
    CMDB_CUDA(cudaMemsetAsync(b->ss.s_key, 0, sizeof(unsigned long long) * B, st));
    unpack_keys_kernel<<<(B * P + 255) / 256, 256, 0, st>>>((const long long *)reduced_keys_device, B * P, P, b->ss.min_val,
                                                            b->ss.min_idx, b->ss.s_key);
    CMDB_CHECK(score_select(b, B, P, false));  // m_test, s*, s_idx, m_star_row (a global row: no bank access needed)
    return score_shard_lookup(b, B, knn_d2_contrib_device);
}

static int shard_finish_enqueue(cmdb_bank *b, const float *knn_d2_sum_device, int B, int P, int fh, int fw, int out_hw,
                                int img_first, int img_step, unsigned want_maps, int64_t *out_ticket);

int cmdb_score_shard_finish_submit(cmdb_bank *b, const float *knn_d2_sum_device, int B, int P, int fh, int fw, int out_hw,
                                   int img_first, int img_step, unsigned want_maps, int64_t *out_ticket) {
    return shard_finish_enqueue(b, knn_d2_sum_device, B, P, fh, fw, out_hw, img_first, img_step, want_maps, out_ticket);
}

static int shard_finish_enqueue(cmdb_bank *b, const float *knn_d2_sum_device, int B, int P, int fh, int fw, int out_hw,
                                int img_first, int img_step, unsigned want_maps, int64_t *out_ticket) {
    CMDB_CHECK(check_score_args(b, knn_d2_sum_device, B, P, "cmdb_score_shard_finish_submit"));
    CMDB_REQUIRE(out_ticket && fh > 0 && fw > 0 && fh * fw == P && B * P <= b->ss.cap_p && B <= b->ss.cap_b, CMDB_ERR_INVALID,
                 "cmdb_score_shard_finish_submit: bad arguments");
    CMDB_REQUIRE((size_t)out_hw * out_hw == b->ss.map_stride, CMDB_ERR_INVALID,
                 "cmdb_score_shard_finish_submit: out_hw differs from cmdb_score_shard_min");
    CMDB_REQUIRE(img_first >= 0 && img_step >= 1, CMDB_ERR_INVALID, "cmdb_score_shard_finish_submit: bad image subset");
    const int slot = b->shard_slot, lane = b->shard_lane;
    CMDB_REQUIRE(!b->pending[slot].active, CMDB_ERR_STATE, "cmdb_score_shard_finish_submit: this round's slot is still outstanding");
    CMDB_CUDA(cudaSetDevice(b->device));
    ScoreScratch &s = b->ss;
    cudaStream_t st = b->stream;
    CMDB_CHECK(score_shard_final(b, B, knn_d2_sum_device));
    CMDB_CHECK(blur_batch(b, B, fh, fw, out_hw, img_first, img_step));
    CMDB_CUDA(cudaEventRecord(b->ev_compute[slot], st));
    CMDB_CUDA(cudaStreamWaitEvent(b->d2h_stream, b->ev_compute[slot], 0));
    // scalar / per-patch prefix of ALL images in one copy, then the maps of the images this rank finished
    const size_t npix = (size_t)out_hw * out_hw;
    want_maps &= 3u;
    if (img_first == 0 && img_step == 1) {
        CMDB_CUDA(cudaMemcpyAsync(s.out_block_host, s.out_block, out_block_extent(b, B, want_maps), cudaMemcpyDeviceToHost, b->d2h_stream));
    } else {
        CMDB_CUDA(cudaMemcpyAsync(s.out_block_host, s.out_block, s.off_map_out, cudaMemcpyDeviceToHost, b->d2h_stream));
        for (int i = img_first; i < B; i += img_step) {
            const size_t o1 = s.off_map_out + sizeof(float) * s.map_stride * i, o2 = s.off_map_pre + sizeof(float) * s.map_stride * i;
            const size_t o3 = s.off_map_u8 + s.map_stride * i;
            CMDB_CUDA(cudaMemcpyAsync(s.out_block_host + o1, s.out_block + o1, sizeof(float) * npix, cudaMemcpyDeviceToHost, b->d2h_stream));
            if (want_maps & 1u) CMDB_CUDA(cudaMemcpyAsync(s.out_block_host + o2, s.out_block + o2, sizeof(float) * npix, cudaMemcpyDeviceToHost, b->d2h_stream));
            if (want_maps & 2u) CMDB_CUDA(cudaMemcpyAsync(s.out_block_host + o3, s.out_block + o3, npix, cudaMemcpyDeviceToHost, b->d2h_stream));
        }
    }
    CMDB_CUDA(cudaEventRecord(b->ev_done[slot], b->d2h_stream));
    cmdb_bank::Pending &pd = b->pending[slot];
    pd.active = true, pd.B = B, pd.P = P, pd.out_hw = out_hw, pd.want = want_maps, pd.host_maps = true;
    pd.img_first = img_first, pd.img_step = img_step;
    pd.ticket = ++b->ticket_counter;
    b->next_slot = (slot + 1) % kResultSlots, b->next_lane = lane ^ 1;
    *out_ticket = pd.ticket;
    return CMDB_OK;
}

// ---- NCCL-free sharded rounds: the two exchanges go through peer-mapped memory (cmdb_comm), fused into the kernels ----

int cmdb_bank_attach_comm(cmdb_bank *b, cmdb_comm *comm) {
    CMDB_REQUIRE(b, CMDB_ERR_INVALID, "cmdb_bank_attach_comm: bank is NULL");
    CMDB_REQUIRE(!b->any_pending(), CMDB_ERR_STATE, "cmdb_bank_attach_comm: a submitted round is outstanding");
    b->comm = comm;
    if (!comm) return CMDB_OK;
    int rank = 0, world = 1;
    size_t bytes = 0;
    unsigned char *local = nullptr, *peers[kMaxRanks];
    CMDB_CHECK(comm_info(comm, &rank, &world, &local, peers, &bytes));
    CMDB_REQUIRE(bytes >= kCommHeaderBytes, CMDB_ERR_INVALID, "cmdb_bank_attach_comm: the peer buffer has %zu bytes, need >= %zu", bytes,
                 (size_t)kCommHeaderBytes);
    CMDB_CUDA(cudaSetDevice(b->device));
    if (!b->shard_ctr) {
        CMDB_CUDA(cudaMalloc(&b->shard_ctr, 4 * sizeof(unsigned int)));
        CMDB_CUDA(cudaMemset(b->shard_ctr, 0, 4 * sizeof(unsigned int)));
        CMDB_CUDA(cudaMalloc(&b->shard_d2, sizeof(float) * 2 * kShardSlots * kShardD2Cap));
        CMDB_CUDA(cudaHostAlloc(&b->shard_abort_host, sizeof(unsigned int), cudaHostAllocMapped));
        *b->shard_abort_host = 0u;
        CMDB_CUDA(cudaHostGetDevicePointer(&b->shard_abort_dev, b->shard_abort_host, 0));
    }
    return CMDB_OK;
}

int cmdb_score_shard_round_submit(cmdb_bank *b, const float *patches, int B, int P, int fh, int fw, int out_hw, int patch_is_device,
                                  int img_first, int img_step, unsigned want_maps, int64_t *out_ticket) {
    CMDB_CHECK(check_score_args(b, patches, B, P, "cmdb_score_shard_round_submit"));
    CMDB_CHECK(check_shard_batch(b, B, "cmdb_score_shard_round_submit"));
    CMDB_REQUIRE(b->comm, CMDB_ERR_STATE, "cmdb_score_shard_round_submit: attach the peer buffers first (cmdb_bank_attach_comm)");
    CMDB_REQUIRE(b->knn_table && b->knn_rows >= b->row_offset + b->fin_rows, CMDB_ERR_STATE,
                 "cmdb_score_shard_round_submit: install the replicated neighbour table first (cmdb_bank_set_knn_table)");
    CMDB_REQUIRE(out_ticket && fh > 0 && fw > 0 && fh * fw == P && out_hw >= 8 && out_hw <= 256, CMDB_ERR_INVALID,
                 "cmdb_score_shard_round_submit: bad arguments");
    int rank = 0, world = 1;
    size_t bytes = 0;
    unsigned char *local = nullptr;
    PeerPtrs peers{};
    CMDB_CHECK(comm_info(b->comm, &rank, &world, &local, peers.p, &bytes));
    CMDB_CUDA(cudaSetDevice(b->device));
    const bool busy = b->any_pending();
    if (!busy && (size_t)out_hw * out_hw != b->ss.map_stride) score_scratch_free(b);
    CMDB_REQUIRE(!busy || (size_t)out_hw * out_hw == b->ss.map_stride, CMDB_ERR_STATE,
                 "cmdb_score_shard_round_submit: out_hw changes while a submitted round is outstanding; wait for it first");
    const int slot = b->next_slot, lane = b->next_lane;   // advanced by shard_finish_enqueue
    CMDB_REQUIRE(!b->pending[slot].active, CMDB_ERR_STATE,
                 "cmdb_score_shard_round_submit: three submitted rounds are outstanding on this handle; wait for one first");
    CMDB_CHECK(stage_alloc(b, B, P, out_hw));
    score_select_slot(b, lane, slot);
    b->shard_slot = slot, b->shard_lane = lane;
    CMDB_CHECK(score_local_min(b, patches, patch_is_device, B, P, -1, -1, b->ev_compute[slot]));
    const unsigned long long epoch = comm_next_score_epoch(b->comm);  // per buffer, not per bank: several banks may share it
    const int xslot = (int)(epoch % kShardSlots);
    CMDB_CHECK(score_shard_exchange_keys(b, B, P, peers, world, rank, xslot, epoch));
    CMDB_CHECK(score_select(b, B, P, false));
    float *contrib = b->shard_d2 + xslot * kShardD2Cap, *d2_sum = b->shard_d2 + (kShardSlots + xslot) * kShardD2Cap;
    CMDB_CHECK(score_shard_lookup(b, B, contrib));
    CMDB_CHECK(score_shard_exchange_d2(b, B, peers, world, rank, xslot, epoch, contrib, d2_sum));
    return shard_finish_enqueue(b, d2_sum, B, P, fh, fw, out_hw, img_first, img_step, want_maps, out_ticket);
}

int cmdb_score_shard_wait(cmdb_bank *b, int64_t ticket, cmdb_score_out *outs) {
    CMDB_REQUIRE(b && outs, CMDB_ERR_INVALID, "cmdb_score_shard_wait: NULL argument");
    for (int slot = 0; slot < kResultSlots; ++slot) {
        cmdb_bank::Pending &pd = b->pending[slot];
        if (!pd.active || pd.ticket != ticket) continue;
        CMDB_CUDA(cudaSetDevice(b->device));
        CMDB_CUDA(cudaEventSynchronize(b->ev_done[slot]));
        pd.active = false;
        if (b->shard_abort_host && *b->shard_abort_host) {
            set_error("cmdb_score_shard_wait: a peer rank did not arrive at the exchange of this round within the timeout (all ranks "
                      "must submit the same rounds in the same order)");
            return CMDB_ERR_CUDA;
        }
        // maps: the images this rank finished; the scalars and per-patch arrays are replicated, so every image gets those
        for (int i = 0; i < pd.B; ++i) {
            const bool mine = i >= pd.img_first && (i - pd.img_first) % pd.img_step == 0;
            scatter_one(b, b->ss.out_block_host_buf[slot], i, pd.P, pd.out_hw, outs + i, mine);
        }
        return CMDB_OK;
    }
    set_error("cmdb_score_shard_wait: ticket %lld is not outstanding", (long long)ticket);
    return CMDB_ERR_STATE;
}

// plain host -> device copy on the handle's stream (stream-ordered with the phases that follow); used by the Python side to
// stage its slice of a round's queries without involving torch's pinned-memory bookkeeping
int cmdb_bank_stage_h2d(cmdb_bank *b, void *dst_device, const void *src_host, size_t bytes) {
    CMDB_REQUIRE(b && dst_device && src_host, CMDB_ERR_INVALID, "cmdb_bank_stage_h2d: NULL argument");
    if (bytes == 0) return CMDB_OK;
    CMDB_CUDA(cudaSetDevice(b->device));
    // on the copy stream, so that it overlaps the kernels already queued on the compute stream; the compute stream picks the
    // data up through an event.  (The caller guarantees that dst_device is not read by earlier, still running work.)
    CMDB_CUDA(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, b->copy_stream));
    CMDB_CUDA(cudaEventRecord(b->ev_stage, b->copy_stream));
    for (auto st : b->lane_stream) CMDB_CUDA(cudaStreamWaitEvent(st, b->ev_stage, 0));  // whichever lane runs the round
    return CMDB_OK;
}

int cmdb_bank_read_device(cmdb_bank *b, int64_t row0, int64_t n_rows, float *out_device) {
    CMDB_REQUIRE(b && out_device && row0 >= 0 && n_rows >= 0 && row0 + n_rows <= b->rows, CMDB_ERR_INVALID,
                 "cmdb_bank_read_device: rows [%lld,%lld) outside [0,%lld)", (long long)row0, (long long)(row0 + n_rows),
                 b ? (long long)b->rows : 0LL);
    if (n_rows == 0) return CMDB_OK;
    CMDB_CUDA(cudaSetDevice(b->device));
    CMDB_CUDA(cudaMemcpyAsync(out_device, b->data + row0 * b->dim, sizeof(float) * (size_t)n_rows * b->dim, cudaMemcpyDeviceToDevice,
                              b->stream));
    CMDB_CUDA(cudaStreamSynchronize(b->stream));
    return CMDB_OK;
}

// test hook (not in the public header): the GEMM epilogue's per-CTA top-2 lists of the last call, [n_cta][n_q][4] floats
int cmdb_debug_read_candidates(cmdb_bank *b, float *out_host, int n_cta, int n_q) {
    CMDB_REQUIRE(b && out_host && b->ss.cand && n_cta >= 1 && n_cta <= 2 * b->num_sms && n_q >= 1 && n_q <= b->ss.cap_p,
                 CMDB_ERR_INVALID, "cmdb_debug_read_candidates: bad arguments");
    CMDB_CUDA(cudaSetDevice(b->device));
    CMDB_CUDA(cudaStreamSynchronize(b->stream));
    CMDB_CUDA(cudaMemcpy2D(out_host, sizeof(float4) * n_q, b->ss.cand, sizeof(float4) * b->ss.cap_p, sizeof(float4) * n_q, n_cta,
                           cudaMemcpyDeviceToHost));
    return CMDB_OK;
}

// test hook (not in the public header): exact float32 scan of the whole bank with the arithmetic of the re-check kernels
// (warp_sqdist, lowest row on ties) -- what the certified pre-filter must reproduce bit for bit
int cmdb_debug_exact_min(cmdb_bank *b, const float *patch_host, int P, float *min_val_out, int64_t *min_idx_out) {
    CMDB_REQUIRE(b && patch_host && min_val_out && min_idx_out && P >= 1 && b->finalized, CMDB_ERR_INVALID, "cmdb_debug_exact_min: bad arguments");
    CMDB_CUDA(cudaSetDevice(b->device));
    float *q = nullptr;
    unsigned long long *keys = nullptr;
    std::vector<unsigned long long> h((size_t)P);
    cudaError_t e = cudaMalloc(&q, sizeof(float) * (size_t)P * b->dim);
    if (e == cudaSuccess) e = cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)P);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q, patch_host, sizeof(float) * (size_t)P * b->dim, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * (size_t)P, b->stream);
    int rc = CMDB_OK;
    if (e == cudaSuccess) rc = score_exact_scan(b, q, P, keys);
    if (e == cudaSuccess && rc == CMDB_OK) e = cudaMemcpyAsync(h.data(), keys, sizeof(unsigned long long) * (size_t)P, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess && rc == CMDB_OK) e = cudaStreamSynchronize(b->stream);
    cudaFree(q), cudaFree(keys);
    if (rc != CMDB_OK) return rc;
    CMDB_CUDA(e);
    for (int i = 0; i < P; ++i) {
        const unsigned int bits = (unsigned int)(h[i] >> 32);
        float d2;
        memcpy(&d2, &bits, sizeof(d2));
        min_val_out[i] = sqrtf(d2);
        min_idx_out[i] = (int64_t)(h[i] & 0xffffffffULL) + b->row_offset;
    }
    return CMDB_OK;
}

// host-only test hooks (not in the public header): the GEMM's tile-schedule stride and the fallback-tier rule
// CMDB_OPT_TIMING = 2: milliseconds since the time base of every stage mark of the last batch on each lane
// (out: float [2][CMDB_T_COUNT + 1 + 4]: the stage marks, then the four marks inside the refine stage); all submitted
// batches must have been waited for
int cmdb_debug_lane_timeline(cmdb_bank *b, float *out) {
    CMDB_REQUIRE(b && out && b->timing == 2 && b->ev_base, CMDB_ERR_STATE, "cmdb_debug_lane_timeline: set CMDB_OPT_TIMING = 2 first");
    CMDB_CUDA(cudaSetDevice(b->device));
    CMDB_CUDA(cudaDeviceSynchronize());
    constexpr int kN = CMDB_T_COUNT + 1 + 4;
    for (int l = 0; l < 2; ++l)
        for (int i = 0; i < kN; ++i) {
            cudaEvent_t e = i <= CMDB_T_COUNT ? b->ev_tl[l][i] : b->ev_dbg[l][i - CMDB_T_COUNT - 1];
            if (cudaEventElapsedTime(out + l * kN + i, b->ev_base, e) != cudaSuccess) {
                (void)cudaGetLastError();
                out[l * kN + i] = -1.f;
            }
        }
    return CMDB_OK;
}

int cmdb_debug_tile_stride(int mt, int G) { return score_tile_stride(mt, G); }
int cmdb_debug_fallback_use_rescan(int fails, int pairs) { return fallback_use_rescan(fails, pairs) ? 1 : 0; }

// test hook (not in the public header): rows [r0, r0 + n) of the neighbour table as packed (d^2 bits << 32 | row) keys
int cmdb_debug_read_knn(cmdb_bank *b, unsigned long long *out_host, long long r0, long long n) {
    CMDB_REQUIRE(b && out_host && b->knn_table && r0 >= 0 && n >= 1 && r0 + n <= b->fin_rows, CMDB_ERR_INVALID,
                 "cmdb_debug_read_knn: bad arguments");
    CMDB_CUDA(cudaSetDevice(b->device));
    CMDB_CUDA(cudaMemcpy(out_host, b->knn_table + (size_t)r0 * 3, sizeof(unsigned long long) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    return CMDB_OK;
}


// ---- late-fusion head on the device -------------------------------------------------------------------------------

static size_t fused_off_modal(int cap_b) { return (sizeof(double) * cap_b + 255) & ~size_t(255); }
static size_t fused_off_map(int cap_b) { return fused_off_modal(cap_b) + ((sizeof(float) * 3 * cap_b + 255) & ~size_t(255)); }
constexpr int kFusedCapB = 32;

static int fused_alloc(cmdb_bank *b0, int out_hw) {
    cmdb_bank::Fused &f = b0->fused;
    const size_t need = fused_off_map(kFusedCapB) + sizeof(double) * (size_t)out_hw * out_hw * kFusedCapB;
    if (f.cap_bytes >= need) return CMDB_OK;
    CMDB_REQUIRE(!f.active[0] && !f.active[1], CMDB_ERR_STATE, "fused scoring: out_hw grows while a submitted batch is outstanding");
    for (int i = 0; i < 2; ++i) {
        cudaFree(f.dev[i]);
        if (f.host[i]) cudaFreeHost(f.host[i]);
        f.dev[i] = f.host[i] = nullptr;
        CMDB_CUDA(cudaMalloc(&f.dev[i], need));
        CMDB_CUDA(cudaMallocHost(&f.host[i], need));
        if (!f.ev_done[i]) CMDB_CUDA(cudaEventCreateWithFlags(&f.ev_done[i], cudaEventDisableTiming));
    }
    f.cap_bytes = need;
    return CMDB_OK;
}

int cmdb_score_fused_batch_submit(cmdb_bank *const *banks, const float *const *patches, const int *P, const int *fh,
                                  const int *fw, int B, int out_hw, int patch_is_device, const cmdb_fusion_head *head,
                                  unsigned flags, int64_t *out_ticket) {
    CMDB_REQUIRE(banks && patches && P && fh && fw && head && out_ticket, CMDB_ERR_INVALID, "cmdb_score_fused_batch_submit: NULL argument");
    const int M = head->n_modal;
    CMDB_REQUIRE(M >= 1 && M <= 3, CMDB_ERR_INVALID, "cmdb_score_fused_batch_submit: n_modal=%d not in [1,3]", M);
    for (int m = 0; m < M; ++m) {
        CMDB_REQUIRE(banks[m], CMDB_ERR_INVALID, "cmdb_score_fused_batch_submit: banks[%d] is NULL", m);
        CMDB_REQUIRE(banks[m]->device == banks[0]->device, CMDB_ERR_INVALID, "cmdb_score_fused_batch_submit: all banks must live on one GPU");
        for (int k = 0; k < m; ++k)
            CMDB_REQUIRE(banks[k] != banks[m], CMDB_ERR_INVALID, "cmdb_score_fused_batch_submit: one handle per modality");
        CMDB_CHECK(check_batch_args(banks[m], patches[m], B, P[m], fh[m], fw[m], out_hw, "cmdb_score_fused_batch_submit"));
        CMDB_REQUIRE(B <= score_max_batch(banks[m]) && B <= kFusedCapB, CMDB_ERR_INVALID,
                     "cmdb_score_fused_batch_submit: batch=%d exceeds the per-call limit %d", B, std::min(kFusedCapB, score_max_batch(banks[m])));
        CMDB_REQUIRE(!banks[m]->pending[banks[m]->next_slot].active, CMDB_ERR_STATE,
                     "cmdb_score_fused_batch_submit: two batches are already outstanding; wait for one first");
    }
    cmdb_bank *b0 = banks[0];
    cmdb_bank::Fused &f = b0->fused;
    CMDB_CUDA(cudaSetDevice(b0->device));
    CMDB_CHECK(fused_alloc(b0, out_hw));
    const int fs = f.next_fs;   // two fused blocks: at most two fused batches are outstanding
    CMDB_REQUIRE(!f.active[fs], CMDB_ERR_STATE, "cmdb_score_fused_batch_submit: two fused batches are already outstanding; wait for one first");
    f.next_fs ^= 1;
    const int npix = out_hw * out_hw;
    const bool keep = (flags & CMDB_FUSED_KEEP_ON_DEVICE) != 0, host_maps = (flags & CMDB_FUSED_NO_HOST_MAPS) == 0;
    if (keep) {
        CMDB_REQUIRE(f.acc_maps && f.acc_npix == npix && f.acc_n + B <= f.acc_cap, CMDB_ERR_CAPACITY,
                     "cmdb_score_fused_batch_submit: device-side result store missing or full (%lld + %d > %lld); call cmdb_eval_reserve",
                     f.acc_n, B, f.acc_cap);
    }
    FuseParams fp{};
    fp.n_modal = M, fp.B = B, fp.npix = npix;
    for (int m = 0; m < M; ++m) {
        cmdb_bank *bm = banks[m];
        const SlotPick pick = take_slot(bm);
        const int slot = pick.rslot;
        CMDB_CHECK(submit_sub_batch(bm, patches[m], patch_is_device, B, P[m], fh[m], fw[m], out_hw, 0u, pick, false));
        bm->pending[slot].ticket = ++bm->ticket_counter;
        f.banks[fs][m] = bm, f.slots[fs][m] = slot;
        fp.maps[m] = reinterpret_cast<const float *>(bm->ss.out_block_buf[slot] + bm->ss.off_map_out);
        fp.map_stride[m] = bm->ss.map_stride;
        fp.tails[m] = reinterpret_cast<const TailResult *>(bm->ss.out_block_buf[slot]);
        fp.s_lambda[m] = head->s_lambda[m], fp.smap_lambda[m] = head->smap_lambda[m];
        fp.det_coef[m] = head->detect_coef[m], fp.seg_coef[m] = head->seg_coef[m];
        if (m > 0) CMDB_CUDA(cudaStreamWaitEvent(b0->stream, bm->ev_compute[slot], 0));
    }
    fp.det_off = head->detect_offset, fp.seg_off = head->seg_offset;
    unsigned char *blk = f.dev[fs];
    fp.out_s = reinterpret_cast<double *>(blk);
    fp.out_s_modal = reinterpret_cast<float *>(blk + fused_off_modal(kFusedCapB));
    fp.out_map = host_maps ? reinterpret_cast<double *>(blk + fused_off_map(kFusedCapB)) : nullptr;
    if (keep) {
        fp.acc_map = f.acc_maps + (size_t)f.acc_n * npix;
        fp.acc_s = f.acc_scores + f.acc_n;
        f.acc_n += B;
    }
    fuse_head_kernel<<<dim3((npix + 1023) / 1024, B), 256, 0, b0->stream>>>(fp);
    CMDB_CUDA(cudaGetLastError());
    // results of this slot: scalars (+ maps) on the d2h stream; the next batch's kernels may already run meanwhile
    CMDB_CUDA(cudaEventRecord(b0->ev_chunk[1], b0->stream));
    CMDB_CUDA(cudaStreamWaitEvent(b0->d2h_stream, b0->ev_chunk[1], 0));
    const size_t bytes = host_maps ? fused_off_map(kFusedCapB) + sizeof(double) * (size_t)npix * B : fused_off_map(kFusedCapB);
    CMDB_CUDA(cudaMemcpyAsync(f.host[fs], blk, bytes, cudaMemcpyDeviceToHost, b0->d2h_stream));
    CMDB_CUDA(cudaEventRecord(f.ev_done[fs], b0->d2h_stream));
    // the fuse kernel read the other banks' result blocks: their slots may only be reused after it (they wait on it through
    // the ticket: a slot is busy until cmdb_score_fused_batch_wait returned)
    f.active[fs] = true, f.n_modal[fs] = M, f.B[fs] = B, f.out_hw[fs] = out_hw;
    f.ticket[fs] = b0->pending[f.slots[fs][0]].ticket;
    *out_ticket = f.ticket[fs];
    return CMDB_OK;
}

int cmdb_score_fused_batch_wait(cmdb_bank *b0, int64_t ticket, cmdb_fused_out *outs) {
    CMDB_REQUIRE(b0 && outs, CMDB_ERR_INVALID, "cmdb_score_fused_batch_wait: NULL argument");
    cmdb_bank::Fused &f = b0->fused;
    for (int fs = 0; fs < 2; ++fs) {
        if (!f.active[fs] || f.ticket[fs] != ticket) continue;
        CMDB_CUDA(cudaSetDevice(b0->device));
        CMDB_CUDA(cudaEventSynchronize(f.ev_done[fs]));
        const int B = f.B[fs], M = f.n_modal[fs], npix = f.out_hw[fs] * f.out_hw[fs];
        std::vector<cmdb_score_out> tmp((size_t)B);
        for (int m = 0; m < M; ++m) {
            for (int i = 0; i < B; ++i) {
                tmp[i] = cmdb_score_out{};
                tmp[i].min_val = outs[i].min_val[m], tmp[i].min_idx = outs[i].min_idx[m];
            }
            CMDB_CHECK(wait_slot(f.banks[fs][m], f.slots[fs][m], tmp.data()));
        }
        const unsigned char *h = f.host[fs];
        const double *hs = reinterpret_cast<const double *>(h);
        const float *hm = reinterpret_cast<const float *>(h + fused_off_modal(kFusedCapB));
        const double *hmap = reinterpret_cast<const double *>(h + fused_off_map(kFusedCapB));
        for (int i = 0; i < B; ++i) {
            if (outs[i].s) *outs[i].s = hs[i];
            if (outs[i].s_modal)
                for (int m = 0; m < M; ++m) outs[i].s_modal[m] = hm[i * 3 + m];
            if (outs[i].s_map) memcpy(outs[i].s_map, hmap + (size_t)i * npix, sizeof(double) * npix);
        }
        f.active[fs] = false;
        return CMDB_OK;
    }
    set_error("cmdb_score_fused_batch_wait: ticket %lld is not outstanding", (long long)ticket);
    return CMDB_ERR_STATE;
}

int cmdb_score_fused_batch(cmdb_bank *const *banks, const float *const *patches, const int *P, const int *fh, const int *fw,
                           int B, int out_hw, int patch_is_device, const cmdb_fusion_head *head, unsigned flags,
                           cmdb_fused_out *outs) {
    CMDB_REQUIRE(banks && patches && P && head && outs && B >= 1, CMDB_ERR_INVALID, "cmdb_score_fused_batch: bad arguments");
    const int M = head->n_modal;
    CMDB_REQUIRE(M >= 1 && M <= 3, CMDB_ERR_INVALID, "cmdb_score_fused_batch: n_modal=%d not in [1,3]", M);
    int bc_max = kFusedCapB;
    for (int m = 0; m < M; ++m) {
        CMDB_REQUIRE(banks[m] && patches[m], CMDB_ERR_INVALID, "cmdb_score_fused_batch: NULL bank / patches");
        bc_max = std::min(bc_max, score_max_batch(banks[m]));
    }
    // sub-batches are pipelined: batch k + 1 is submitted before the host waits for batch k
    int64_t prev_ticket = 0;
    int prev_b0 = -1;
    for (int b0 = 0; b0 < B; b0 += bc_max) {
        const int bc = std::min(bc_max, B - b0);
        const float *pp[3] = {nullptr, nullptr, nullptr};
        for (int m = 0; m < M; ++m) pp[m] = patches[m] + (size_t)b0 * P[m] * banks[m]->dim;
        int64_t t = 0;
        CMDB_CHECK(cmdb_score_fused_batch_submit(banks, pp, P, fh, fw, bc, out_hw, patch_is_device, head, flags, &t));
        if (prev_b0 >= 0) CMDB_CHECK(cmdb_score_fused_batch_wait(banks[0], prev_ticket, outs + prev_b0));
        prev_ticket = t, prev_b0 = b0;
    }
    return cmdb_score_fused_batch_wait(banks[0], prev_ticket, outs + prev_b0);
}


// ---- device-side result store: replaces pixel_preds.extend / predictions.append of 50 176 scalars per image
//      (multiple_features.py:996-1001); the fused maps stay in HBM until cmdb_eval_* consumes them ----

int cmdb_eval_reserve(cmdb_bank *b, int64_t n_images, int out_hw) {
    CMDB_REQUIRE(b && n_images >= 1 && out_hw >= 8 && out_hw <= 256, CMDB_ERR_INVALID, "cmdb_eval_reserve: bad arguments");
    cmdb_bank::Fused &f = b->fused;
    CMDB_REQUIRE(!f.active[0] && !f.active[1], CMDB_ERR_STATE, "cmdb_eval_reserve: a submitted batch is outstanding");
    CMDB_CUDA(cudaSetDevice(b->device));
    CMDB_CUDA(cudaStreamSynchronize(b->stream));
    cudaFree(f.acc_maps), cudaFree(f.acc_scores);
    f.acc_maps = f.acc_scores = nullptr, f.acc_cap = f.acc_n = 0, f.acc_npix = out_hw * out_hw;
    CMDB_CUDA(cudaMalloc(&f.acc_maps, sizeof(double) * (size_t)n_images * f.acc_npix));
    CMDB_CUDA(cudaMalloc(&f.acc_scores, sizeof(double) * (size_t)n_images));
    f.acc_cap = n_images;
    return CMDB_OK;
}

int cmdb_eval_count(cmdb_bank *b, int64_t *out_n) {
    CMDB_REQUIRE(b && out_n, CMDB_ERR_INVALID, "cmdb_eval_count: bad arguments");
    *out_n = b->fused.acc_n;
    return CMDB_OK;
}

int cmdb_eval_reset(cmdb_bank *b) {
    CMDB_REQUIRE(b, CMDB_ERR_INVALID, "cmdb_eval_reset: bank is NULL");
    CMDB_REQUIRE(!b->fused.active[0] && !b->fused.active[1], CMDB_ERR_STATE, "cmdb_eval_reset: a submitted batch is outstanding");
    b->fused.acc_n = 0;
    return CMDB_OK;
}

int cmdb_eval_read(cmdb_bank *b, int64_t first, int64_t n, double *maps_host, double *scores_host) {
    cmdb_bank::Fused &f = b->fused;
    CMDB_REQUIRE(b && first >= 0 && n >= 0 && first + n <= f.acc_n, CMDB_ERR_INVALID, "cmdb_eval_read: images [%lld,%lld) outside [0,%lld)",
                 (long long)first, (long long)(first + n), (long long)f.acc_n);
    if (n == 0) return CMDB_OK;
    CMDB_CUDA(cudaSetDevice(b->device));
    if (maps_host)
        CMDB_CUDA(cudaMemcpyAsync(maps_host, f.acc_maps + (size_t)first * f.acc_npix, sizeof(double) * (size_t)n * f.acc_npix,
                                  cudaMemcpyDeviceToHost, b->stream));
    if (scores_host)
        CMDB_CUDA(cudaMemcpyAsync(scores_host, f.acc_scores + first, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, b->stream));
    CMDB_CUDA(cudaStreamSynchronize(b->stream));
    return CMDB_OK;
}

int cmdb_upsample_blur(int device, const float *map_host, int fh, int fw, int out_hw, float *out_host, float *out_pre_host,
                       uint8_t *out_u8_host) {
    CMDB_REQUIRE(map_host && out_host && fh > 0 && fw > 0, CMDB_ERR_INVALID, "cmdb_upsample_blur: bad arguments");
    CMDB_REQUIRE(out_hw >= 8 && out_hw <= 256, CMDB_ERR_INVALID, "cmdb_upsample_blur: out_hw=%d not in [8,256]", out_hw);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        set_error("cmdb_upsample_blur: no CUDA device; this library has no CPU fallback");
        return CMDB_ERR_CUDA;
    }
    CMDB_CUDA(cudaSetDevice(device));
    const size_t npix = (size_t)out_hw * out_hw;
    float *in = nullptr, *pre = nullptr, *o = nullptr, *mx = nullptr;
    unsigned char *u8 = nullptr, *tmp = nullptr;
    cudaError_t e = cudaMalloc(&in, sizeof(float) * fh * fw);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, npix);
    if (e == cudaSuccess) e = cudaMalloc(&mx, sizeof(float) * 17);
    if (e == cudaSuccess) e = cudaMalloc(&pre, sizeof(float) * npix);
    if (e == cudaSuccess) e = cudaMalloc(&o, sizeof(float) * npix);
    if (e == cudaSuccess) e = cudaMalloc(&u8, npix);
    if (e == cudaSuccess) e = cudaMemcpy(in, map_host, sizeof(float) * fh * fw, cudaMemcpyHostToDevice);
    int rc = CMDB_OK;
    if (e == cudaSuccess) rc = upsample_blur_launch(nullptr, 1, 0, 1, npix, in, fh, fw, out_hw, pre, o, u8, tmp, mx, mx + 1);
    if (e == cudaSuccess && rc == CMDB_OK) e = cudaMemcpy(out_host, o, sizeof(float) * npix, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == CMDB_OK && out_pre_host) e = cudaMemcpy(out_pre_host, pre, sizeof(float) * npix, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == CMDB_OK && out_u8_host) e = cudaMemcpy(out_u8_host, u8, npix, cudaMemcpyDeviceToHost);
    cudaFree(in), cudaFree(pre), cudaFree(o), cudaFree(u8), cudaFree(tmp), cudaFree(mx);
    if (rc != CMDB_OK) return rc;
    CMDB_CUDA(e);
    return CMDB_OK;
}

}  // extern "C"
