// extern "C" entry points that orchestrate the kernels: coreset selection, projection, scoring (single GPU and the
// row-sharded phases), stand-alone upsample+blur.  Declarations and reference citations: include/cmdiad_b200.h.
#include <math.h>

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

// host block of a slot -> the caller's buffers
static void scatter_outputs(const cmdb_bank *b, const unsigned char *h, int B, int P, int out_hw, cmdb_score_out *outs,
                            int img_first = 0, int img_step = 1) {
    const ScoreScratch &s = b->ss;
    const size_t npix = (size_t)out_hw * out_hw;
    for (int i = img_first; i < B; i += img_step) {
        cmdb_score_out *out = outs + i;
        TailResult tr;
        memcpy(&tr, h + sizeof(TailResult) * i, sizeof(tr));
        if (out->min_val) memcpy(out->min_val, h + s.off_min_val + sizeof(float) * (size_t)i * P, sizeof(float) * P);
        if (out->min_idx) memcpy(out->min_idx, h + s.off_min_idx + sizeof(long long) * (size_t)i * P, sizeof(long long) * P);
        if (out->s_map) memcpy(out->s_map, h + s.off_map_out + sizeof(float) * s.map_stride * i, sizeof(float) * npix);
        if (out->s_map_pre) memcpy(out->s_map_pre, h + s.off_map_pre + sizeof(float) * s.map_stride * i, sizeof(float) * npix);
        if (out->s_map_u8) memcpy(out->s_map_u8, h + s.off_map_u8 + s.map_stride * i, npix);
        if (out->s) *out->s = tr.s;
        if (out->s_star) *out->s_star = tr.s_star;
        if (out->s_idx) *out->s_idx = tr.s_idx;
        if (out->w) *out->w = tr.w;
        if (out->m_star_knn) out->m_star_knn[0] = tr.knn0, out->m_star_knn[1] = tr.knn1;
        if (out->nn_idx)
            for (int k = 0; k < 3; ++k) out->nn_idx[k] = tr.nn_idx[k];
    }
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
                                s.map_out, s.map_u8, s.map_tmp, s.map_max);
}

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

// 16-byte key header + one 8-byte flagged word per two halves of the row
static unsigned int mailbox_slot_stride(int d) { return 16u + 8u * (unsigned int)((d + 1) / 2 + 1); }

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
    return (size_t)2 * (size_t)(world > 0 ? world : 1) * mailbox_slot_stride(d_proj_max > 0 ? d_proj_max : 1);
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
    sh.slot_stride = mailbox_slot_stride(d_proj);
    CMDB_REQUIRE(mb_bytes >= cmdb_coreset_mailbox_bytes(sh.world, d_proj), CMDB_ERR_INVALID,
                 "cmdb_coreset_select_sharded: mailbox has %zu bytes, need %zu", mb_bytes,
                 cmdb_coreset_mailbox_bytes(sh.world, d_proj));
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
static int submit_sub_batch(cmdb_bank *b, const float *src, int is_device, int bc, int P, int fh, int fw, int out_hw,
                            unsigned want, int slot) {
#define CMDB_MARK(i)                                                   \
    do {                                                               \
        if (b->timing) CMDB_CUDA(cudaEventRecord(b->ev[i], b->stream)); \
    } while (0)
    ScoreScratch &s = b->ss;
    cudaStream_t st = b->stream;
    CMDB_CHECK(stage_alloc(b, bc, P, out_hw));
    score_select_slot(b, slot);
    CMDB_MARK(CMDB_T_STAGE_IN);
    // this slot's q_f32 was last read by the batch submitted two calls ago
    CMDB_CHECK(score_local_min(b, src, is_device, bc, P, CMDB_T_GEMM, CMDB_T_REFINE, b->ev_compute[slot]));
    CMDB_MARK(CMDB_T_MAP);
    CMDB_CHECK(blur_batch(b, bc, fh, fw, out_hw));
    // min_val / min_idx / maps do not depend on the re-weighting: their copy overlaps it
    CMDB_CUDA(cudaEventRecord(b->ev_chunk[0], st));
    CMDB_CUDA(cudaStreamWaitEvent(b->d2h_stream, b->ev_chunk[0], 0));
    CMDB_CUDA(cudaMemcpyAsync(s.out_block_host + s.off_min_val, s.out_block + s.off_min_val,
                              out_block_extent(b, bc, want) - s.off_min_val, cudaMemcpyDeviceToHost, b->d2h_stream));
    CMDB_MARK(CMDB_T_REWEIGHT);
    CMDB_CHECK(score_reweight(b, bc, P, true));
    CMDB_MARK(CMDB_T_OUT);
    CMDB_CUDA(cudaEventRecord(b->ev_compute[slot], st));
    CMDB_CUDA(cudaStreamWaitEvent(b->d2h_stream, b->ev_compute[slot], 0));
    CMDB_CUDA(cudaMemcpyAsync(s.out_block_host, s.out_block, s.off_min_val, cudaMemcpyDeviceToHost, b->d2h_stream));
    CMDB_CUDA(cudaEventRecord(b->ev_done[slot], b->d2h_stream));
    if (b->timing) CMDB_CUDA(cudaEventRecord(b->ev[CMDB_T_COUNT], b->d2h_stream));
#undef CMDB_MARK
    cmdb_bank::Pending &pd = b->pending[slot];
    pd.active = true, pd.B = bc, pd.P = P, pd.out_hw = out_hw, pd.want = want;
    return CMDB_OK;
}

static int wait_slot(cmdb_bank *b, int slot, cmdb_score_out *outs) {
    cmdb_bank::Pending &pd = b->pending[slot];
    CMDB_CUDA(cudaEventSynchronize(b->ev_done[slot]));
    pd.active = false;
    scatter_outputs(b, b->ss.out_block_host_buf[slot], pd.B, pd.P, pd.out_hw, outs);
    if (b->timing) b->ev_valid = true;
    return CMDB_OK;
}

static int check_batch_args(cmdb_bank *b, const float *patches, int B, int P, int fh, int fw, int out_hw, const char *fn) {
    CMDB_CHECK(check_score_args(b, patches, B, P, fn));
    CMDB_REQUIRE(fh > 0 && fw > 0 && fh * fw == P, CMDB_ERR_INVALID, "%s: feature_map_dims %dx%d != P=%d", fn, fh, fw, P);
    CMDB_REQUIRE(out_hw >= 8 && out_hw <= 256, CMDB_ERR_INVALID, "%s: out_hw=%d not in [8,256]", fn, out_hw);
    CMDB_CUDA(cudaSetDevice(b->device));
    if (b->ss.map_stride && (size_t)out_hw * out_hw != b->ss.map_stride) {  // map stride is fixed per scratch
        CMDB_REQUIRE(!b->pending[0].active && !b->pending[1].active, CMDB_ERR_STATE,
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
    for (int b0 = 0; b0 < B; b0 += bc_max) {
        const int bc = std::min(bc_max, B - b0);
        const int slot = b->next_slot;
        CMDB_REQUIRE(!b->pending[slot].active, CMDB_ERR_STATE, "cmdb_score_batch: two submitted batches are outstanding; wait for one first");
        b->next_slot ^= 1;
        CMDB_CHECK(submit_sub_batch(b, patches + (size_t)b0 * P * b->dim, patch_is_device, bc, P, fh, fw, out_hw,
                                    want_mask_of(outs + b0, bc), slot));
        CMDB_CHECK(wait_slot(b, slot, outs + b0));
    }
    return CMDB_OK;
}

int cmdb_score_batch_submit(cmdb_bank *b, const float *patches, int B, int P, int fh, int fw, int out_hw, int patch_is_device,
                            unsigned want_maps, int64_t *out_ticket) {
    CMDB_REQUIRE(out_ticket, CMDB_ERR_INVALID, "cmdb_score_batch_submit: out_ticket is NULL");
    CMDB_CHECK(check_batch_args(b, patches, B, P, fh, fw, out_hw, "cmdb_score_batch_submit"));
    CMDB_REQUIRE(B <= score_max_batch(b), CMDB_ERR_INVALID, "cmdb_score_batch_submit: batch=%d exceeds the per-call limit %d", B,
                 score_max_batch(b));
    const int slot = b->next_slot;
    CMDB_REQUIRE(!b->pending[slot].active, CMDB_ERR_STATE,
                 "cmdb_score_batch_submit: two batches are already outstanding; call cmdb_score_batch_wait first");
    CMDB_CHECK(submit_sub_batch(b, patches, patch_is_device, B, P, fh, fw, out_hw, want_maps & 3u, slot));
    b->next_slot ^= 1;
    b->pending[slot].ticket = ++b->ticket_counter;
    *out_ticket = b->pending[slot].ticket;
    return CMDB_OK;
}

int cmdb_score_batch_wait(cmdb_bank *b, int64_t ticket, cmdb_score_out *outs) {
    CMDB_REQUIRE(b && outs, CMDB_ERR_INVALID, "cmdb_score_batch_wait: NULL argument");
    for (int slot = 0; slot < 2; ++slot)
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
    CMDB_REQUIRE(!b->pending[0].active && !b->pending[1].active, CMDB_ERR_STATE,
                 "cmdb_score_shard_min: a submitted batch is outstanding on this handle; wait for it first");
    if ((size_t)out_hw * out_hw != b->ss.map_stride) score_scratch_free(b);
    CMDB_CHECK(stage_alloc(b, B, P, out_hw));
    score_select_slot(b, 0);
    CMDB_CHECK(score_local_min(b, patches, patch_is_device, B, P, -1, -1));
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

// host-only test hooks (not in the public header): the GEMM's tile-schedule stride and the fallback-tier rule
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
    if (e == cudaSuccess) e = cudaMalloc(&mx, sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&pre, sizeof(float) * npix);
    if (e == cudaSuccess) e = cudaMalloc(&o, sizeof(float) * npix);
    if (e == cudaSuccess) e = cudaMalloc(&u8, npix);
    if (e == cudaSuccess) e = cudaMemcpy(in, map_host, sizeof(float) * fh * fw, cudaMemcpyHostToDevice);
    int rc = CMDB_OK;
    if (e == cudaSuccess) rc = upsample_blur_launch(nullptr, 1, 0, 1, npix, in, fh, fw, out_hw, pre, o, u8, tmp, mx);
    if (e == cudaSuccess && rc == CMDB_OK) e = cudaMemcpy(out_host, o, sizeof(float) * npix, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == CMDB_OK && out_pre_host) e = cudaMemcpy(out_pre_host, pre, sizeof(float) * npix, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == CMDB_OK && out_u8_host) e = cudaMemcpy(out_u8_host, u8, npix, cudaMemcpyDeviceToHost);
    cudaFree(in), cudaFree(pre), cudaFree(o), cudaFree(u8), cudaFree(tmp), cudaFree(mx);
    if (rc != CMDB_OK) return rc;
    CMDB_CUDA(e);
    return CMDB_OK;
}

}  // extern "C"
