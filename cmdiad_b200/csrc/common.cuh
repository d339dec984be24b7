// Shared declarations of the cmdiad_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/cmdiad_b200.h"

namespace cmdb {

void set_error(const char *fmt, ...);

#define CMDB_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            cmdb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
            (void)cudaGetLastError();                                                                \
            return CMDB_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)

#define CMDB_CHECK(expr)           \
    do {                           \
        int _s = (expr);           \
        if (_s != CMDB_OK) return _s; \
    } while (0)

#define CMDB_REQUIRE(cond, status, ...)   \
    do {                                  \
        if (!(cond)) {                    \
            cmdb::set_error(__VA_ARGS__); \
            return status;                \
        }                                 \
    } while (0)

constexpr int kNumSMsDefault = 148;
constexpr int kMaxRanks = 8;         // GPUs of one NVSwitch box
constexpr int kScoreBN = 256;        // bank rows per GEMM tile
constexpr int kScoreBM = 128;        // query rows per GEMM tile
constexpr int kResultSlots = 3;      // result blocks / outstanding submitted calls per handle (the compute lanes stay two)
constexpr int kMaxStageChunks = 4;   // host query batches are staged and multiplied in up to this many chunks
constexpr int kWorkCap = 16384;      // capacity of ScoreScratch::work_list
// Fallback tiers of the certified pre-filter.  Tier 1, exact rescan: ~R/296 rows x D x 4 B per (query, producer) pair, HBM
// bound (~0.32 us per pair at 200k x 768) -- RIGOROUS: the result equals the exact float32 scan, lowest row on ties.
// Tier 2, 3-term GEMM over the compacted uncertified queries + exact re-check of the 4 best candidates: at least one sweep
// of the hi + lo bank (~0.25 ms), then ~0.55 us per query -- cheaper when a call queues very many pairs, but NOT certified:
// among more than 4 rows within the float32 resolution of |a|^2 + |b|^2 - 2ab it may return another one of them (not
// necessarily the exact minimum / lowest row).  Round 1 picked the cheaper tier; round 2 keeps the contract instead: tier 1 whenever the
// pairs fit the work list (<= 5 ms of rescans in the worst case), tier 2 only beyond it (banks dominated by rows that
// float32 cannot tell apart, where the reference's own mm-form argmin is arbitrary as well).
__host__ __device__ inline bool fallback_use_rescan(int fails, int pairs) {
    (void)fails;
    return pairs <= kWorkCap;
}
constexpr int kScoreBK = 64;         // fp16 K elements per pipeline stage (one 128-byte swizzle row)

// device scratch of one scoring call, sized at finalize time
struct ScoreScratch {
    int cap_p = 0;                  // padded query-row capacity of a sub-batch (multiple of 128)
    int cap_b = 0;                  // image capacity of a sub-batch
    float *q_f32 = nullptr;         // [cap_p, D]  normalised query patches (the slot selected by score_select_slot)
    float *q_f32_buf[kResultSlots] = {};   // per result slot: batches k+1 and k+2 are staged while batch k is still scored
    __half *q_hi = nullptr;         // [cap_p, D]  split-fp16 query operand
    __half *q_lo = nullptr;
    int *q_scale_exp = nullptr;     // device [cap_p]: per-row exponent e_q with q_hi + q_lo = q * 2^e_q
    float *q_norm = nullptr;        // [cap_p] ||q|| (rounded up)            -- certified pre-filter, score_tail.cu
    float *q_eps = nullptr;         // [cap_p] ||q - q_hi * 2^-e_q|| (rounded up)
    int *fail_list = nullptr;       // [cap_p] query rows whose pre-filter result could not be certified
    // device control block of the fallback: [0] uncertified queries, [1] (query, producer) pairs to rescan,
    // [2] rows of the 3-term GEMM fallback, [3] pairs of the exact rescan, [4] queries the rescan finishes
    // ([2..4] are derived from [0..1] by the last block of refine_cert_kernel: pairs that fit the work list -> rescan, more ->
    // GEMM), [5] / [6] last-block tickets of the certificate / rescan kernels; the per-image argmax keys (s_key) follow at
    // fail_ctl + 8 in the same allocation so that one memset clears both
    int *fail_ctl = nullptr;
    int *fail_count_host = nullptr; // pinned copy of [0..1] (statistics / adaptive mode)
    bool sched_pair = false;            // the last first-pass GEMM ran on CTA pairs (cta_group::2)
    bool sched_pair_last = false;       // ... and the most recent GEMM launch of any kind
    int sched_run_last = 1;             // N tiles per visit (score_gemm_run) of the most recent GEMM launch
    int chunk_tiles = 0, mt_total = 0;  // M-tile layout of the last first-pass GEMM: chunks of chunk_tiles tiles (rescan needs it)
    int2 *work_list = nullptr;      // [kWorkCap] (query row, producer) pairs whose producer may hide rows inside the band
    unsigned long long *best_key = nullptr;  // [cap_p] running exact (d^2 bits << 32 | row) of the uncertified queries
    void *tmap_qhi = nullptr;       // host CUtensorMap objects for q_hi / q_lo
    void *tmap_qlo = nullptr;
    float *m_test = nullptr;        // [D] patch[s_idx]
    float *m_star = nullptr;        // [D] bank[min_idx[s_idx]]
    float *nn_rows = nullptr;       // [3, D] rows of the 3 nearest neighbours (sharded mode)
    unsigned long long *top3 = nullptr;  // merged 3 smallest w_dist keys
    int n_topk_blocks = 0;
    unsigned int *done_counter = nullptr;  // last-block-done counter of reweight_kernel
    float4 *cand = nullptr;         // [n_ctas, cap_p] per-CTA running top-2 (val1, idx1, val2, idx2) of the GEMM epilogue
    size_t cand_window_bytes = 0;   // L2 access-policy window over cand (0 = persistence not available / disabled)
    float cand_hit_ratio = 0.f;
    float *min_val = nullptr;       // [cap_p]
    long long *min_idx = nullptr;   // [cap_p]
    unsigned long long *s_key = nullptr;    // packed argmax key of min_val
    unsigned long long *topk_keys = nullptr;  // [n_blocks*3] per-block w_dist top-3 packed keys
    void *tail = nullptr;           // TailResult
    // results live in ONE device block (tail | min_val | min_idx | map_out | map_pre | map_u8) mirrored by a pinned
    // host block, so a scoring call ends with a single device->host copy
    unsigned char *out_block = nullptr;       // (current slot)
    unsigned char *out_block_host = nullptr;  // cudaMallocHost
    // per RESULT SLOT (kResultSlots of them, cycled independently of the two lanes): the results of batches k - 1 and k
    // travel to the host / wait for the caller while batch k + 1 is already enqueued behind batch k - 1 on its lane
    unsigned char *out_block_buf[kResultSlots] = {};
    unsigned char *out_block_host_buf[kResultSlots] = {};
    size_t off_min_val = 0, off_min_idx = 0, off_map_out = 0, off_map_pre = 0, off_map_u8 = 0, out_block_bytes = 0;
    size_t map_stride = 0;  // pixels reserved per image in the map sections
    float *map_pre = nullptr;       // [out_hw^2]
    float *map_out = nullptr;
    unsigned char *map_u8 = nullptr;
    unsigned char *map_tmp = nullptr;  // horizontally blurred 8-bit image between the two blur kernels
    float *map_max = nullptr;
    int map_cap = 0;
};

}  // namespace cmdb

struct cmdb_bank {
    int device = 0;
    int dim = 0;
    int num_sms = cmdb::kNumSMsDefault;
    int64_t capacity = 0;
    int64_t rows = 0;
    int64_t row_offset = 0;
    int score_impl = CMDB_SCORE_TCGEN05;
    // distance GEMM mode: 0 = certified hi.hi pre-filter with FP32-equivalent fallback for uncertified queries (default),
    // 3 = FP32-equivalent split for every query, 1 = uncertified hi.hi pre-filter (diagnostics)
    int prefilter_terms = 0;
    // error-bound inputs of the certificate, true units, rounded up (cmdb_bank_finalize)
    float cert_bmax = 0.f;    // max ||b||
    float cert_eb_max = 0.f;  // max ||b - b_hi * 2^-scale_exp||
    unsigned int *cert_buf = nullptr;  // device [2]: the two maxima as float bits
    // optional table of the three nearest bank rows of every bank row, as packed (d^2 bits << 32 | row) keys
    // (cmdb_bank_build_knn): turns the per-image w_dist pass into a lookup
    unsigned long long *knn_table = nullptr;  // [knn_rows][3]
    // rows the table covers: fin_rows for a table built on this handle (cmdb_bank_build_knn); the GLOBAL row count for a
    // replicated table installed on a row-sharded handle (cmdb_bank_set_knn_table)
    long long knn_rows = 0;
    // statistics of the last scoring call / adaptive fallback to the direct 3-term GEMM
    int64_t last_queries = 0;
    int last_mode = 0;           // GEMM mode the last call actually ran
    bool fail_pending = false;   // fail_count_host holds the count of a finished-or-in-flight call
    int direct_calls_left = 0;   // certified mode: calls to run directly with 3 terms before probing the pre-filter again
    // Two compute LANES: scoring call k runs on lane k & 1 (its own stream and its own copy of every scratch buffer, ss_store),
    // so the tail of a batch / round -- certificate, rescans, exchanges, re-weighting, maps -- overlaps the distance GEMM of
    // the next one (the GEMM is one persistent CTA per SM; the small tail kernels fit beside it).  `stream` and `ss` are the
    // lane currently selected (score_select_slot); everything that is not scoring runs on whichever lane is current.
    cudaStream_t stream = nullptr;
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    // per lane: side stream for the (normally empty) tier-2 launches and the counters copy, which run beside the rescan
    cudaStream_t lane_aux[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork[2] = {}, ev_join[2] = {};
    cmdb::ScoreScratch ss_store[2];
    int *last_fail_host = nullptr;   // pinned certificate counters of the most recent certified call (either lane)
    cudaStream_t copy_stream = nullptr;   // host -> device staging of query chunks, overlapped with the GEMM of earlier chunks
    cudaEvent_t ev_chunk[cmdb::kMaxStageChunks] = {};
    cudaStream_t d2h_stream = nullptr;    // device -> host copies of the results
    cudaEvent_t ev_done[cmdb::kResultSlots] = {};   // result slot's block is in the pinned host memory
    cudaEvent_t ev_compute[cmdb::kResultSlots] = {};   // the kernels of the call that used the slot are done (its q_f32 may be overwritten)
    cudaEvent_t ev_fail = nullptr;        // the certificate counters of the last certified call are on the host
    cudaEvent_t ev_stage = nullptr;       // cmdb_bank_stage_h2d: the staged bytes are on the device
    struct Pending {
        bool active = false;
        int B = 0, P = 0, out_hw = 0;
        unsigned want = 0;
        bool host_maps = true;   // false: the per-modality maps stay in HBM (fused late-fusion head), only scalars travel
        int img_first = 0, img_step = 1;  // sharded rounds: the images whose maps this rank finished
        long long ticket = 0;
    } pending[cmdb::kResultSlots];
    bool any_pending() const {
        for (const auto &p : pending)
            if (p.active) return true;
        return false;
    }
    // queries are normalised on the device right after staging: (q - q_mean) / q_std, one IEEE subtract + one IEEE divide
    // like the reference's torch expression (multiple_features.py:90, 976-977); cmdb_bank_set_query_norm
    bool q_norm_enabled = false;
    float q_mean = 0.f, q_std = 1.f;
    // late-fusion head on the device (cmdb_score_fused_batch*; this handle is modality 0 of the fused call): per slot the
    // fused float64 maps / image scores / lambda-scaled per-modality s in one device block mirrored by a pinned host block
    struct Fused {
        unsigned char *dev[2] = {nullptr, nullptr};
        unsigned char *host[2] = {nullptr, nullptr};
        size_t cap_bytes = 0;
        cudaEvent_t ev_done[2] = {};
        bool active[2] = {false, false};
        int next_fs = 0;   // fused block of the next fused call (two fused batches may be outstanding)
        int n_modal[2] = {0, 0}, B[2] = {0, 0}, out_hw[2] = {0, 0};
        cmdb_bank *banks[2][3] = {};
        int slots[2][3] = {};
        long long ticket[2] = {0, 0};
        // device-resident accumulation of the fused maps over predict calls (cmdb_eval_*)
        double *acc_maps = nullptr;     // [acc_cap][npix]
        double *acc_scores = nullptr;   // [acc_cap]
        long long acc_cap = 0, acc_n = 0;
        int acc_npix = 0;
    } fused;
    long long ticket_counter = 0;
    int next_slot = 0;    // result slot of the next submitted call (cycles through kResultSlots)
    int next_lane = 0;    // compute lane of the next submitted call (alternates)
    int shard_lane = 0;   // lane of the sharded round in progress
    int shard_slot = 0;   // result slot of the sharded round in progress (cmdb_score_shard_min .. finish)
    cmdb_comm *comm = nullptr;             // peer buffers for NCCL-free sharded rounds (cmdb_bank_attach_comm; not owned)
    unsigned int *shard_ctr = nullptr;     // [0] device: last-block counter of the push kernels
    unsigned int *shard_abort_host = nullptr;  // pinned + mapped: set by a kernel whose peers never arrived
    unsigned int *shard_abort_dev = nullptr;   // device alias of shard_abort_host
    float *shard_d2 = nullptr;             // device [2][kShardD2Cap]: local contributions / sums of the neighbour distances
    float *data = nullptr;  // [capacity, dim] float32 row-major
    // scoring layout (cmdb_bank_finalize)
    bool finalized = false;
    int64_t fin_rows = 0;
    int64_t fin_rows_pad = 0;  // multiple of kScoreBN
    __half *hi = nullptr;      // [fin_rows_pad, dim] fp16(x * 2^scale_exp)
    __half *lo = nullptr;      // [fin_rows_pad, dim] fp16(x * 2^scale_exp - hi)
    float *norm = nullptr;     // [fin_rows_pad] ||x * 2^scale_exp||^2 ; +inf on padding rows
    int scale_exp = 0;
    void *tmap_hi = nullptr;  // host copies of the CUtensorMap objects (128 B each)
    void *tmap_lo = nullptr;
    void *tmap_hi2 = nullptr;  // box of 128 bank rows (CTA-pair kernels)
    void *tmap_lo2 = nullptr;
    cmdb::ScoreScratch ss;
    int timing = 0;
    cudaEvent_t ev[CMDB_T_COUNT + 1] = {};
    bool ev_valid = false;
    // CMDB_OPT_TIMING = 2 (diagnostics): per-lane copies of the stage events and a common time base, so that the
    // overlap of the two lanes can be read (cmdb_debug_lane_timeline)
    cudaEvent_t ev_tl[2][CMDB_T_COUNT + 1] = {};
    cudaEvent_t ev_base = nullptr;
    cudaEvent_t ev_dbg[2][4] = {};   // inside the refine stage: before / after the certificate kernel, after the rescan, after the tier-2 launches
    int cur_slot = 0;
    double *stats_buf = nullptr;  // 2 doubles on device
    unsigned int *absmax_buf = nullptr;
};

namespace cmdb {

// bank.cu
int bank_max_abs(cmdb_bank *b, const float *x, int64_t n, float *out_host);
int pick_scale_exp(float absmax);
void launch_normalize(cudaStream_t stream, int num_sms, float *x, int64_t n, float mean, float stdv);
void launch_split_rows(cudaStream_t stream, int num_sms, const float *x, int64_t n_rows, int64_t n_pad, int dim,
                       int scale_exp, __half *hi, __half *lo, float *norm, float pad_norm, unsigned int *cert_buf);

// project.cu
int project_rows(cmdb_bank *b, const float *x_dev, int64_t n_rows, int D, const int32_t *indptr_h,
                 const int32_t *indices_h, const double *data_h, int d_proj, double *z_dev);

// comm.cu
int comm_info(cmdb_comm *c, int *rank, int *world, unsigned char **local, unsigned char **peers, size_t *bytes);
unsigned long long comm_next_score_epoch(cmdb_comm *c);

// coreset.cu
// layout of a rank's peer-mapped buffer (cmdb_comm): [0, 256) per-rank "shard arrived" flags | key slots
// [parity][rank][CTA] x 32 B | at kCommHeaderBytes: replica of the whole projected bank in storage type
constexpr unsigned int kCommReadyOff = 0;
constexpr unsigned int kCommKeysOff = 256;
constexpr unsigned int kCommMaxCtas = 192;
constexpr size_t kCommCoresetBytes = 256 * 1024;  // coreset flags + key slots (>= 256 + 2 * kMaxRanks * kCommMaxCtas * 32); cleared by cmdb_comm_reset
// row-sharded SCORING rounds exchange through the same buffer (never cleared: every word is rewritten with the round's epoch):
//   flags  [4 slots][2 kinds][kMaxRanks] u64 at kCommScoreFlagsOff
//   d2     [4 slots][kMaxRanks][kShardD2Cap] float at kCommScoreD2Off        (squared neighbour distances, 2 per image)
//   keys   [4 slots][kMaxRanks][kShardKeysCap] int64 at kCommScoreKeysOff    (packed (min distance, global row) per query)
// (4 slots, round r uses slot r & 3: with two lanes per rank a rank may start round r + 2 while a peer still reads round r)
constexpr size_t kCommScoreFlagsOff = kCommCoresetBytes;
constexpr size_t kCommScoreD2Off = kCommScoreFlagsOff + 4096;
constexpr int kShardD2Cap = 64;
constexpr size_t kCommScoreKeysOff = kCommScoreD2Off + 16384;
constexpr int kShardKeysCap = 128 * 1024;   // queries per round (32 images x 3136 patches = 100 352)
constexpr int kShardSlots = 4;
constexpr size_t kCommHeaderBytes = 34 * 1024 * 1024;  // >= kCommScoreKeysOff + kShardSlots * kMaxRanks * kShardKeysCap * 8; the replica follows
struct ShardCtx {  // row-sharded coreset loop
    int world, rank;
    long long row_offset, n_total;
    unsigned char *peers[kMaxRanks];
    size_t comm_bytes;
    const double *z0_host;  // float64 projection of global row 0
};
int coreset_greedy_dev(cmdb_bank *b, const double *z_dev, int64_t N, int d, int64_t n_select, int dtype_mode,
                       int64_t *out_idx_host, const int64_t *force_idx_host, void *out_min_last_host, const ShardCtx *sh);
int coreset_greedy(cmdb_bank *b, const double *z_dev, int64_t N, int d, int64_t n_select, int dtype_mode,
                   int64_t *out_idx_host);
int coreset_rownorms(int device, const void *z_host, const void *last_host, int64_t n_rows, int d, int dtype_mode,
                     void *out_host);

// score_gemm.cu
int score_scratch_alloc(cmdb_bank *b, int B, int P_img, int out_hw);
void score_select_slot(cmdb_bank *b, int lane, int rslot);  // lane's stream / scratch / q_f32, result pointers at block `rslot`
int score_max_batch(const cmdb_bank *b);  // images per internal sub-batch (shared-memory bound of reweight_kernel)
void score_scratch_free(cmdb_bank *b);
int score_make_tensor_maps(cmdb_bank *b);
// q_f32 -> q_hi / q_lo (device-side scale selection).  compact = true: only the rows of ss.fail_list (count on the
// device), written to rows 0..count-1
int score_query_prep(cmdb_bank *b, int P, bool compact, int row0 = 0);  // rows [row0, row0 + P) (row0 % 128 == 0)
// q_hi / q_lo -> cand via the tcgen05 distance GEMM with `terms` MMAs per K step; compact = true: the M extent is the
// device-side fail count
int score_gemm_candidates(cmdb_bank *b, int P, int terms, bool compact, int *n_cand_out, int row0 = 0);
void q_split_rows(cmdb_bank *b, const float *rows_dev, int n);
int score_gemm_groups();              // epilogue warp groups per CTA: producers = groups * CTAs
int score_tile_stride(int mt, int G); // host copy of the GEMM's tile schedule stride
int score_gemm_run(int nt, int mt_units, int G);  // N tiles per visit of an M tile for a launch of that shape
// stage the queries (src: host or device, [B*P_img, dim]) + candidates + refine, all modes
// stage_after: event the copy stream waits for before it overwrites q_f32 (nullptr: everything queued so far)
int score_local_min(cmdb_bank *b, const float *src, int src_is_device, int B, int P_img, int ev_gemm, int ev_refine,
                    cudaEvent_t stage_after = nullptr);

// The distance GEMM keeps one persistent CTA per SM with ~200 KB of dynamic shared memory, i.e. the SM runs in its largest
// shared-memory carve-out.  For the other lane's small kernels to become co-resident with it they must ask for the SAME
// carve-out (an SM is only re-configured when it is idle); one function per translation unit sets that preference once.
#define CMDB_PREFER_MAX_SMEM(kernel) \
    (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
void tail_prefer_carveout();   // score_tail.cu
void gemm_prefer_carveout();   // score_gemm.cu
void api_prefer_carveout();    // api.cu
void bank_prefer_carveout();   // bank.cu

// score_tail.cu
struct TailResult {  // device-side result block (ScoreScratch::tail)
    float s, s_star, w, knn0, knn1;
    int s_idx;
    long long nn_idx[3];
    long long m_star_row;  // global row of m_star
};
int score_simt_candidates(cmdb_bank *b, int P, int *n_cand_out);
int score_refine(cmdb_bank *b, int B, int P_img, int n_cand, bool compact);
int score_refine_certified(cmdb_bank *b, int B, int P_img, int n_cand);
int score_select(cmdb_bank *b, int B, int P_img, bool local_m_star);
int score_reweight(cmdb_bank *b, int B, int P_img, bool fused);
int score_build_knn_table(cmdb_bank *b, long long row_first, long long row_count);
int score_exact_scan(cmdb_bank *b, const float *q_dev, int P, unsigned long long *keys_dev);
int score_shard_lookup(cmdb_bank *b, int B, float *contrib_dev);
int score_shard_final(cmdb_bank *b, int B, const float *d2_sum_dev);
// peer-memory form of the two exchanges of a sharded round (no NCCL): see score_tail.cu
struct PeerPtrs {
    unsigned char *p[kMaxRanks];
};
int score_shard_exchange_keys(cmdb_bank *b, int B, int P_img, const PeerPtrs &peers, int world, int rank, int slot, unsigned long long epoch);
int score_shard_exchange_d2(cmdb_bank *b, int B, const PeerPtrs &peers, int world, int rank, int slot, unsigned long long epoch,
                            const float *contrib_dev, float *d2_sum_dev);
int score_merge_top3(cmdb_bank *b, int n_ranks, int B);
int score_final(cmdb_bank *b, int B);
int upsample_blur_launch(cudaStream_t stream, int n_img, int img_first, int img_step, size_t map_stride, const float *map_dev,
                         int fh, int fw, int out_hw, float *pre_dev, float *out_dev, unsigned char *u8_dev,
                         unsigned char *tmp_dev, float *mx_dev, float *mx_part_dev /* [n images][16] band maxima */);

}  // namespace cmdb
