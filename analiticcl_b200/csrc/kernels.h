// kernels.h -- launch interface of the sm_100a kernels of the variant-lookup path (see kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"

namespace anl {

// Buffers of one launch over `n` queries (all device pointers).
struct LaunchBuffers {
  const uint8_t* queries = nullptr;  // [n_total][query_stride] encoded query rows
  const uint32_t* qlist = nullptr;   // optional: indices into `queries` (rerun of selected queries); nullptr = identity
  const uint8_t* qblob = nullptr;    // optional: raw UTF-8 of the queries (confusable triage + edit scripts)
  const uint32_t* qboff = nullptr;   //           byte offsets into qblob, n_total + 1 entries
  uint32_t* rec_query = nullptr;     // optional (confusables): per pool record the row of its query, bp.pool_cap entries
  ConfWork* conf_work = nullptr;     // optional: queue of (record, query) pairs for the confusable kernel, bp.pool_cap entries
  uint32_t n = 0;                    // number of queries in this launch
  uint32_t* hits = nullptr;          // [n][hit_cap] gather ids of candidate instances
  uint32_t* hit_count = nullptr;     // [n]
  uint32_t* qflags = nullptr;        // [n] QF_* bits
  OutRec* out = nullptr;             // packed result pool, bp.pool_cap records
  uint32_t* out_gid = nullptr;       // optional (sharded mode): global gather id per pool record
  OutHead* out_head = nullptr;       // [n] per-query header: offset / count into the pool, max_freq
  void* scratch = nullptr;           // score kernel scratch: score_scratch_bytes(...)
  unsigned int* work = nullptr;      // [0],[1] work-stealing counters, [2] pool cursor, [3] confusable queue length,
                                     // [4] staged-node queue length, [5] exact-stage work counter, [8..17] pair-list path
                                     // (zeroed by the launchers); WORK_SLOTS entries
  // pair-list score stage (launch_score_pairs): dense slot base per query, the shape-sorted pair list, packed features
  uint32_t* qbase = nullptr;         // [n]
  uint4* pairs = nullptr;            // [pair_cap] by sorted position: {query (0xFFFFFFFF = no pair), candidate (gather id),
                                     //             dense slot = qbase[query] + position in its hit list, 0}
  uint32_t* pair_res = nullptr;      // [pair_cap] by dense slot: ld | lcs << 8 | prefix << 16 | suffix << 24, or 0xFFFFFFFF
  uint32_t pair_cap = 0;
  uint32_t* pair_hist = nullptr;     // [PAIR_TABLE] pairs per shape
  uint32_t* pair_first = nullptr;    // [PAIR_TABLE] first sorted position of a shape
  uint32_t* pair_cursor = nullptr;   // [PAIR_TABLE]
  QEntry* queue = nullptr;           // optional (split probe path): staged nodes of the whole launch, queue_cap entries
  uint32_t queue_cap = 0;
  QCtx* qctx = nullptr;              //          per-query context for the exact stage, [n]
  Counters* counters = nullptr;      // accumulated work counters (zeroed by the caller when wanted)
  cudaStream_t aux_stream = nullptr;    // optional side stream (+ two events): launch_score runs the long-query class on it,
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // beside the short-query class, when `scratch` has room for both grids
  size_t scratch_bytes = 0;             // size of `scratch` (only needed with aux_stream)
  cudaEvent_t ev_bloom_done = nullptr;  // optional: recorded by launch_probe between the Bloom and the exact stage
};

// Query normalisation on the device (normalize_to_alphabet, src/anahash.rs:50-80): raw text -> encoded rows.
cudaError_t launch_encode(const DeviceIndex* d_ix, const BatchParams& bp, const uint8_t* qblob, const uint32_t* qboff, uint32_t n,
                          uint8_t* rows, uint8_t* status, cudaStream_t stream);
// Candidate generation: deletion neighbourhood x insertion multisets -> Bloom -> table -> postings.
// With lb.queue set (and no stop-at-exact-match) the work is split over two kernels: the Bloom stage appends
// the nodes that pass the filter to a global queue, and exact_kernel looks them up one node per lane.  The
// caller must check work[4] <= queue_cap afterwards (else: run again without the queue).
cudaError_t launch_probe(const DeviceIndex* d_ix, const DeviceIndex& h_ix, const BatchParams& bp, const LaunchBuffers& lb,
                         int sm_count, cudaStream_t stream);
// Bit-parallel OSA prefilter of the hit lists (exact rejection of candidates far beyond the edit distance);
// optional, run between launch_probe and launch_score.
cudaError_t launch_prefilter(const DeviceIndex* d_ix, const BatchParams& bp, const LaunchBuffers& lb, int sm_count, cudaStream_t stream);
// Scoring (true Damerau-Levenshtein, LCS, prefix, suffix, case, f64 score) fused with ranking,
// cropping and cut-off.
cudaError_t launch_score(const DeviceIndex* d_ix, const DeviceIndex& h_ix, const BatchParams& bp, const LaunchBuffers& lb,
                         int sm_count, cudaStream_t stream);
// The same stage (prefilter + score) over a global, shape-sorted pair list: every lane of the DP holds a pair of the
// same matrix shape, whatever query it belongs to.  ev_filter_done (optional) is recorded after the filter / sort
// kernels.  The caller checks work[WORK_PAIR_TOTAL] <= lb.pair_cap afterwards.
cudaError_t launch_score_pairs(const DeviceIndex* d_ix, const DeviceIndex& h_ix, const BatchParams& bp, const LaunchBuffers& lb,
                               int sm_count, cudaStream_t stream, cudaEvent_t ev_filter_done);
static const int WORK_SLOTS = 24;       // entries of LaunchBuffers::work
static const int WORK_PAIR_TOTAL = 8;   // pairs that passed the filter = dense slots needed
static const uint32_t PAIR_TABLE = 5120;  // entries of the per-shape tables
// Confusable rescoring on the device (only when launch_score filled lb.conf_work): edit script + pattern
// matching per queued pair, then re-rank / crop / cut-off per query in place.  Queries the device cannot
// settle get HEAD_HOST_FINISH in their header.
cudaError_t launch_confusables(const DeviceIndex* d_ix, const BatchParams& bp, const LaunchBuffers& lb, int sm_count,
                               cudaStream_t stream);
cudaError_t launch_finish(const BatchParams& bp, const LaunchBuffers& lb, int sm_count, cudaStream_t stream);
size_t score_scratch_bytes(const BatchParams& bp, int sm_count, uint32_t n_queries);
// device-only forms of the index, built once at upload: 32-byte anagram records; single postings folded into their slots
cudaError_t finish_device_index(const Key192* ana_key, const uint32_t* ana_inst_off, uint32_t n_anagrams, AnaRec* ana_rec, Slot* table,
                                uint64_t slots, const uint32_t* post_ana, const uint8_t* post_cls, cudaStream_t stream);
cudaError_t configure_kernels();
// Unicode Alphabetic ranges ([n][2], inclusive) for the character classes of the device edit script; per device
cudaError_t upload_alphabetic_ranges(const uint32_t* ranges, uint32_t n);
unsigned long long kernel_launches();  // process-wide count of kernel launches issued by this library
// Lexicon-sharded mode: merge the all-gathered per-shard survivor lists (see merge_kernel).
cudaError_t launch_shard_flagcheck(const uint32_t* flags_all, uint32_t n, uint32_t n_shards, uint64_t stride, unsigned int* res,
                                   int sm_count, cudaStream_t stream);
cudaError_t launch_merge(const BatchParams& bp, uint32_t n, uint32_t n_shards, const OutHead* heads_all, uint32_t head_stride,
                         const OutRec* recs_all, const uint32_t* gids_all, uint32_t rec_stride, const uint32_t* qflags_in, uint32_t* qflags,
                         OutRec* out, OutHead* out_head, void* scratch, uint32_t scratch_cap, unsigned int* work, int sm_count,
                         cudaStream_t stream);
size_t merge_scratch_bytes(int sm_count, uint32_t n, uint32_t scratch_cap);

// Export stage (export.cu): packed pool + headers -> the caller-visible arrays in query order.
struct ExportBuffers {
  uint32_t n = 0;
  const OutHead* head = nullptr;        // [n]
  const uint32_t* qflags = nullptr;     // [n] QF_*
  const uint8_t* enc_status = nullptr;  // optional [n] ENC_* of the encode kernel
  const OutRec* pool = nullptr;
  uint32_t* tile_sum = nullptr;         // [export_tiles(n)]
  uint32_t* loff = nullptr;             // [n + 1] batch-local CSR offsets
  uint32_t* oflags = nullptr;           // [n] ANL_QUERY_* flags
  void* out = nullptr;                  // anl_variant[out_cap]
  uint32_t out_cap = 0;
  ExportSummary* summary = nullptr;
};
uint32_t export_tiles(uint32_t n);
cudaError_t launch_export(const ExportBuffers& eb, cudaStream_t stream);
// results of re-run queries -> the batch's pool (records behind `base`, headers and flags replaced)
cudaError_t launch_patch(uint32_t m, const uint32_t* qlist, const OutHead* rr_head, const uint32_t* rr_qflags, const OutRec* rr_out,
                         OutHead* head, uint32_t* qflags, OutRec* pool, uint32_t base, uint32_t pool_cap, unsigned int* pool_cursor,
                         uint32_t rr_total, cudaStream_t stream);
// off64[i] = base + loff[i], i < n: the batch's slice of the call-wide u64 offsets
cudaError_t launch_offsets(uint32_t n, const uint32_t* loff, uint64_t base, uint64_t* off64, cudaStream_t stream);

}  // namespace anl
