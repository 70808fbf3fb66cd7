// engine.h -- device residency of the index and the batched lookup pipeline (H2D, kernels, export, D2H).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/analiticcl_b200.h"
#include "device_types.h"
#include "host_model.h"
#include "hostpool.h"
#include "kernels.h"

namespace anl {

// Vec<Vec<VariantResult>> as a CSR in DMA-able host memory: the export stage's arrays land here by D2H copy.
struct ResultSet {
  DmaBuffer<uint64_t> offsets;  // n + 1
  DmaBuffer<anl_variant> variants;
  DmaBuffer<uint32_t> flags;  // per query: ANL_QUERY_* bits
};

// Everything one pass over a batch of queries needs, on the device and in pinned host memory.
// Buffers are capacity-based and recycled through Engine's cache: steady-state calls allocate nothing.
static const int EV_PER_RUN = 8;
struct DeviceBatch {
  // ---- capacities (what the buffers can hold) ----
  uint32_t cap_n = 0, cap_pool = 0;
  size_t cap_rows_bytes = 0, cap_hits = 0, cap_scratch = 0;
  // ---- current contents ----
  uint32_t n = 0;
  BatchParams bp;
  anl_search_params params;
  const char* blob = nullptr;  // raw queries (borrowed for the duration of the call, or owned_blob)
  std::string owned_blob;
  std::vector<uint64_t> offsets;    // n + 1, relative to blob
  std::vector<uint8_t> host_flags;  // per query: ENC_* (1 = too long, empty by construction; 2 = too long, unsupported)
  // pinned host buffers
  uint8_t* h_rows = nullptr;   // [cap_n][cap_stride]
  OutHead* h_head = nullptr;   // [cap_n]
  uint32_t* h_flags = nullptr; // [cap_n]
  uint32_t* h_hitcnt = nullptr;// [cap_n]
  OutRec* h_out = nullptr;     // [cap_pool]
  unsigned int* h_work = nullptr;  // [WORK_SLOTS]
  ExportSummary* h_summary = nullptr;
  // device buffers
  uint8_t* d_rows = nullptr;
  uint8_t* d_qblob = nullptr;   // raw query bytes (encode kernel, confusable stage), grow-only
  uint32_t* d_qboff = nullptr;  // n + 1 byte offsets into d_qblob
  uint32_t* h_qboff = nullptr;  // pinned staging of the offsets
  char* h_qblob = nullptr;      // pinned staging of the query text
  size_t cap_qblob = 0, cap_qboff = 0;
  bool has_qblob = false;
  QEntry* d_queue = nullptr;        // split probe path: staged nodes between the Bloom stage and the exact stage
  QCtx* d_qctx = nullptr;           //                   per-query context of the exact stage
  size_t cap_queue = 0, cap_qctx = 0;
  bool split = false;               // use the split probe path for this batch
  uint8_t* d_enc_status = nullptr;  // per query: ENC_* result of the encode kernel (or of the host encoder)
  uint8_t* h_enc_status = nullptr;
  bool dev_encode = false;          // the rows of this batch were encoded on the device
  uint32_t* d_rec_query = nullptr;  // per pool record: the row of its query (confusable triage), sized like d_out
  ConfWork* d_conf_work = nullptr;  // queue of (record, query) pairs for the confusable kernel, sized like d_out
  bool dev_conf = false;            // this batch's confusables are rescored on the device (HEAD_HOST_FINISH marks the rest)
  uint32_t* d_hits = nullptr;
  uint32_t* d_hit_count = nullptr;
  uint32_t* d_qflags = nullptr;
  OutRec* d_out = nullptr;
  uint32_t* d_gid = nullptr;  // sharded mode: global gather id per pool record (sized like d_out)
  OutHead* d_head = nullptr;
  void* d_scratch = nullptr;
  unsigned int* d_work = nullptr;
  Counters* d_counters = nullptr;
  // pair-list score stage (kernels.cu "Kernels 2p"): dense slot base per query, shape-sorted pair list, packed features
  bool use_pairs = false;
  uint32_t* d_qbase = nullptr;      // [cap_n]
  uint4* d_pairs = nullptr;         // [cap_pairs]
  uint32_t* d_pair_res = nullptr;   // [cap_pairs]
  uint32_t* d_pair_tab = nullptr;   // [3][PAIR_TABLE]: histogram, first position, cursor per shape
  size_t cap_pairs = 0;
  // export stage (export.cu): final arrays of this batch, in query order
  anl_variant* d_final = nullptr;   // [cap_pool]
  uint32_t* d_loff = nullptr;       // [cap_n + 1] batch-local CSR offsets
  uint32_t* d_oflags = nullptr;     // [cap_n]
  uint64_t* d_off64 = nullptr;      // [cap_n] call-wide offsets (written when the batch's base is known)
  uint32_t* d_tile_sum = nullptr;   // [export_tiles(cap_n)]
  ExportSummary* d_summary = nullptr;
  // buffers of the hit-overflow rerun (queries with more than hit_cap candidate instances), grow-only
  uint32_t* rr_qlist = nullptr;
  uint32_t* rr_hits = nullptr;
  uint32_t* rr_hit_count = nullptr;
  uint32_t* rr_qflags = nullptr;
  OutHead* rr_head = nullptr;
  OutRec* rr_out = nullptr;
  ConfWork* rr_conf_work = nullptr;
  uint32_t* rr_rec_query = nullptr;
  uint8_t* rr_scratch = nullptr;
  unsigned int* rr_work = nullptr;  // [8] the rerun's own counters (the batch's hold its pool cursor)
  size_t rr_cap_m = 0, rr_cap_hits = 0, rr_cap_pool = 0, rr_cap_scratch = 0;
  cudaStream_t stream = nullptr;    // this batch's own stream (copies + default launches)
  cudaEvent_t uploaded = nullptr;   // recorded after the H2D copy of the query rows
  cudaStream_t aux = nullptr;       // side stream: the long-query class of the score kernel runs beside the short one
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_done = nullptr;    // end of the most recent chain that is not one of the per-run timing events
  // EV_PER_RUN events (start, after the Bloom stage, after probe, after prefilter, after score, after confusables,
  // after finish, after export) per run since the last timings() call
  std::vector<cudaEvent_t> events;
  uint32_t runs_recorded = 0;
  cudaEvent_t last_done = nullptr;  // end event of the most recent run
  bool ran = false;
  bool settled = false;        // the summary of the last run has been read and every exception it reported is dealt with
  bool sharded = false;        // built against a lexicon shard: results come from shard_export / shard_merge
  bool merged = false;         // shard_merge has produced the final lists in d_out / d_head
  int final_mode = FINISH_FULL;  // finish mode of the merged result (bp.finish_mode is FINISH_SHARD while scoring)
  uint64_t reruns = 0;
  uint64_t results = 0;
};

struct ShardComm;  // shard_comm.cu: NCCL communicator + receive buffers of the lexicon-sharded mode
struct ShardStepStats {
  float score_ms, exchange_ms, merge_ms;
  uint64_t bytes_received, records_local, records_total;
};
bool shard_unique_id(uint8_t* id128, std::string* err);

class Engine {
 public:
  explicit Engine(HostModel* hm) : hm_(hm) {}
  ~Engine();
  bool upload(int device, std::string* err);  // device side of build()
  bool ensure_msets(uint32_t J, std::string* err);
  bool ensure_confusable_table(std::string* err);  // device copy of the confusable prefilter table + vocabulary text
  bool make_batch_params(const anl_search_params& p, BatchParams* bp, uint32_t* needed_j, std::string* err) const;

  // copy_blob: keep a private copy of the query text (device-batch API) or borrow it (one-shot call)
  // sync: wait for the H2D copy before returning (otherwise it is ordered before the kernels by an event)
  DeviceBatch* create_batch(const char* blob, const uint64_t* offsets, uint64_t n, const anl_search_params& p,
                            bool copy_blob, bool sync, std::string* err, int* status);
  // one pass: probe, prefilter, score, confusables, finish, export -- ends with the (async) download of the summary
  bool run_batch(DeviceBatch* b, cudaStream_t stream, std::string* err);
  // Waits for the run and deals with everything its summary reports: result pool / staged-node queue overflows
  // (re-run with larger buffers) and hit-list overflows (those queries are run again with an exact capacity and
  // patched into the pool).  Afterwards the export arrays of the batch are final unless needs_host_finish().
  bool settle(DeviceBatch* b, std::string* err, int* status);
  bool needs_host_finish(const DeviceBatch* b) const;
  uint32_t export_total(const DeviceBatch* b) const { return b->h_summary->total; }
  // fast path: D2H of the export arrays into out[qbase ..] / variants[vbase ..] on the batch's stream (no wait)
  bool issue_download(DeviceBatch* b, ResultSet* out, uint64_t qbase, uint64_t vbase, std::string* err);
  // slow path: queries the device could not finish go through the host post-pass; writes the batch's part of `out`
  // (whose variants may grow) and returns the number of variants written
  bool finish_on_host(DeviceBatch* b, ResultSet* out, uint64_t qbase, uint64_t vbase, uint64_t* written, std::string* err,
                      int* status);
  bool wait_download(DeviceBatch* b, std::string* err);
  // settle + download + wait into a fresh result set (device-batch API, sharded merge)
  bool fetch_batch(DeviceBatch* b, ResultSet* out, std::string* err, int* status);
  void free_batch(DeviceBatch* b);  // returns the buffers to the cache
  // stage_ms[7]: Bloom stage, exact stage (the whole fused probe kernel when the split path is off), prefilter,
  // score/rank, confusables, finish, export
  bool timings(DeviceBatch* b, float* stage_ms, std::string* err);
  bool counters(DeviceBatch* b, anl_counters* out, std::string* err);

  // lexicon-sharded mode (see include/analiticcl_b200.h)
  bool shard_export_size(DeviceBatch* b, uint64_t* n_records, uint32_t* max_per_query, std::string* err, int* status);
  bool shard_export(DeviceBatch* b, void* d_heads, void* d_records, void* d_gids, void* d_flags, std::string* err);
  bool shard_merge(DeviceBatch* b, uint32_t n_shards, const void* d_heads_all, const void* d_records_all,
                   const void* d_gids_all, const void* d_flags_all, uint64_t record_stride, uint32_t max_survivors,
                   ResultSet* out, std::string* err, int* status);

  // the same with the exchange inside the library (shard_comm.cu): NCCL communicator over the shards' ranks, then per
  // batch: score against the shard, exchange all shards' survivors, merge.  out == nullptr: results stay on the device.
  bool shard_comm_init(const uint8_t* id128, int rank, int n_ranks, std::string* err);
  void shard_comm_free();
  bool shard_step(DeviceBatch* b, ResultSet* out, ShardStepStats* stats, std::string* err, int* status);
  bool shard_merge_strided(DeviceBatch* b, uint32_t n_shards, const void* d_heads_all, const void* d_records_all,
                           const void* d_gids_all, const void* d_flags_all, uint64_t query_stride, uint64_t record_stride,
                           uint32_t max_survivors, ResultSet* out, std::string* err, int* status);

  const DeviceIndex& host_view() const { return h_ix_; }
  bool uploaded() const { return d_ix_ != nullptr; }
  cudaStream_t stream() const { return stream_; }
  int device() const { return device_; }
  HostModel* host_model() const { return hm_; }

 private:
  // post-pass of one query's device records -> final variants (confusables, re-sort, cut-off); appends to `out`
  void finish_query_variants(const DeviceBatch& b, uint64_t qi, const OutRec* recs, uint32_t count, double max_freq,
                             std::vector<anl_variant>* out) const;
  void finish_query(const DeviceBatch& b, uint64_t qi, const OutRec* recs, uint32_t count, double max_freq,
                    std::vector<anl_variant>* out) const;
  bool rerun_overflowed(DeviceBatch* b, std::string* err, int* status);  // hit-list overflows: re-run, patch, re-export
  bool launch_export_chain(DeviceBatch* b, cudaStream_t st, std::string* err);
  bool ensure_capacity(DeviceBatch* b, uint32_t n, uint32_t stride, uint32_t hit_cap, uint32_t pool_cap, size_t scratch,
                       std::string* err);
  bool grow_pool(DeviceBatch* b, uint32_t pool_cap, bool keep_records, std::string* err);
  bool grow_pairs(DeviceBatch* b, size_t pair_cap, std::string* err);
  bool launch_score_stage(DeviceBatch* b, const LaunchBuffers& lb, cudaStream_t st, cudaEvent_t ev_filter, std::string* err);
  bool relaunch_from_score(DeviceBatch* b, std::string* err);  // score stage .. export + summary download, on the batch's stream
  void destroy_batch(DeviceBatch* b);
  void release_index();

  HostModel* hm_;
  int device_ = -1;
  int sm_count_ = 148;
  cudaStream_t stream_ = nullptr;
  DeviceIndex h_ix_{};
  DeviceIndex* d_ix_ = nullptr;
  std::vector<void*> index_allocs_;
  void* d_mset_ = nullptr;
  std::vector<void*> conf_allocs_;
  size_t conf_uploaded_ = (size_t)-1;  // number of confusables the device table was built from
  size_t conf_vocab_ = 0;              // vocabulary size the device text blob was built from
  std::mutex cache_m_;
  std::vector<DeviceBatch*> cache_;  // idle batches whose buffers can be reused
  ShardComm* shard_comm_ = nullptr;
};

// anl_find_variants_batch over one or more replicas of the index (one Engine per device): the batch is cut into chunks
// that go round-robin to the devices, every device is driven by its own host thread with several chunks in flight,
// and the chunks' result arrays land in `out` in query order.  src/bin/analiticcl.rs:418-482 (process_par) is the
// reference's counterpart: rayon over the queries of one process.
bool find_variants_batch_multi(const std::vector<Engine*>& engines, const char* blob, const uint64_t* offsets, uint64_t n,
                               const anl_search_params& p, ResultSet* out, std::string* err, int* status);

}  // namespace anl
