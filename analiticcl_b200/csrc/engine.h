// engine.h -- device residency and the batched lookup pipeline (H2D, two kernels, D2H, host post-pass).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/analiticcl_b200.h"
#include "device_types.h"
#include "host_model.h"
#include "kernels.h"

namespace anl {

struct ResultSet {
  std::vector<uint64_t> offsets;  // n + 1
  std::vector<anl_variant> variants;
  std::vector<uint32_t> flags;  // per query: bit 0 = empty input
};

// Everything one pass over a batch of queries needs on the device.
struct DeviceBatch {
  uint32_t n = 0;
  BatchParams bp;
  anl_search_params params;
  // host copies
  std::string blob;               // raw queries (needed by the confusable post-pass)
  std::vector<uint64_t> offsets;  // n + 1
  std::vector<uint8_t> host_flags;  // per query: 1 = resolved on the host as empty, 2 = unsupported length
  uint8_t* h_rows = nullptr;        // pinned, [n][stride]
  // device buffers
  uint8_t* d_rows = nullptr;
  uint32_t* d_hits = nullptr;
  uint32_t* d_hit_count = nullptr;
  uint32_t* d_qflags = nullptr;
  OutRec* d_out = nullptr;
  uint32_t* d_out_count = nullptr;
  void* d_scratch = nullptr;
  unsigned int* d_work = nullptr;
  Counters* d_counters = nullptr;
  // one event triple (start, after probe, after score) per run since the last timings() call
  std::vector<cudaEvent_t> events;
  uint32_t runs_recorded = 0;
  cudaEvent_t last_done = nullptr;  // end event of the most recent run
  bool ran = false;
  uint64_t reruns = 0;
  uint64_t results = 0;
};

class Engine {
 public:
  explicit Engine(HostModel* hm) : hm_(hm) {}
  ~Engine();
  bool upload(int device, std::string* err);  // device side of build()
  bool ensure_msets(uint32_t J, std::string* err);
  bool make_batch_params(const anl_search_params& p, BatchParams* bp, uint32_t* needed_j, std::string* err) const;

  DeviceBatch* create_batch(const char* blob, const uint64_t* offsets, uint64_t n, const anl_search_params& p,
                            std::string* err, int* status);
  bool run_batch(DeviceBatch* b, cudaStream_t stream, std::string* err);
  bool fetch_batch(DeviceBatch* b, ResultSet* out, std::string* err, int* status);
  void free_batch(DeviceBatch* b);
  bool timings(DeviceBatch* b, float* probe_ms, float* score_ms, std::string* err);
  bool counters(DeviceBatch* b, anl_counters* out, std::string* err);

  // whole pipeline with internal chunking
  bool find_variants_batch(const char* blob, const uint64_t* offsets, uint64_t n, const anl_search_params& p, ResultSet* out,
                           std::string* err, int* status);

  const DeviceIndex& host_view() const { return h_ix_; }
  bool uploaded() const { return d_ix_ != nullptr; }
  cudaStream_t stream() const { return stream_; }

 private:
  // post-pass of one query's device records -> final variants (confusables, re-sort, cut-off)
  void finish_query(const DeviceBatch& b, uint64_t qi, const OutRec* recs, uint32_t count, std::vector<anl_variant>* out) const;
  bool rerun_overflow(DeviceBatch* b, const std::vector<uint32_t>& which, const std::vector<uint32_t>& hit_counts,
                      const std::vector<uint32_t>& out_counts, std::vector<std::vector<OutRec>>* recs, std::string* err,
                      int* status);
  template <class T>
  bool dev_alloc(T** p, size_t count, std::string* err);
  void release_index();

  HostModel* hm_;
  int device_ = -1;
  int sm_count_ = 148;
  cudaStream_t stream_ = nullptr;
  DeviceIndex h_ix_{};
  DeviceIndex* d_ix_ = nullptr;
  std::vector<void*> index_allocs_;
  void* d_mset_ = nullptr;
};

}  // namespace anl
