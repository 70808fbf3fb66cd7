// kernels.h -- launch interface of the two sm_100a kernels of the variant-lookup path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"

namespace anl {

// Buffers of one launch over `n` queries (all device pointers).
struct LaunchBuffers {
  const uint8_t* queries;  // [n_total][query_stride] encoded query rows
  const uint32_t* qlist;   // optional: indices into `queries` (rerun of selected queries); nullptr = identity
  uint32_t n;              // number of queries in this launch
  uint32_t* hits;          // [n][hit_cap] gather ids of candidate instances
  uint32_t* hit_count;     // [n]
  uint32_t* qflags;        // [n] QF_* bits
  OutRec* out;             // packed result pool, bp.pool_cap records
  OutHead* out_head;       // [n] per-query header: offset / count into the pool, max_freq
  void* scratch;           // score kernel scratch: score_scratch_bytes(...)
  unsigned int* work;      // [0],[1] work-stealing counters, [2] pool cursor (zeroed by the launchers)
  Counters* counters;      // accumulated work counters (zeroed by the caller when wanted)
};

// Candidate generation: deletion neighbourhood x insertion multisets -> Bloom -> table -> postings.
cudaError_t launch_probe(const DeviceIndex* d_ix, const DeviceIndex& h_ix, const BatchParams& bp, const LaunchBuffers& lb,
                         int sm_count, cudaStream_t stream);
// Scoring (true Damerau-Levenshtein, LCS, prefix, suffix, case, f64 score) fused with ranking,
// cropping and cut-off.
cudaError_t launch_score(const DeviceIndex* d_ix, const DeviceIndex& h_ix, const BatchParams& bp, const LaunchBuffers& lb,
                         int sm_count, cudaStream_t stream);
size_t score_scratch_bytes(const BatchParams& bp, int sm_count, uint32_t n_queries);
cudaError_t configure_kernels();

}  // namespace anl
