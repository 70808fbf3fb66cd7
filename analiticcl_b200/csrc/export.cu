// export.cu -- the last stage of a pass: the packed result pool of a batch becomes the caller-visible result arrays.
//
// The score / finish kernels leave, per query, a header {max_freq, offset, count} and `count` 16-byte records
// {dist_score, vocab id, raw frequency} somewhere in the packed pool (reservation order, not query order).  The
// C ABI returns Vec<Vec<VariantResult>> (src/types.rs:326-332) as a CSR: u64 offsets[n + 1] + 32-byte records
// {vocab_id, dist_score, freq_score, via} in query order.  That final form is produced HERE, on the device, so the
// host side of anl_find_variants_batch does no per-record (and, in the common case, no per-query) work: the arrays
// are DMA-ed straight into the pinned result set.
//
//   count_kernel   : per tile of EXPORT_TILE queries, the number of records (+ the batch summary: queries that
//                    need a re-run or a host finish)
//   export_kernel  : tile base = sum of the tile counts before it; exclusive scan inside the tile; records written
//                    in query order with the frequency normalised (freq / max_freq, the same IEEE division as
//                    src/lib.rs:1523); per-query API flags
//   offsets_kernel : local u32 offsets + the call-wide base -> u64 offsets (run when the base is known)
#include "kernels.h"

#include "../../include/analiticcl_b200.h"
#include "kernel_common.cuh"

namespace anl {

constexpr uint32_t EXPORT_THREADS = 256;

__device__ __forceinline__ uint32_t query_count(const OutHead& h, uint32_t qf) {
  // a query that will be re-run (hit list overflow) or was skipped holds no records yet
  if (qf & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED | QF_OUT_OVERFLOW)) return 0;
  return h.count & ~HEAD_HOST_FINISH;
}

__global__ void __launch_bounds__(EXPORT_THREADS)
count_kernel(uint32_t n, const OutHead* __restrict__ head, const uint32_t* __restrict__ qflags, uint32_t* __restrict__ tile_sum,
             ExportSummary* __restrict__ summary) {
  __shared__ uint32_t s_red[3][EXPORT_THREADS / 32];
  const uint32_t q0 = blockIdx.x * EXPORT_TILE;
  uint32_t sum = 0, rerun = 0, hostfin = 0;
  for (uint32_t k = threadIdx.x; k < EXPORT_TILE; k += EXPORT_THREADS) {
    const uint32_t q = q0 + k;
    if (q >= n) break;
    const uint32_t qf = qflags[q];
    const OutHead h = head[q];
    sum += query_count(h, qf);
    rerun += (qf & (QF_HIT_OVERFLOW | QF_UNSUPPORTED)) == QF_HIT_OVERFLOW;  // (a pool overflow shows in the pool cursor)
    hostfin += !(qf & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED | QF_OUT_OVERFLOW)) && (h.count & HEAD_HOST_FINISH);
  }
  sum = __reduce_add_sync(FULL, sum);
  rerun = __reduce_add_sync(FULL, rerun);
  hostfin = __reduce_add_sync(FULL, hostfin);
  const uint32_t w = threadIdx.x >> 5;
  if (lane_id() == 0) {
    s_red[0][w] = sum;
    s_red[1][w] = rerun;
    s_red[2][w] = hostfin;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0, b = 0, c = 0;
    for (uint32_t i = 0; i < EXPORT_THREADS / 32; ++i) {
      a += s_red[0][i];
      b += s_red[1][i];
      c += s_red[2][i];
    }
    tile_sum[blockIdx.x] = a;
    atomicAdd(&summary->total, a);
    if (b) atomicAdd(&summary->n_rerun, b);
    if (c) atomicAdd(&summary->n_host_finish, c);
  }
}

__global__ void __launch_bounds__(EXPORT_THREADS)
export_kernel(uint32_t n, const OutHead* __restrict__ head, const uint32_t* __restrict__ qflags, const uint8_t* __restrict__ enc_status,
              const OutRec* __restrict__ pool, const uint32_t* __restrict__ tile_sum, uint32_t* __restrict__ loff,
              uint32_t* __restrict__ oflags, anl_variant* __restrict__ out, uint32_t out_cap) {
  __shared__ uint32_t s_cnt[EXPORT_TILE];
  __shared__ uint32_t s_warp[EXPORT_THREADS / 32];
  __shared__ uint32_t s_base;
  const uint32_t q0 = blockIdx.x * EXPORT_TILE;
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  // base of this tile: the records of all tiles before it
  {
    uint32_t b = 0;
    for (uint32_t t = threadIdx.x; t < blockIdx.x; t += EXPORT_THREADS) b += tile_sum[t];
    b = __reduce_add_sync(FULL, b);
    if (lane == 0) s_warp[w] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t a = 0;
      for (uint32_t i = 0; i < EXPORT_THREADS / 32; ++i) a += s_warp[i];
      s_base = a;
    }
    __syncthreads();
  }
  // exclusive scan of the tile's counts: every thread owns EXPORT_TILE / EXPORT_THREADS consecutive queries
  constexpr uint32_t PER = EXPORT_TILE / EXPORT_THREADS;
  uint32_t cnt[PER], mine = 0;
#pragma unroll
  for (uint32_t k = 0; k < PER; ++k) {
    const uint32_t q = q0 + threadIdx.x * PER + k;
    uint32_t c = 0, of = 0;
    if (q < n) {
      const uint32_t qf = qflags[q];
      c = query_count(head[q], qf);
      const uint32_t es = enc_status ? enc_status[q] : (uint32_t)ENC_OK;
      if ((qf & QF_EMPTY) && es == ENC_OK) of |= ANL_QUERY_EMPTY;
      if ((qf & QF_UNSUPPORTED) || es == ENC_TOO_LONG_UNSUPPORTED) of |= ANL_QUERY_UNSUPPORTED;
      oflags[q] = of;
    }
    cnt[k] = c;
    mine += c;
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  __syncthreads();  // (s_warp is reused)
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  uint32_t before = 0;
  for (uint32_t i = 0; i < w; ++i) before += s_warp[i];
  uint32_t run = s_base + before + incl - mine;
#pragma unroll
  for (uint32_t k = 0; k < PER; ++k) {
    const uint32_t q = q0 + threadIdx.x * PER + k;
    s_cnt[threadIdx.x * PER + k] = run;
    if (q < n) loff[q] = run;
    run += cnt[k];
  }
  if (q0 + EXPORT_TILE >= n && threadIdx.x == EXPORT_THREADS - 1) loff[n] = run;  // (the last tile closes the CSR)
  __syncthreads();
  // records in query order: a group of 8 lanes per query (most queries hold a handful of records)
  const uint32_t group = threadIdx.x >> 3, gl = threadIdx.x & 7;
  for (uint32_t k = group; k < EXPORT_TILE; k += EXPORT_THREADS / 8) {
    const uint32_t q = q0 + k;
    if (q >= n) break;
    const uint32_t qf = qflags[q];
    const OutHead h = head[q];
    const uint32_t c = query_count(h, qf);
    const uint32_t dst = s_cnt[k];
    const OutRec* __restrict__ src = pool + h.offset;
    for (uint32_t i = gl; i < c; i += 8) {
      if (dst + i >= out_cap) break;  // (cannot happen: the caller sizes `out` like the pool)
      const OutRec r = src[i];
      const double f = (double)r.freq;
      anl_variant v;
      v.vocab_id = r.vocab_id & ~OUT_SKIP_CONFUSABLES;
      v.dist_score = r.dist_score;
      v.freq_score = h.max_freq > 0.0 ? __ddiv_rn(f, h.max_freq) : f;  // src/lib.rs:1521-1525
      v.via = ANL_NO_VIA;
      out[dst + i] = v;
    }
  }
}

__global__ void offsets_kernel(uint32_t n, const uint32_t* __restrict__ loff, uint64_t base, uint64_t* __restrict__ off64) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off64[i] = base + loff[i];
}

// Results of re-run queries (hit-list overflow: run again with an exact capacity into buffers of their own) join
// the batch's pool: records appended behind the pool cursor, header and flags of the query replaced.  One warp per
// re-run query.
__global__ void __launch_bounds__(128)
patch_kernel(uint32_t m, const uint32_t* __restrict__ qlist, const OutHead* __restrict__ rr_head, const uint32_t* __restrict__ rr_qflags,
             const OutRec* __restrict__ rr_out, OutHead* __restrict__ head, uint32_t* __restrict__ qflags, OutRec* __restrict__ pool,
             uint32_t base, uint32_t pool_cap, unsigned int* pool_cursor, uint32_t rr_total) {
  const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(pool_cursor, rr_total);
  if (k >= m) return;
  const OutHead h = rr_head[k];
  const uint32_t c = h.count & ~HEAD_HOST_FINISH;
  const bool fits = (unsigned long long)base + h.offset + c <= (unsigned long long)pool_cap;  // (the host grew the pool first)
  if (fits)
    for (uint32_t i = lane; i < c; i += 32) pool[base + h.offset + i] = rr_out[h.offset + i];
  if (lane == 0) {
    const uint32_t q = qlist[k];
    OutHead o;
    o.max_freq = h.max_freq;
    o.offset = base + h.offset;
    o.count = fits ? h.count : 0;
    head[q] = o;
    qflags[q] = rr_qflags[k] | (fits ? 0u : QF_OUT_OVERFLOW);
  }
}

cudaError_t launch_patch(uint32_t m, const uint32_t* qlist, const OutHead* rr_head, const uint32_t* rr_qflags, const OutRec* rr_out,
                         OutHead* head, uint32_t* qflags, OutRec* pool, uint32_t base, uint32_t pool_cap, unsigned int* pool_cursor,
                         uint32_t rr_total, cudaStream_t stream) {
  if (m == 0) return cudaSuccess;
  patch_kernel<<<(m * 32 + 127) / 128, 128, 0, stream>>>(m, qlist, rr_head, rr_qflags, rr_out, head, qflags, pool, base, pool_cap,
                                                         pool_cursor, rr_total);
  count_launch(1);
  return cudaGetLastError();
}

uint32_t export_tiles(uint32_t n) { return (n + EXPORT_TILE - 1) / EXPORT_TILE; }

cudaError_t launch_export(const ExportBuffers& eb, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(eb.summary, 0, sizeof(ExportSummary), stream);
  if (e != cudaSuccess) return e;
  if (eb.n == 0) return cudaMemsetAsync(eb.loff, 0, sizeof(uint32_t), stream);
  const uint32_t tiles = export_tiles(eb.n);
  count_kernel<<<tiles, EXPORT_THREADS, 0, stream>>>(eb.n, eb.head, eb.qflags, eb.tile_sum, eb.summary);
  export_kernel<<<tiles, EXPORT_THREADS, 0, stream>>>(eb.n, eb.head, eb.qflags, eb.enc_status, eb.pool, eb.tile_sum, eb.loff,
                                                      eb.oflags, reinterpret_cast<anl_variant*>(eb.out), eb.out_cap);
  count_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_offsets(uint32_t n, const uint32_t* loff, uint64_t base, uint64_t* off64, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  offsets_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, loff, base, off64);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace anl
