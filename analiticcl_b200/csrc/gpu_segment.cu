// gpu_segment.cu -- the batch producer of find_all_matches (src/search.rs:190-336, src/lib.rs:1822-1903) on the device.
//
// search.cpp segments a text on the host cores: token boundaries (maximal runs of non-alphabetic characters),
// boundary strengths, the hard-delimited batches and, per batch, every 1..max_ngram-gram as a byte span.  On running
// text that is a scan over every byte plus ~3 spans per token; here it is a handful of kernels over the text in HBM:
//   class_kernel      per byte: character start?  alphabetic?  (UTF-8 decoded in place; Unicode Alphabetic ranges)
//   flag kernels+scan boundary begins / ends -> boundary records, strengths (src/search.rs:238-258)
//   closer scan       hard boundaries that close a batch (src/lib.rs:1822) -> batch descriptors
//   count / emit      one thread per batch replays find_match_ngrams (src/search.rs:262-312) for every order: first the
//                     number of spans, then -- after a scan -- the spans themselves, in the producer's order
// The spans and the batch index come back to the host in one copy each; the lookups they feed are the GPU batches of
// anl_find_variants_batch as before.  Device-wide scans are CUB's (CUDA toolkit headers).
#include <cub/device/device_scan.cuh>

#include <cstring>
#include <mutex>

#include "kernel_common.cuh"
#include "search.h"
#include "unicode_tables.h"

namespace anl {

namespace {

#define GS_TRY(expr)                                                                      \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      *err = std::string("CUDA error in the device segmentation: ") + cudaGetErrorString(_e) + " at " #expr; \
      return false;                                                                       \
    }                                                                                     \
  } while (0)

__constant__ uint32_t c_seg_alpha[2 * 800];
__constant__ uint32_t c_seg_n_alpha;

__device__ __forceinline__ bool dev_is_alphabetic(uint32_t cp) {
  if (cp < 0x80) return (cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z');
  uint32_t lo = 0, hi = c_seg_n_alpha;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (cp > c_seg_alpha[2 * mid + 1]) lo = mid + 1; else hi = mid;
  }
  return lo < c_seg_n_alpha && cp >= c_seg_alpha[2 * lo];
}

// cls[i]: 0 = continuation byte, 1 = start of an alphabetic character, 2 = start of any other character
__global__ void class_kernel(const uint8_t* __restrict__ text, uint64_t n, uint8_t* __restrict__ cls) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = text[i];
  if ((c & 0xC0) == 0x80) {
    cls[i] = 0;
    return;
  }
  uint32_t l = c < 0x80 ? 1 : ((c & 0xE0) == 0xC0 ? 2 : ((c & 0xF0) == 0xE0 ? 3 : ((c & 0xF8) == 0xF0 ? 4 : 1)));
  if (i + l > n) l = (uint32_t)(n - i);
  uint32_t cp = c;
  if (l > 1) {
    cp = c & (0xFFu >> (l + 1));
    for (uint32_t k = 1; k < l; ++k) cp = (cp << 6) | (text[i + k] & 0x3Fu);
  }
  cls[i] = dev_is_alphabetic(cp) ? 1 : 2;
}
// class of the character before byte i (0 = none)
__device__ __forceinline__ uint32_t prev_class(const uint8_t* __restrict__ cls, uint64_t i) {
  while (i > 0) {
    --i;
    if (cls[i]) return cls[i];
  }
  return 0;
}
// a boundary begins at a non-alphabetic character that follows an alphabetic one (or the start of the text), and ends
// where the next alphabetic character starts
__global__ void flag_kernel(const uint8_t* __restrict__ cls, uint64_t n, uint32_t* __restrict__ fbegin, uint32_t* __restrict__ fend) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = cls[i];
  uint32_t b = 0, e = 0;
  if (c) {
    const uint32_t p = prev_class(cls, i);
    b = (c == 2 && p != 2) ? 1u : 0u;
    e = (c == 1 && p == 2) ? 1u : 0u;
  }
  fbegin[i] = b;
  fend[i] = e;
}
struct DBoundary {
  uint64_t begin, end;
  int32_t strength, pad;
};
__global__ void boundary_fill_kernel(const uint32_t* __restrict__ fbegin, const uint32_t* __restrict__ fend,
                                     const uint32_t* __restrict__ sbegin, const uint32_t* __restrict__ send, uint64_t n,
                                     DBoundary* __restrict__ bounds) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (fbegin[i]) bounds[sbegin[i]].begin = i;  // (exclusive scans: the index of this boundary)
  if (fend[i]) bounds[send[i]].end = i;        // the boundary that ends here is the send[i]-th (ends never precede their begins)
}
// src/search.rs:238-258: the last boundary and every boundary longer than one byte is hard; ' - _ are weak; else normal
__global__ void strength_kernel(const uint8_t* __restrict__ text, DBoundary* __restrict__ bounds, uint64_t nb, uint32_t* __restrict__ closer) {
  const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (b >= nb) return;
  DBoundary d = bounds[b];
  const uint64_t len = d.end - d.begin;
  int s;
  if (b + 1 == nb || len > 1) {
    s = BOUNDARY_HARD;
  } else {
    const char c = len == 1 ? (char)text[d.begin] : 0;
    s = (c == '\'' || c == '-' || c == '_') ? BOUNDARY_WEAK : BOUNDARY_NORMAL;
  }
  bounds[b].strength = s;
  // a hard boundary closes the batch that runs up to it -- unless it sits at the very start of its batch, which only the
  // first boundary of a text that begins with one can do (src/lib.rs:1822, list_batches in search.cpp)
  closer[b] = (s == BOUNDARY_HARD && !(b == 0 && d.begin == 0)) ? 1u : 0u;
}
struct DBatch {
  uint64_t begin, begin_index, end_index;
};
__global__ void batch_kernel(const DBoundary* __restrict__ bounds, const uint32_t* __restrict__ closer, const uint32_t* __restrict__ cscan,
                             uint64_t nb, DBatch* __restrict__ batches) {
  const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (b >= nb || !closer[b]) return;
  const uint32_t k = cscan[b];  // exclusive: the batch this boundary closes
  batches[k].end_index = b;
  // the next batch starts behind this boundary
  batches[k + 1].begin = bounds[b].end;
  batches[k + 1].begin_index = b + 1;
}

__device__ __forceinline__ bool usable(const uint8_t* __restrict__ text, uint64_t b, uint64_t e) {
  return e > b && !(e - b == 1 && text[b] == ' ');
}
// find_match_ngrams (src/search.rs:262-312) for one batch and one order; EMIT = false counts, true writes
template <bool EMIT>
__device__ __forceinline__ uint32_t batch_ngrams(const uint8_t* __restrict__ text, const DBoundary* __restrict__ bounds, uint64_t nbounds,
                                                 uint32_t order, uint64_t begin, uint64_t end, SegmentSpan* __restrict__ out) {
  uint32_t c = 0;
  for (uint64_t i = 0; i + order - 1 < nbounds; ++i) {
    const uint64_t rb = bounds[i + order - 1].begin;
    if (rb > end) break;
    if (usable(text, begin, rb)) {
      if (EMIT) out[c] = SegmentSpan{(size_t)begin, (size_t)rb, order};
      ++c;
    }
    begin = bounds[i].end;
  }
  if (begin < end && usable(text, begin, end)) {
    // Match::internal_boundaries (src/search.rs:99-116) counts with a first/last window: the first
    // inner boundary only opens the window, every later one extends it.
    long long first = -1;
    uint64_t last_plus1 = 0;
    for (uint64_t k = 0; k < nbounds; ++k)
      if (bounds[k].begin > begin && bounds[k].end < end) {
        if (first < 0) first = (long long)k; else last_plus1 = k + 1;
      }
    const uint64_t inner = (first < 0 || (uint64_t)first >= last_plus1) ? 0 : last_plus1 - (uint64_t)first;
    if (inner == order) {
      if (EMIT) out[c] = SegmentSpan{(size_t)begin, (size_t)end, order};
      ++c;
    }
  }
  return c;
}
__global__ void seg_count_kernel(const uint8_t* __restrict__ text, const DBoundary* __restrict__ bounds, const DBatch* __restrict__ batches,
                                 uint64_t nbatch, uint32_t max_ngram, uint32_t* __restrict__ count) {
  const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (k >= nbatch) return;
  const DBatch d = batches[k];
  const DBoundary* bb = bounds + d.begin_index;
  const uint64_t nbounds = d.end_index + 1 - d.begin_index, end = bounds[d.end_index].begin;
  uint32_t c = 0;
  for (uint32_t order = 1; order <= max_ngram; ++order) c += batch_ngrams<false>(text, bb, nbounds, order, d.begin, end, nullptr);
  count[k] = c;
}
__global__ void seg_emit_kernel(const uint8_t* __restrict__ text, const DBoundary* __restrict__ bounds, const DBatch* __restrict__ batches,
                                uint64_t nbatch, uint32_t max_ngram, const uint64_t* __restrict__ first, SegmentSpan* __restrict__ segs) {
  const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (k >= nbatch) return;
  const DBatch d = batches[k];
  const DBoundary* bb = bounds + d.begin_index;
  const uint64_t nbounds = d.end_index + 1 - d.begin_index, end = bounds[d.end_index].begin;
  SegmentSpan* out = segs + first[k];
  for (uint32_t order = 1; order <= max_ngram; ++order) out += batch_ngrams<true>(text, bb, nbounds, order, d.begin, end, out);
}
__global__ void widen_kernel(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// Scratch comes from the device's stream-ordered memory pool (release threshold raised once per device, ensure_tables):
// a dozen cudaMalloc / cudaFree pairs per call cost more than the kernels; from the pool they are free after the first call.
struct Pool {
  std::vector<void*> ptrs;
  ~Pool() {
    for (void* p : ptrs) cudaFreeAsync(p, 0);
  }
  template <class T>
  bool alloc(T** out, size_t count, std::string* err) {
    void* p = nullptr;
    if (cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), 0) != cudaSuccess) {
      cudaGetLastError();
      *err = "device segmentation: out of device memory";
      return false;
    }
    ptrs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return true;
  }
};
template <class T>
bool exclusive_scan(Pool& pool, const T* in, T* out, uint64_t n, std::string* err) {
  size_t bytes = 0;
  GS_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n));
  uint8_t* tmp = nullptr;
  if (!pool.alloc(&tmp, bytes, err)) return false;
  GS_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n));
  return true;
}
inline unsigned blocks_for(uint64_t n) { return (unsigned)std::max<uint64_t>(1, (n + 255) / 256); }

std::mutex g_tab_m;
bool g_tab_loaded[64] = {false};
bool ensure_tables(int device, std::string* err) {
  std::lock_guard<std::mutex> lk(g_tab_m);
  if (device >= 0 && device < 64 && g_tab_loaded[device]) return true;
  std::vector<uint32_t> alpha(2 * (size_t)anl_unicode::kAlphabeticRanges_len);
  for (unsigned i = 0; i < anl_unicode::kAlphabeticRanges_len; ++i) {
    alpha[2 * i] = anl_unicode::kAlphabeticRanges[i][0];
    alpha[2 * i + 1] = anl_unicode::kAlphabeticRanges[i][1];
  }
  const uint32_t n = anl_unicode::kAlphabeticRanges_len;
  if (n > 800) {
    *err = "alphabetic range table too large for the device";
    return false;
  }
  {
    cudaMemPool_t mp = nullptr;
    uint64_t keep = 8ull << 30;  // freed scratch stays with the pool up to this much instead of going back to the driver
    if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess && mp) cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
    cudaGetLastError();
  }
  GS_TRY(cudaMemcpyToSymbol(c_seg_alpha, alpha.data(), alpha.size() * sizeof(uint32_t)));
  GS_TRY(cudaMemcpyToSymbol(c_seg_n_alpha, &n, sizeof n));
  if (device >= 0 && device < 64) g_tab_loaded[device] = true;
  return true;
}

}  // namespace

// segment_text (search.cpp) on `device`; the same SegmentedText.  Texts of 2 GiB and more are left to the host.
bool segment_text_device(int device, const std::string& text, uint32_t max_ngram, SegmentedText* stp, std::string* err,
                         PodBuffer<Boundary>* bounds_out, PodBuffer<BatchDesc>* batches_out) {
  static_assert(sizeof(DBoundary) == sizeof(Boundary) && sizeof(DBatch) == sizeof(BatchDesc), "device records mirror the host's");
  SegmentedText& st = *stp;
  if (bounds_out) bounds_out->clear();
  if (batches_out) batches_out->clear();
  st.segs.clear();
  st.batch_first.resize(1);
  st.batch_first[0] = 0;
  const uint64_t n = text.size();
  if (n == 0) return true;
  if (n >= 0x7FFFFFF0ull) {
    *err = "text too large for the device segmentation";
    return false;
  }
  GS_TRY(cudaSetDevice(device));
  if (!ensure_tables(device, err)) return false;
  PhaseTimer pt;
  Pool pool;
  uint8_t *d_text = nullptr, *d_cls = nullptr;
  uint32_t *fbegin = nullptr, *fend = nullptr, *sbegin = nullptr, *send = nullptr;
  if (!pool.alloc(&d_text, n, err) || !pool.alloc(&d_cls, n, err) || !pool.alloc(&fbegin, n + 1, err) || !pool.alloc(&fend, n + 1, err) ||
      !pool.alloc(&sbegin, n + 1, err) || !pool.alloc(&send, n + 1, err))
    return false;
  GS_TRY(cudaMemcpy(d_text, text.data(), n, cudaMemcpyHostToDevice));
  class_kernel<<<blocks_for(n), 256>>>(d_text, n, d_cls);
  flag_kernel<<<blocks_for(n), 256>>>(d_cls, n, fbegin, fend);
  GS_TRY(cudaMemset(fbegin + n, 0, 4));
  GS_TRY(cudaMemset(fend + n, 0, 4));
  if (!exclusive_scan(pool, fbegin, sbegin, n + 1, err) || !exclusive_scan(pool, fend, send, n + 1, err)) return false;
  uint32_t nbegin = 0, nend = 0;
  GS_TRY(cudaMemcpy(&nbegin, sbegin + n, 4, cudaMemcpyDeviceToHost));
  GS_TRY(cudaMemcpy(&nend, send + n, 4, cudaMemcpyDeviceToHost));
  // the text always ends with a boundary: the open non-alphabetic run at its end, or one of length zero
  const bool open_tail = nbegin > nend;
  const uint64_t nb = (uint64_t)nbegin + (open_tail ? 0 : 1);
  DBoundary* d_bounds = nullptr;
  uint32_t *closer = nullptr, *cscan = nullptr;
  if (!pool.alloc(&d_bounds, nb, err) || !pool.alloc(&closer, nb + 1, err) || !pool.alloc(&cscan, nb + 1, err)) return false;
  boundary_fill_kernel<<<blocks_for(n), 256>>>(fbegin, fend, sbegin, send, n, d_bounds);
  {
    // the last boundary's end (open run) or the whole zero-length boundary at the end of the text
    DBoundary last;
    memset(&last, 0, sizeof last);
    if (open_tail) {
      GS_TRY(cudaMemcpy(&last, d_bounds + nb - 1, sizeof last, cudaMemcpyDeviceToHost));
      last.end = n;
    } else {
      last.begin = last.end = n;
    }
    GS_TRY(cudaMemcpy(d_bounds + nb - 1, &last, sizeof last, cudaMemcpyHostToDevice));
  }
  strength_kernel<<<blocks_for(nb), 256>>>(d_text, d_bounds, nb, closer);
  GS_TRY(cudaMemset(closer + nb, 0, 4));
  if (!exclusive_scan(pool, closer, cscan, nb + 1, err)) return false;
  uint32_t nbatch = 0;
  GS_TRY(cudaMemcpy(&nbatch, cscan + nb, 4, cudaMemcpyDeviceToHost));
  count_launch(6);
  if (bounds_out) {
    bounds_out->resize(nb);
    GS_TRY(cudaMemcpy(bounds_out->data(), d_bounds, nb * sizeof(DBoundary), cudaMemcpyDeviceToHost));
  }
  if (nbatch == 0) return true;
  DBatch* d_batches = nullptr;
  uint32_t* scount = nullptr;
  uint64_t *scount64 = nullptr, *sfirst = nullptr;
  if (!pool.alloc(&d_batches, (size_t)nbatch + 1, err) || !pool.alloc(&scount, nbatch, err) || !pool.alloc(&scount64, (size_t)nbatch + 1, err) ||
      !pool.alloc(&sfirst, (size_t)nbatch + 1, err))
    return false;
  {
    DBatch first;
    first.begin = 0;
    first.begin_index = 0;
    first.end_index = 0;
    GS_TRY(cudaMemcpy(d_batches, &first, sizeof first, cudaMemcpyHostToDevice));
  }
  batch_kernel<<<blocks_for(nb), 256>>>(d_bounds, closer, cscan, nb, d_batches);
  seg_count_kernel<<<blocks_for(nbatch), 256>>>(d_text, d_bounds, d_batches, nbatch, max_ngram, scount);
  widen_kernel<<<blocks_for(nbatch), 256>>>(scount, nbatch, scount64);
  GS_TRY(cudaMemset(scount64 + nbatch, 0, 8));
  if (!exclusive_scan(pool, scount64, sfirst, (uint64_t)nbatch + 1, err)) return false;
  uint64_t nsegs = 0;
  GS_TRY(cudaMemcpy(&nsegs, sfirst + nbatch, 8, cudaMemcpyDeviceToHost));
  SegmentSpan* d_segs = nullptr;
  if (!pool.alloc(&d_segs, nsegs, err)) return false;
  seg_emit_kernel<<<blocks_for(nbatch), 256>>>(d_text, d_bounds, d_batches, nbatch, max_ngram, sfirst, d_segs);
  count_launch(4);
  GS_TRY(cudaGetLastError());
  pt.lap("device segmentation: kernels");
  st.segs.resize(nsegs);
  st.batch_first.resize((size_t)nbatch + 1);
  if (nsegs) GS_TRY(cudaMemcpy(st.segs.data(), d_segs, nsegs * sizeof(SegmentSpan), cudaMemcpyDeviceToHost));
  GS_TRY(cudaMemcpy(st.batch_first.data(), sfirst, ((size_t)nbatch + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (batches_out) {
    batches_out->resize(nbatch);
    GS_TRY(cudaMemcpy(batches_out->data(), d_batches, (size_t)nbatch * sizeof(DBatch), cudaMemcpyDeviceToHost));
  }
  pt.lap("device segmentation: download");
  return true;
}

}  // namespace anl
