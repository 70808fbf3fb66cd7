// gpu_build.cu -- the index build (src/lib.rs:192-245) on the device.
//
// HostModel::build_index does the same work on the host cores (anagram values of every indexed entry, instances in
// (key, id) order, anagram arrays, symmetric-delete postings in (fingerprint, anagram, class) order, Bloom words,
// open-addressing table).  For the 10 M-entry lexicon of BASELINE config 5 that is tens of seconds; here every phase
// is a kernel or a device-wide primitive:
//   key_kernel        192-bit prime-product key per entry (overflow = build error, as on the host)
//   3 x radix sort    instances by (key, id): stable LSD passes over the three 64-bit limbs (entries arrive in id order)
//   boundary + scan   anagram ranks; anagram arrays and instance rows written in gather order
//   posting kernels   1 + distinct classes postings per anagram (count, scan, generate in (anagram, class) order)
//   radix sort        postings by fingerprint (stable: equal fingerprints stay in (anagram, class) order)
//   group kernels     distinct fingerprints -> Bloom words (atomic OR) and slot records
//   radix sort + scan slots by home position; linear probing becomes a prefix maximum: pos_i = max(home_i, pos_{i-1} + 1)
// The sorts and scans are CUB's device-wide primitives (the CUDA toolkit's own headers) -- a one-off build step, not
// the lookup path.  The result is downloaded into the same HostIndex the host build fills, so persistence, has() and
// the upload path are shared.  Every array equals the host build's bit for bit except `table`: the host inserts keys in
// fingerprint order, this build in home-slot order; both are valid linear-probing layouts of the same keys (checked by
// tests/test_gpu_index_build.py, which also compares lookups).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstring>

#include "host_model.h"
#include "kernel_common.cuh"

namespace anl {

namespace {

#define GB_TRY(expr)                                                                       \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      *err = std::string("CUDA error in the GPU index build: ") + cudaGetErrorString(_e) + " at " #expr; \
      return false;                                                                        \
    }                                                                                      \
  } while (0)

// device buffers freed on scope exit
struct DevPool {
  std::vector<void*> ptrs;
  ~DevPool() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <class T>
  bool alloc(T** out, size_t count, std::string* err) {
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) {
      *err = std::string("GPU index build: out of device memory (") + cudaGetErrorString(e) + ")";
      return false;
    }
    ptrs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return true;
  }
  void release(void* p) {
    for (auto& q : ptrs)
      if (q == p) {
        cudaFree(q);
        q = nullptr;
      }
  }
};

__device__ __forceinline__ bool key_mul_dev(Key192& k, uint64_t m) {
  const uint64_t l0 = k.w0 * m, h0 = __umul64hi(k.w0, m);
  const uint64_t l1 = k.w1 * m, h1 = __umul64hi(k.w1, m);
  const uint64_t l2 = k.w2 * m, h2 = __umul64hi(k.w2, m);
  const uint64_t r1 = l1 + h0;
  const uint64_t c1 = r1 < l1;
  uint64_t r2 = l2 + h1;
  uint64_t c2 = r2 < l2;
  r2 += c1;
  c2 += (r2 < c1);
  k.w0 = l0;
  k.w1 = r1;
  k.w2 = r2;
  return (h2 + c2) == 0;
}
__device__ __forceinline__ uint32_t key_bits_dev(const Key192& k) {
  if (k.w2) return 128 + 64 - __clzll((long long)k.w2);
  if (k.w1) return 64 + 64 - __clzll((long long)k.w1);
  if (k.w0) return 64 - __clzll((long long)k.w0);
  return 0;
}

struct BuildStats {
  unsigned int first_overflow;   // smallest entry index whose key exceeds 192 bits (0xFFFFFFFF = none)
  unsigned int max_key_bits, max_len, max_charcount;
  unsigned long long charcount_mask[4];
  unsigned int class_seen[8];    // bit s: symbol s occurs
  unsigned int long_list;        // a posting list longer than 65535
  unsigned int overflow_slots;   // keys placed beyond the end of the table (wrapped by the serial tail pass)
};

// rows: [n][stride] = len, flags, symbols (entries in id order)
__global__ void key_kernel(const uint8_t* __restrict__ rows, uint32_t stride, uint32_t n, const uint32_t* __restrict__ prime_of,
                           uint64_t* __restrict__ w0, uint64_t* __restrict__ w1, uint64_t* __restrict__ w2, BuildStats* st) {
  __shared__ uint32_t s_prime[256];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_prime[i] = prime_of[i];
  __syncthreads();
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint8_t* row = rows + (size_t)k * stride;
  const uint32_t len = row[0];
  Key192 key{1, 0, 0};
  bool ok = true;
  uint64_t pp = 1;
  for (uint32_t i = 0; i < len && ok; ++i) {
    pp *= s_prime[row[2 + i]];
    if (pp >> 53) {
      ok = key_mul_dev(key, pp);
      pp = 1;
    }
  }
  if (ok && pp > 1) ok = key_mul_dev(key, pp);
  if (!ok) atomicMin(&st->first_overflow, k);
  w0[k] = key.w0;
  w1[k] = key.w1;
  w2[k] = key.w2;
  atomicMax(&st->max_key_bits, ok ? key_bits_dev(key) : 0u);
}

__global__ void iota_kernel(uint32_t* v, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n, uint64_t* __restrict__ dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

// lexicon-sharded build: keep the instances whose key hashes to this shard
__global__ void shard_flag_kernel(const uint64_t* __restrict__ w0, const uint64_t* __restrict__ w1, const uint64_t* __restrict__ w2,
                                  const uint32_t* __restrict__ perm, uint32_t n, uint32_t shard, uint32_t n_shards, uint32_t* __restrict__ flag) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const uint32_t k = perm[g];
  flag[g] = hash_key(w0[k], w1[k], w2[k]) % n_shards == shard ? 1u : 0u;
}
__global__ void shard_compact_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                                     uint32_t n, uint32_t* __restrict__ perm_out, uint32_t* __restrict__ gid_out) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n || !flag[g]) return;
  perm_out[pos[g]] = perm[g];
  gid_out[pos[g]] = g;  // position in the global (key, id) order: the tie-break key on every shard
}

// does instance g (in gather order) start a new anagram?
__global__ void boundary_kernel(const uint64_t* __restrict__ w0, const uint64_t* __restrict__ w1, const uint64_t* __restrict__ w2,
                                const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ flag) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  bool start = g == 0;
  if (!start) {
    const uint32_t a = perm[g], b = perm[g - 1];
    start = w0[a] != w0[b] || w1[a] != w1[b] || w2[a] != w2[b];
  }
  flag[g] = start ? 1u : 0u;
}

// instance arrays in gather order + the anagram arrays; rank_incl = inclusive scan of the boundary flags
__global__ void instance_kernel(const uint8_t* __restrict__ rows, uint32_t stride, const uint32_t* __restrict__ ids,
                                const uint32_t* __restrict__ freqs, const uint64_t* __restrict__ w0, const uint64_t* __restrict__ w1,
                                const uint64_t* __restrict__ w2, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ flag,
                                const uint32_t* __restrict__ rank_incl, uint32_t n, uint32_t out_stride, uint8_t* __restrict__ inst_rows,
                                uint32_t* __restrict__ inst_vocab, uint32_t* __restrict__ inst_freq, Key192* __restrict__ ana_key,
                                uint32_t* __restrict__ ana_inst_off, uint16_t* __restrict__ ana_charcount, BuildStats* st) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const uint32_t k = perm[g];
  const uint8_t* row = rows + (size_t)k * stride;
  uint8_t* out = inst_rows + (size_t)g * out_stride;
  const uint32_t len = row[0];
  for (uint32_t i = 0; i < out_stride; ++i) out[i] = i < len + 2 ? row[i] : 0;
  inst_vocab[g] = ids[k];
  inst_freq[g] = freqs[k];
  atomicMax(&st->max_len, len);
  for (uint32_t i = 0; i < len; ++i) atomicOr(&st->class_seen[row[2 + i] >> 5], 1u << (row[2 + i] & 31));
  if (flag[g]) {
    const uint32_t r = rank_incl[g] - 1;
    ana_key[r] = Key192{w0[k], w1[k], w2[k]};
    ana_inst_off[r] = g;
    ana_charcount[r] = (uint16_t)len;
    atomicMax(&st->max_charcount, len);
    atomicOr(&st->charcount_mask[len >> 6], 1ull << (len & 63));
  }
}

// distinct classes of an anagram (from its first instance), ascending, into a 256-bit set
__device__ __forceinline__ void class_set(const uint8_t* row, uint32_t* bits /*[8]*/) {
  for (int i = 0; i < 8; ++i) bits[i] = 0;
  const uint32_t len = row[0];
  for (uint32_t i = 0; i < len; ++i) bits[row[2 + i] >> 5] |= 1u << (row[2 + i] & 31);
}
__global__ void posting_count_kernel(const uint8_t* __restrict__ inst_rows, uint32_t stride, const uint32_t* __restrict__ ana_inst_off,
                                     uint32_t nana, int sd, uint32_t* __restrict__ count) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nana) return;
  const uint8_t* row = inst_rows + (size_t)ana_inst_off[r] * stride;
  uint32_t c = 1;  // the self posting
  if (sd >= 1 && row[0] > 1) {  // (the empty value is never a node: no empty leaves, src/iterators.rs:177)
    uint32_t bits[8];
    class_set(row, bits);
    for (int i = 0; i < 8; ++i) c += __popc(bits[i]);
  }
  count[r] = c;
}
// postings of anagram r at [first[r], first[r] + count): one per distinct class, ascending, then the self posting --
// the order (anagram, class as a byte, POST_SELF = 0xFF last) that a stable sort by fingerprint must preserve
__global__ void posting_gen_kernel(const uint8_t* __restrict__ inst_rows, uint32_t stride, const uint32_t* __restrict__ ana_inst_off,
                                   uint32_t nana, int sd, const uint32_t* __restrict__ first, uint64_t* __restrict__ fp,
                                   uint32_t* __restrict__ ana, uint8_t* __restrict__ cls) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nana) return;
  const uint8_t* row = inst_rows + (size_t)ana_inst_off[r] * stride;
  const uint32_t len = row[0];
  uint64_t h = 0;
  for (uint32_t i = 0; i < len; ++i) h += class_rnd(row[2 + i]);
  uint32_t p = first[r];
  if (sd >= 1 && len > 1) {
    uint32_t bits[8];
    class_set(row, bits);
    for (uint32_t w = 0; w < 8; ++w) {
      uint32_t m = bits[w];
      while (m) {
        const uint32_t x = w * 32 + (uint32_t)__ffs(m) - 1;
        m &= m - 1;
        fp[p] = h - class_rnd(x);
        ana[p] = r;
        cls[p] = (uint8_t)x;
        ++p;
      }
    }
  }
  fp[p] = h;
  ana[p] = r;
  cls[p] = POST_SELF;
}
__global__ void posting_gather_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ ana, const uint8_t* __restrict__ cls,
                                      uint32_t n, uint32_t* __restrict__ post_ana, uint8_t* __restrict__ post_cls) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  post_ana[t] = ana[order[t]];
  post_cls[t] = cls[order[t]];
}
__global__ void group_flag_kernel(const uint64_t* __restrict__ fp, uint32_t n, uint32_t* __restrict__ flag) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) flag[t] = (t == 0 || fp[t] != fp[t - 1]) ? 1u : 0u;
}
// one record per distinct fingerprint: {fp, first posting, home slot}; Bloom words on the way (OR is order-free)
__global__ void group_kernel(const uint64_t* __restrict__ fp, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ gid_incl,
                             uint32_t n, uint64_t slot_mask, uint64_t word_mask, uint64_t* __restrict__ g_fp, uint32_t* __restrict__ g_off,
                             uint32_t* __restrict__ g_home, unsigned long long* __restrict__ bloom) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n || !flag[t]) return;
  const uint32_t gi = gid_incl[t] - 1;
  const uint64_t f = fp[t];
  g_fp[gi] = f;
  g_off[gi] = t;
  g_home[gi] = (uint32_t)fp_index(f, slot_mask);
  atomicOr(bloom + fp_index(f, word_mask), (unsigned long long)bloom_mask(f));
}
// linear probing in home-slot order is a prefix maximum: pos_i = max(home_i, pos_{i-1} + 1) = i + max_{j<=i}(home_j - j)
__global__ void slack_kernel(const uint32_t* __restrict__ home_sorted, uint32_t n, long long* __restrict__ v) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (long long)home_sorted[i] - (long long)i;
}
struct MaxOp {
  __device__ __forceinline__ long long operator()(long long a, long long b) const { return a > b ? a : b; }
};
__global__ void place_kernel(const uint32_t* __restrict__ order, const long long* __restrict__ vmax, const uint64_t* __restrict__ g_fp,
                             const uint32_t* __restrict__ g_off, uint32_t ngroups, uint32_t nposts, uint64_t slots, Slot* __restrict__ table,
                             BuildStats* st) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ngroups) return;
  const uint32_t gi = order[i];
  const uint32_t end = gi + 1 < ngroups ? g_off[gi + 1] : nposts;
  const uint32_t cnt = end - g_off[gi];
  if (cnt > 0xFFFFu) atomicOr(&st->long_list, 1u);
  const uint64_t pos = (uint64_t)((long long)i + vmax[i]);
  if (pos < slots) {
    table[pos] = Slot{g_fp[gi], g_off[gi], (uint16_t)cnt, 0};
  } else {
    atomicAdd(&st->overflow_slots, 1u);  // runs past the end of the table: the serial pass below wraps it
  }
}
// the few keys whose run reaches past the last slot, in order, into the first free slots from 0 (what linear probing
// does when these keys are inserted last); one thread
__global__ void wrap_kernel(const uint32_t* __restrict__ order, const long long* __restrict__ vmax, const uint64_t* __restrict__ g_fp,
                            const uint32_t* __restrict__ g_off, uint32_t ngroups, uint32_t nposts, uint64_t slots, Slot* __restrict__ table) {
  if (blockIdx.x || threadIdx.x) return;
  uint64_t next = 0;
  // positions grow with i: the overflowing keys are a suffix
  uint32_t lo = ngroups;
  while (lo > 0 && (uint64_t)((long long)(lo - 1) + vmax[lo - 1]) >= slots) --lo;
  for (uint32_t i = lo; i < ngroups; ++i) {
    const uint32_t gi = order[i];
    const uint32_t end = gi + 1 < ngroups ? g_off[gi + 1] : nposts;
    while (table[next].post_cnt != 0) ++next;
    table[next] = Slot{g_fp[gi], g_off[gi], (uint16_t)(end - g_off[gi]), 0};
  }
}

template <class K, class V>
bool radix_sort_pairs(DevPool& pool, K* keys_in, K* keys_out, V* vals_in, V* vals_out, uint32_t n, int end_bit, std::string* err) {
  size_t bytes = 0;
  GB_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit));
  uint8_t* tmp = nullptr;
  if (!pool.alloc(&tmp, bytes, err)) return false;
  GB_TRY(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit));
  GB_TRY(cudaDeviceSynchronize());
  pool.release(tmp);
  return true;
}
bool inclusive_sum(DevPool& pool, const uint32_t* in, uint32_t* out, uint32_t n, std::string* err) {
  size_t bytes = 0;
  GB_TRY(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, (int)n));
  uint8_t* tmp = nullptr;
  if (!pool.alloc(&tmp, bytes, err)) return false;
  GB_TRY(cub::DeviceScan::InclusiveSum(tmp, bytes, in, out, (int)n));
  GB_TRY(cudaDeviceSynchronize());
  pool.release(tmp);
  return true;
}
bool exclusive_sum(DevPool& pool, const uint32_t* in, uint32_t* out, uint32_t n, std::string* err) {
  size_t bytes = 0;
  GB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n));
  uint8_t* tmp = nullptr;
  if (!pool.alloc(&tmp, bytes, err)) return false;
  GB_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n));
  GB_TRY(cudaDeviceSynchronize());
  pool.release(tmp);
  return true;
}
template <class T, class Vec>
bool download(const T* d, size_t n, Vec* v, std::string* err) {
  v->resize(n);
  if (n) GB_TRY(cudaMemcpy(v->data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
  return true;
}
inline unsigned blocks_for(uint64_t n, unsigned threads = 256) { return (unsigned)std::max<uint64_t>(1, (n + threads - 1) / threads); }

}  // namespace

// The device build of HostModel::build_index: same checks, same arrays (see the header comment for `table`).
bool gpu_build_index(HostModel* hm, int sd, uint32_t shard, uint32_t n_shards, int device, std::string* err) {
  PhaseTimer pt;
  HostIndex ix;
  ix.sd = sd;
  if (n_shards == 0 || shard >= n_shards) {
    *err = "invalid shard";
    return false;
  }
  ix.shard = shard;
  ix.n_shards = n_shards;
  if (!hm->check_variant_support(n_shards, err)) return false;
  if (hm->alphabet.size() + 1 > 168) {
    *err = "alphabet has more classes than there are primes (168)";
    return false;
  }
  for (uint32_t s = 0; s <= hm->alphabet.size(); ++s) ix.prime_of[s] = kPrimes[s];
  if (hm->decoder.size() > 0xFFFFFFF0ull) {
    *err = "vocabulary too large";
    return false;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    *err = "no CUDA device available; the variant-lookup path has no CPU fallback";
    return false;
  }
  if (device >= 0) GB_TRY(cudaSetDevice(device));

  // ---- host: the indexed entries as fixed-stride rows, in id order -------------------------------------------------
  const std::vector<VocabEntry>& dec = hm->decoder;
  std::vector<uint32_t> ids;
  ids.reserve(dec.size());
  uint32_t max_len_all = 0;
  for (size_t id = 0; id < dec.size(); ++id) {
    const VocabEntry& v = dec[id];
    if (!(v.vocabtype & VT_INDEXED) || v.syms.empty()) continue;
    if (v.syms.size() > (size_t)ANL_MAX_SYMBOLS) {
      *err = "lexicon entry longer than " + std::to_string(ANL_MAX_SYMBOLS) + " symbols: " + v.text;
      return false;
    }
    ids.push_back((uint32_t)id);
    max_len_all = std::max<uint32_t>(max_len_all, (uint32_t)v.syms.size());
  }
  const uint32_t n_all = (uint32_t)ids.size();
  if (n_all == 0) {
    *err = "no indexed vocabulary entries";
    return false;
  }
  const uint32_t in_stride = ((max_len_all + 2) + 15) & ~15u;
  RawVec<uint8_t> h_rows((size_t)n_all * in_stride);
  std::vector<uint32_t> h_freq(n_all);
  parallel_ranges(n_all, 1u << 14, [&](unsigned, uint64_t lo, uint64_t hi) {
    for (uint64_t k = lo; k < hi; ++k) {
      const VocabEntry& v = dec[ids[k]];
      uint8_t* row = h_rows.data() + k * in_stride;
      memset(row, 0, in_stride);
      row[0] = (uint8_t)v.syms.size();
      row[1] = v.first_lower ? ROW_FIRST_LOWER : 0;
      memcpy(row + 2, v.syms.data(), v.syms.size());
      h_freq[k] = v.frequency;
    }
  });
  pt.lap("gpu build: pack rows");

  DevPool pool;
  uint8_t* d_rows = nullptr;
  uint32_t *d_ids = nullptr, *d_freq = nullptr, *d_prime = nullptr;
  BuildStats* d_st = nullptr;
  if (!pool.alloc(&d_rows, h_rows.size(), err) || !pool.alloc(&d_ids, n_all, err) || !pool.alloc(&d_freq, n_all, err) ||
      !pool.alloc(&d_prime, 256, err) || !pool.alloc(&d_st, 1, err))
    return false;
  GB_TRY(cudaMemcpy(d_rows, h_rows.data(), h_rows.size(), cudaMemcpyHostToDevice));
  GB_TRY(cudaMemcpy(d_ids, ids.data(), (size_t)n_all * 4, cudaMemcpyHostToDevice));
  GB_TRY(cudaMemcpy(d_freq, h_freq.data(), (size_t)n_all * 4, cudaMemcpyHostToDevice));
  GB_TRY(cudaMemcpy(d_prime, ix.prime_of, sizeof ix.prime_of, cudaMemcpyHostToDevice));
  BuildStats st0;
  memset(&st0, 0, sizeof st0);
  st0.first_overflow = 0xFFFFFFFFu;
  GB_TRY(cudaMemcpy(d_st, &st0, sizeof st0, cudaMemcpyHostToDevice));

  // ---- keys ----------------------------------------------------------------------------------------------------------
  uint64_t *w0 = nullptr, *w1 = nullptr, *w2 = nullptr;
  if (!pool.alloc(&w0, n_all, err) || !pool.alloc(&w1, n_all, err) || !pool.alloc(&w2, n_all, err)) return false;
  key_kernel<<<blocks_for(n_all), 256>>>(d_rows, in_stride, n_all, d_prime, w0, w1, w2, d_st);
  count_launch(1);
  BuildStats st;
  GB_TRY(cudaMemcpy(&st, d_st, sizeof st, cudaMemcpyDeviceToHost));
  if (st.first_overflow != 0xFFFFFFFFu) {
    *err = "anagram value of lexicon entry exceeds 192 bits: " + dec[ids[st.first_overflow]].text;
    return false;
  }
  pt.lap("gpu build: keys");

  // ---- instances in (key, id) order: stable LSD radix passes over the limbs that hold bits ----------------------------
  uint32_t *perm = nullptr, *perm2 = nullptr;
  uint64_t *kin = nullptr, *kout = nullptr;
  if (!pool.alloc(&perm, n_all, err) || !pool.alloc(&perm2, n_all, err) || !pool.alloc(&kin, n_all, err) || !pool.alloc(&kout, n_all, err))
    return false;
  iota_kernel<<<blocks_for(n_all), 256>>>(perm, n_all);
  const int limbs = st.max_key_bits > 128 ? 3 : (st.max_key_bits > 64 ? 2 : 1);
  for (int l = 0; l < limbs; ++l) {
    const uint64_t* src = l == 0 ? w0 : (l == 1 ? w1 : w2);
    gather_u64_kernel<<<blocks_for(n_all), 256>>>(src, perm, n_all, kin);
    const int bits = l + 1 < limbs ? 64 : (int)st.max_key_bits - 64 * l;
    if (!radix_sort_pairs(pool, kin, kout, perm, perm2, n_all, bits, err)) return false;
    std::swap(perm, perm2);
  }
  count_launch(1 + limbs);
  pt.lap("gpu build: sort instances");

  // ---- lexicon shard: keep this shard's anagrams, remember the global positions -----------------------------------------
  uint32_t n = n_all;
  uint32_t* d_gid = nullptr;
  if (n_shards > 1) {
    uint32_t *flag = nullptr, *pos = nullptr, *perm_s = nullptr;
    if (!pool.alloc(&flag, n_all, err) || !pool.alloc(&pos, n_all, err) || !pool.alloc(&perm_s, n_all, err) || !pool.alloc(&d_gid, n_all, err))
      return false;
    shard_flag_kernel<<<blocks_for(n_all), 256>>>(w0, w1, w2, perm, n_all, shard, n_shards, flag);
    if (!exclusive_sum(pool, flag, pos, n_all, err)) return false;
    uint32_t last_pos = 0, last_flag = 0;
    GB_TRY(cudaMemcpy(&last_pos, pos + n_all - 1, 4, cudaMemcpyDeviceToHost));
    GB_TRY(cudaMemcpy(&last_flag, flag + n_all - 1, 4, cudaMemcpyDeviceToHost));
    n = last_pos + last_flag;
    if (n == 0) {
      *err = "shard holds no anagrams";
      return false;
    }
    shard_compact_kernel<<<blocks_for(n_all), 256>>>(perm, flag, pos, n_all, perm_s, d_gid);
    count_launch(2);
    perm = perm_s;
  }

  // ---- anagram boundaries, instance arrays in gather order ------------------------------------------------------------------
  uint32_t *bflag = nullptr, *rank_incl = nullptr;
  if (!pool.alloc(&bflag, n, err) || !pool.alloc(&rank_incl, n, err)) return false;
  boundary_kernel<<<blocks_for(n), 256>>>(w0, w1, w2, perm, n, bflag);
  if (!inclusive_sum(pool, bflag, rank_incl, n, err)) return false;
  uint32_t nana = 0;
  GB_TRY(cudaMemcpy(&nana, rank_incl + n - 1, 4, cudaMemcpyDeviceToHost));
  // (the shard's own longest entry sizes its rows, like the host build)
  uint32_t shard_max_len = max_len_all;
  if (n_shards > 1) {
    // max_len of the shard is only known after the pass below; a first pass over the lengths is cheap on the host side:
    // rows keep the global stride when it already is the smallest multiple of 16 that fits (the common case)
    std::vector<uint32_t> h_perm(n);
    GB_TRY(cudaMemcpy(h_perm.data(), perm, (size_t)n * 4, cudaMemcpyDeviceToHost));
    shard_max_len = 0;
    for (uint32_t g = 0; g < n; ++g) shard_max_len = std::max<uint32_t>(shard_max_len, h_rows[(size_t)h_perm[g] * in_stride]);
  }
  ix.norm_stride = ((shard_max_len + 2) + 15) & ~15u;
  uint8_t* d_inst_rows = nullptr;
  uint32_t *d_inst_vocab = nullptr, *d_inst_freq = nullptr, *d_ana_off = nullptr;
  Key192* d_ana_key = nullptr;
  uint16_t* d_ana_cc = nullptr;
  if (!pool.alloc(&d_inst_rows, (size_t)n * ix.norm_stride, err) || !pool.alloc(&d_inst_vocab, n, err) || !pool.alloc(&d_inst_freq, n, err) ||
      !pool.alloc(&d_ana_key, nana, err) || !pool.alloc(&d_ana_off, (size_t)nana + 1, err) || !pool.alloc(&d_ana_cc, nana, err))
    return false;
  instance_kernel<<<blocks_for(n), 256>>>(d_rows, in_stride, d_ids, d_freq, w0, w1, w2, perm, bflag, rank_incl, n, ix.norm_stride,
                                          d_inst_rows, d_inst_vocab, d_inst_freq, d_ana_key, d_ana_off, d_ana_cc, d_st);
  GB_TRY(cudaMemcpy(d_ana_off + nana, &n, 4, cudaMemcpyHostToDevice));
  count_launch(2);
  pt.lap("gpu build: instance + anagram arrays");

  // ---- postings: count, scan, generate, sort by fingerprint -------------------------------------------------------------------
  uint32_t *pcount = nullptr, *pfirst = nullptr;
  if (!pool.alloc(&pcount, nana, err) || !pool.alloc(&pfirst, nana, err)) return false;
  posting_count_kernel<<<blocks_for(nana), 256>>>(d_inst_rows, ix.norm_stride, d_ana_off, nana, sd, pcount);
  if (!exclusive_sum(pool, pcount, pfirst, nana, err)) return false;
  uint32_t lastc = 0, lastf = 0;
  GB_TRY(cudaMemcpy(&lastc, pcount + nana - 1, 4, cudaMemcpyDeviceToHost));
  GB_TRY(cudaMemcpy(&lastf, pfirst + nana - 1, 4, cudaMemcpyDeviceToHost));
  const uint64_t nposts64 = (uint64_t)lastc + lastf;
  if (nposts64 > 0x7FFFFFF0ull) {
    *err = "too many postings for the GPU index build";
    return false;
  }
  const uint32_t nposts = (uint32_t)nposts64;
  uint64_t *fp_in = nullptr, *fp_sorted = nullptr;
  uint32_t *p_ana = nullptr, *order_in = nullptr, *order = nullptr;
  uint8_t* p_cls = nullptr;
  if (!pool.alloc(&fp_in, nposts, err) || !pool.alloc(&fp_sorted, nposts, err) || !pool.alloc(&p_ana, nposts, err) ||
      !pool.alloc(&p_cls, nposts, err) || !pool.alloc(&order_in, nposts, err) || !pool.alloc(&order, nposts, err))
    return false;
  posting_gen_kernel<<<blocks_for(nana), 256>>>(d_inst_rows, ix.norm_stride, d_ana_off, nana, sd, pfirst, fp_in, p_ana, p_cls);
  iota_kernel<<<blocks_for(nposts), 256>>>(order_in, nposts);
  if (!radix_sort_pairs(pool, fp_in, fp_sorted, order_in, order, nposts, 64, err)) return false;
  uint32_t* d_post_ana = nullptr;
  uint8_t* d_post_cls = nullptr;
  if (!pool.alloc(&d_post_ana, nposts, err) || !pool.alloc(&d_post_cls, nposts, err)) return false;
  posting_gather_kernel<<<blocks_for(nposts), 256>>>(order, p_ana, p_cls, nposts, d_post_ana, d_post_cls);
  count_launch(5);
  pt.lap("gpu build: postings");

  // ---- distinct fingerprints: Bloom words, slot records, table ------------------------------------------------------------------
  uint32_t *gflag = nullptr, *gid_incl = nullptr;
  if (!pool.alloc(&gflag, nposts, err) || !pool.alloc(&gid_incl, nposts, err)) return false;
  group_flag_kernel<<<blocks_for(nposts), 256>>>(fp_sorted, nposts, gflag);
  if (!inclusive_sum(pool, gflag, gid_incl, nposts, err)) return false;
  uint32_t ngroups = 0;
  GB_TRY(cudaMemcpy(&ngroups, gid_incl + nposts - 1, 4, cudaMemcpyDeviceToHost));
  ix.table_keys = ngroups;
  uint64_t slots = 1;
  while (slots < (uint64_t)ngroups * 2) slots <<= 1;
  uint64_t words = 1;
  while (words * 2 < ngroups) words <<= 1;  // (the sizing rules of the host build)
  uint64_t bloom_max_bytes = 128ull << 20;
  if (const char* e = getenv("ANL_BLOOM_MAX_MB")) bloom_max_bytes = (uint64_t)std::max(1, atoi(e)) << 20;
  while (words * 8 > bloom_max_bytes && words * 8 >= ngroups) words >>= 1;
  Slot* d_table = nullptr;
  unsigned long long* d_bloom = nullptr;
  uint64_t* g_fp = nullptr;
  uint32_t *g_off = nullptr, *g_home = nullptr, *g_home_sorted = nullptr, *g_order_in = nullptr, *g_order = nullptr;
  long long *slack = nullptr, *slack_max = nullptr;
  if (!pool.alloc(&d_table, slots, err) || !pool.alloc(&d_bloom, words, err) || !pool.alloc(&g_fp, ngroups, err) ||
      !pool.alloc(&g_off, ngroups, err) || !pool.alloc(&g_home, ngroups, err) || !pool.alloc(&g_home_sorted, ngroups, err) ||
      !pool.alloc(&g_order_in, ngroups, err) || !pool.alloc(&g_order, ngroups, err) || !pool.alloc(&slack, ngroups, err) ||
      !pool.alloc(&slack_max, ngroups, err))
    return false;
  GB_TRY(cudaMemset(d_table, 0, slots * sizeof(Slot)));
  GB_TRY(cudaMemset(d_bloom, 0, words * sizeof(unsigned long long)));
  group_kernel<<<blocks_for(nposts), 256>>>(fp_sorted, gflag, gid_incl, nposts, slots - 1, words - 1, g_fp, g_off, g_home, d_bloom);
  iota_kernel<<<blocks_for(ngroups), 256>>>(g_order_in, ngroups);
  int home_bits = 1;
  while ((1ull << home_bits) < slots) ++home_bits;
  if (!radix_sort_pairs(pool, g_home, g_home_sorted, g_order_in, g_order, ngroups, home_bits, err)) return false;
  slack_kernel<<<blocks_for(ngroups), 256>>>(g_home_sorted, ngroups, slack);
  {
    size_t bytes = 0;
    GB_TRY(cub::DeviceScan::InclusiveScan(nullptr, bytes, slack, slack_max, MaxOp(), (int)ngroups));
    uint8_t* tmp = nullptr;
    if (!pool.alloc(&tmp, bytes, err)) return false;
    GB_TRY(cub::DeviceScan::InclusiveScan(tmp, bytes, slack, slack_max, MaxOp(), (int)ngroups));
    GB_TRY(cudaDeviceSynchronize());
    pool.release(tmp);
  }
  place_kernel<<<blocks_for(ngroups), 256>>>(g_order, slack_max, g_fp, g_off, ngroups, nposts, slots, d_table, d_st);
  wrap_kernel<<<1, 32>>>(g_order, slack_max, g_fp, g_off, ngroups, nposts, slots, d_table);
  count_launch(8);
  GB_TRY(cudaDeviceSynchronize());
  GB_TRY(cudaMemcpy(&st, d_st, sizeof st, cudaMemcpyDeviceToHost));
  if (st.long_list) {
    *err = "posting list too long";
    return false;
  }
  pt.lap("gpu build: Bloom filter + table");

  // ---- download into the host index ----------------------------------------------------------------------------------------------
  ix.max_len = st.max_len;
  ix.max_key_bits = 0;
  ix.max_charcount = st.max_charcount;
  for (int w = 0; w < 4; ++w) ix.charcount_mask[w] = st.charcount_mask[w];
  for (uint32_t s = 0; s < 256; ++s)
    if (st.class_seen[s >> 5] & (1u << (s & 31))) ix.active_classes.push_back((uint8_t)s);
  if (!download(d_ana_key, nana, &ix.ana_key, err) || !download(d_ana_off, (size_t)nana + 1, &ix.ana_inst_off, err) ||
      !download(d_ana_cc, nana, &ix.ana_charcount, err) || !download(d_inst_vocab, n, &ix.inst_vocab, err) ||
      !download(d_inst_freq, n, &ix.inst_freq, err) || !download(d_inst_rows, (size_t)n * ix.norm_stride, &ix.inst_rows, err) ||
      !download(d_table, slots, &ix.table, err) || !download(reinterpret_cast<uint64_t*>(d_bloom), words, &ix.bloom, err) ||
      !download(d_post_ana, nposts, &ix.post_ana, err) || !download(d_post_cls, nposts, &ix.post_cls, err))
    return false;
  if (n_shards > 1 && !download(d_gid, n, &ix.inst_gid, err)) return false;
  // max_key_bits of the (shard's) anagrams, like the host build
  {
    const unsigned nt_max = host_threads();
    std::vector<uint32_t> mx(nt_max, 0);
    parallel_ranges(nana, 1u << 16, [&](unsigned t, uint64_t lo, uint64_t hi) {
      uint32_t m = 0;
      for (uint64_t r = lo; r < hi; ++r) {
        const Key192& k = ix.ana_key[r];
        const uint32_t b = k.w2 ? 192 - (uint32_t)__builtin_clzll(k.w2) : (k.w1 ? 128 - (uint32_t)__builtin_clzll(k.w1) : (k.w0 ? 64 - (uint32_t)__builtin_clzll(k.w0) : 0));
        m = std::max(m, b);
      }
      mx[t] = m;
    });
    for (uint32_t m : mx) ix.max_key_bits = std::max(ix.max_key_bits, m);
  }
  pt.lap("gpu build: download");
  hm->index = std::move(ix);
  hm->build_language_model();
  hm->built = true;
  return true;
}

}  // namespace anl
