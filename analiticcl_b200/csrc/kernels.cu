// kernels.cu -- the hand-written sm_100a kernels of the variant-lookup hot path.
//
//   encode_kernel    : query normalisation (normalize_to_alphabet, src/anahash.rs:50-80).
//   bloom_kernel     : candidate generation, stage 1.  Replaces the search side of find_nearest_anahashes
//                      (src/lib.rs:1143-1308): BFS over deletions (src/iterators.rs:153-187) + linear scan of
//                      sortedindex[charcount] with a bignum modulo per test (src/lib.rs:1268-1281) become a
//                      canonical enumeration of the neighbourhood with one add + one Bloom word per node.
//   exact_kernel     : candidate generation, stage 2: table slot, postings, exact 192-bit verification.
//   probe_kernel     : both stages fused in one kernel (StopAtExactMatch, reruns, queue overflow).
//   prefilter_kernel : bit-parallel OSA distance of every candidate; exact rejection of the far ones.
//   score_kernel     : candidate scoring + ranking.  Replaces gather_instances (src/lib.rs:1311-1402),
//                      damerau_levenshtein / longest_common_substring_length / common_prefix_length /
//                      common_suffix_length (src/distance.rs:101-231) and score_and_rank
//                      (src/lib.rs:1405-1653) up to and including crop and cut-off.
//   confusable_kernel, finish_kernel : rescore_confusables (src/lib.rs:1656-1663) on the device, then
//                      re-rank / crop / cut-off in place.
//   merge_kernel     : lexicon-sharded mode, ranks the all-gathered per-shard survivors.
//
// One warp owns one query at a time in the per-query kernels (queries are independent,
// src/bin/analiticcl.rs:445-448); warps pull queries from a global counter, so the grids are persistent:
// SMs x resident CTAs.  Integer / byte work only -- no tensor cores (see DESIGN.md for the rooflines).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include <atomic>

#include "device_types.h"
#include "editscript_fixed.h"
#include "kernel_common.cuh"
#include "kernels.h"

namespace anl {

// process-wide count of kernel launches issued by this library (bench.py reports it as gpu_launches); atomic: with
// several devices every device has its own dispatcher thread
static std::atomic<unsigned long long> g_kernel_launches{0};
unsigned long long kernel_launches() { return g_kernel_launches.load(std::memory_order_relaxed); }
void count_launch(unsigned n) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

long long merge_grid(int sm_count, uint32_t n);

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------

// DistanceThreshold -> absolute distance for an input of `len` symbols (src/lib.rs:982-1012).
__device__ __forceinline__ uint32_t apply_threshold(const Threshold& t, uint32_t len) {
  if (t.kind == 2) {
    uint32_t h = len >> 1;  // floor(len as f64 / 2.0) as u8 (saturating)
    if (h > 255u) h = 255u;
    return min(t.value & 0xFFu, h);
  }
  float f = floorf(__fmul_rn((float)len, t.ratio));  // (len as f32 * x).floor() as u8 (saturating)
  uint32_t v = (f != f || f <= 0.f) ? 0u : (f >= 255.f ? 255u : (uint32_t)f);
  return min(v, t.kind == 0 ? 12u : (t.value & 0xFFu));
}

// k *= m over 192 bits; false if the product does not fit (then it cannot equal any indexed key).
__device__ __forceinline__ bool mul192(uint64_t& w0, uint64_t& w1, uint64_t& w2, uint64_t m) {
  uint64_t l0 = w0 * m, h0 = __umul64hi(w0, m);
  uint64_t l1 = w1 * m, h1 = __umul64hi(w1, m);
  uint64_t l2 = w2 * m, h2 = __umul64hi(w2, m);
  uint64_t r1 = l1 + h0;
  uint64_t c1 = r1 < l1;
  uint64_t r2 = l2 + h1;
  uint64_t c2 = r2 < l2;
  r2 += c1;
  c2 += (r2 < c1);
  w0 = l0;
  w1 = r1;
  w2 = r2;
  return (h2 + c2) == 0;
}

__device__ __forceinline__ bool ccbit(const uint64_t* mask, uint32_t cc) {
  return cc < 256u && ((mask[cc >> 6] >> (cc & 63)) & 1ull);
}

// ================================================================================================
// Kernel 0: query normalisation (normalize_to_alphabet, src/anahash.rs:50-80)
// ================================================================================================
// One thread per query: greedy matching of alphabet members at each character position, in alphabet-file
// order (the members starting with the current byte are listed in (line, member) order); an unmatched
// character becomes the UNK symbol.  Writes the encoded row (len, flags, symbols) the other kernels read.
__device__ __forceinline__ uint32_t dev_u8len(uint32_t c) {
  if (c < 0x80) return 1;
  if ((c & 0xE0) == 0xC0) return 2;
  if ((c & 0xF0) == 0xE0) return 3;
  if ((c & 0xF8) == 0xF0) return 4;
  return 1;
}
__global__ void __launch_bounds__(128)
encode_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ qblob,
              const uint32_t* __restrict__ qboff, uint32_t n, uint8_t* __restrict__ rows, uint8_t* __restrict__ status) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t b0 = qboff[q], len = qboff[q + 1] - b0;
  const uint8_t* __restrict__ s = qblob + b0;
  uint8_t* __restrict__ row = rows + (size_t)q * bp.query_stride;
  const uint32_t cap = bp.query_stride - 2;
  const AlphaMember* __restrict__ members = ix->alpha_members;
  const uint32_t unk = ix->unk_symbol;
  uint32_t count = 0, i = 0;
  while (i < len) {
    const uint32_t c = s[i];
    const AlphaFirst af = ix->alpha_first[c];
    uint32_t step = 0, sym = unk;
    for (uint32_t m = af.first; m < (uint32_t)af.first + af.count; ++m) {
      const uint32_t ml = members[m].len;
      if (i + ml > len) continue;
      bool eq = true;
      for (uint32_t k = 1; k < ml && eq; ++k) eq = s[i + k] == members[m].bytes[k];
      if (eq) {
        step = ml;
        sym = members[m].seqnr;
        break;
      }
    }
    if (step == 0) step = dev_u8len(c);
    if (count < cap) row[2 + count] = (uint8_t)sym;
    i += step;
    ++count;
  }
  uint32_t flags = 0;
  if (len > 0) {
    // is the first character lowercase?  (src/lib.rs:1374)
    const uint32_t c0 = s[0];
    uint32_t cp = c0;
    if (c0 >= 0x80) {
      uint32_t l = dev_u8len(c0);
      if (l > len) l = len;
      cp = c0 & (0xFFu >> (l + 1));
      for (uint32_t k = 1; k < l; ++k) cp = (cp << 6) | (s[k] & 0x3Fu);
      // binary search in the inclusive [lo, hi] ranges
      uint32_t lo = 0, hi = ix->n_lower_ranges;
      const uint32_t* __restrict__ t = ix->lower_ranges;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) / 2;
        if (cp > t[2 * mid + 1]) lo = mid + 1; else hi = mid;
      }
      if (lo < ix->n_lower_ranges && cp >= t[2 * lo]) flags |= Q_FIRST_LOWER;
    } else if (cp >= 'a' && cp <= 'z') {
      flags |= Q_FIRST_LOWER;
    }
  }
  uint8_t st = ENC_OK;
  if (count > (uint32_t)ANL_MAX_SYMBOLS) {
    // longer than the device rows hold.  If even after max_anagram_distance deletions the query is longer
    // than the longest indexed entry, the result is empty (exact); otherwise it is outside the supported range
    const uint32_t ka = apply_threshold(bp.max_anagram, count);
    st = (count > ix->max_charcount + ka) ? ENC_TOO_LONG_EMPTY : ENC_TOO_LONG_UNSUPPORTED;
    count = 0;
  }
  row[0] = (uint8_t)count;
  row[1] = (uint8_t)flags;
  status[q] = st;
}

// ================================================================================================
// Upload-time transformations of the index (device only)
// ================================================================================================
// The host index keeps keys, instance offsets, slots and postings as separate arrays (and so does its file format).
// On the device every separate array is another random access per verified posting, so the upload folds them:
// anagram records {key, first instance, count} in one 32-byte sector, and the only posting of a slot into the slot.
__global__ void ana_rec_kernel(const Key192* __restrict__ key, const uint32_t* __restrict__ inst_off, uint32_t n, AnaRec* __restrict__ rec) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  AnaRec a;
  a.key = key[r];
  a.inst_off = inst_off[r];
  a.inst_cnt = inst_off[r + 1] - inst_off[r];
  rec[r] = a;
}
__global__ void inline_slots_kernel(Slot* __restrict__ table, uint64_t slots, const uint32_t* __restrict__ post_ana,
                                    const uint8_t* __restrict__ post_cls) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= slots) return;
  Slot s = table[i];
  if (s.post_cnt == 1 && !(s.pad & SLOT_INLINE)) {
    s.pad = (uint16_t)(SLOT_INLINE | post_cls[s.post_off]);
    s.post_off = post_ana[s.post_off];
    table[i] = s;
  }
}
cudaError_t finish_device_index(const Key192* ana_key, const uint32_t* ana_inst_off, uint32_t n_anagrams, AnaRec* ana_rec, Slot* table,
                                uint64_t slots, const uint32_t* post_ana, const uint8_t* post_cls, cudaStream_t stream) {
  if (n_anagrams) ana_rec_kernel<<<(n_anagrams + 255) / 256, 256, 0, stream>>>(ana_key, ana_inst_off, n_anagrams, ana_rec);
  if (slots) inline_slots_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>(table, slots, post_ana, post_cls);
  g_kernel_launches += 2;
  return cudaGetLastError();
}

// ================================================================================================
// Kernel 1: candidate generation
// ================================================================================================
// Every node X = D + I' of the query's neighbourhood (D: a sub-multiset of the focus after d deletions,
// I': a multiset of inserted symbols) is fingerprinted with the LINEAR multiset hash of
// device_types.h: mhash(X) = mhash(F) - sum rnd(deleted) + sum rnd(inserted).  One 64-bit add per node,
// one 8-byte Bloom word per probe.  Nodes that pass the filter are staged (self-contained 32-byte
// records) and looked up exactly 32 at a time; every posting found is verified EXACTLY with
// multi-limb arithmetic on the prime-product keys, by cross-multiplication instead of division:
//     key(C) * prod(deleted)  ==  key(F) * prod(inserted) * p_x        (256-bit products)
// so the only per-query bignum work is key(F) itself.
//
// The kernel is written for a small instruction footprint (ncu showed the previous, fully inlined
// version stalled 57 % of the time on instruction fetch: 92 KB of SASS against a 32 KB L1.5 I-cache):
// the exact stage is one non-inlined function, cold general-case code is kept out of line, and
// per-symbol work runs in short loops over packed bytes instead of unrolled register arrays.
#ifndef ANL_K1_WARPS
#define ANL_K1_WARPS 8
#endif
#ifndef ANL_K1_MIN_CTAS
#define ANL_K1_MIN_CTAS 4
#endif
#ifndef ANL_K1_DCH
#define ANL_K1_DCH 64
#endif
constexpr int K1_WARPS = ANL_K1_WARPS;
constexpr int DCH = ANL_K1_DCH;  // deletion entries with insertion budget buffered per pass
constexpr int SQ = 64;   // staging queue capacity (filter positives waiting for the exact lookup)

// deleted symbols packed one per byte (ascending, unused bytes 0xFF), number of deletions in the top byte
__device__ __forceinline__ uint32_t dd_count(uint64_t dd) { return (uint32_t)(dd >> 56); }
// true iff one of the low six bytes of v equals the byte b
__device__ __forceinline__ bool has_byte6(uint64_t v, uint32_t b) {
  const uint64_t x = (v ^ (0x0000010101010101ULL * b)) & 0x0000FFFFFFFFFFFFULL;
  return ((x - 0x0000010101010101ULL) & ~x & 0x0000808080808080ULL) != 0;
}

struct __align__(8) DEntry {  // an element of the deletion neighbourhood that still has insertion budget
  uint64_t h;                 // mhash(D)
  uint64_t dprod;             // product of the primes of the deleted symbols (< 2^60)
  uint64_t dd;                // deleted symbols, packed
};
struct __align__(16) SEntry {  // a node X = D + I' that passed the Bloom filter
  uint64_t h;                  // mhash(X)
  uint64_t dprod;
  uint64_t dd;
  uint32_t t;                  // index of I' in the multiset table (unused when isz == 0)
  uint8_t isz;                 // |I'|
  uint8_t imax;                // largest symbol of I' (0 if empty)
  uint8_t pad[2];
};
struct K1Warp {
  SEntry sq[SQ];
  DEntry dch[DCH];
  uint64_t nprod[32];  // exact stage: product of the primes of I' per staged node (0 = does not fit: general path)
  uint64_t kF[3];      // exact key of the focus (valid iff kF_ok)
  uint32_t pfx[33];    // exact stage: exclusive prefix of posting counts (+ sentinel)
  uint32_t poff[32];   // exact stage: first posting of each staged node (inline slot: the anagram rank)
  uint32_t pinl[32];   // exact stage: 0xFFFFFFFF, or the class of the slot's inline posting
  uint32_t stat[6];    // per-warp work counters of the non-inlined stages: slots, postings, anagrams, instances, probes, passes
  uint32_t binomL[8];  // C(L, d)
  uint32_t kF_ok, nhits, L, ka;
  uint32_t* hits_q;    // hit list of the current query
  uint8_t sorted[256];
};
struct K1Shared {
  DeviceIndex ix;     // block-local copy of the model constants (pointers, masks, small tables)
  uint64_t rnd[256];  // class_rnd of every symbol
  uint32_t hit_cap;
  K1Warp w[K1_WARPS];
};

// a[3] * m -> r[4]
__device__ __forceinline__ void mul192x64(uint64_t a0, uint64_t a1, uint64_t a2, uint64_t m, uint64_t& r0, uint64_t& r1,
                                          uint64_t& r2, uint64_t& r3) {
  const uint64_t l0 = a0 * m, h0 = __umul64hi(a0, m);
  const uint64_t l1 = a1 * m, h1 = __umul64hi(a1, m);
  const uint64_t l2 = a2 * m, h2 = __umul64hi(a2, m);
  r0 = l0;
  r1 = l1 + h0;
  const uint64_t c1 = r1 < l1;
  r2 = l2 + h1;
  uint64_t c2 = r2 < l2;
  r2 += c1;
  c2 += r2 < c1;
  r3 = h2 + c2;
}

// General exact verification (cold): key(F - deleted + I' + x) == key(C) by multiplying out the node's
// symbols.  Used when key(F) itself does not fit 192 bits or the inserted product does not fit 53 bits.
__device__ __noinline__ bool verify_general(const K1Shared& S, const K1Warp& W, const SEntry& s, uint32_t x, const Key192& ck) {
  uint64_t k0 = 1, k1 = 0, k2 = 0, pp = 1;
  bool ok = true;
  const uint32_t d = dd_count(s.dd);
  uint32_t dp = 0;  // next deleted symbol to drop (both lists ascend)
  #pragma unroll 1
  for (uint32_t p = 0; p < W.L && ok; ++p) {
    const uint32_t sym = W.sorted[p];
    if (dp < d && ((s.dd >> (8 * dp)) & 0xFF) == sym) {
      ++dp;
      continue;
    }
    pp *= S.ix.prime_of[sym];
    if (pp >> 53) {  // next factor (< 2^10) could overflow 64 bits: flush
      ok = mul192(k0, k1, k2, pp);
      pp = 1;
    }
  }
  if (ok && s.isz) {
    const uint64_t cls = __ldg(reinterpret_cast<const uint64_t*>(S.ix.mset + s.t) + 1);  // cls[6] | j | maxcls
    #pragma unroll 1
    for (uint32_t a = 0; a < s.isz && ok; ++a) {
      pp *= S.ix.prime_of[(cls >> (8 * a)) & 0xFF];
      if (pp >> 53) {
        ok = mul192(k0, k1, k2, pp);
        pp = 1;
      }
    }
  }
  if (ok && x != POST_SELF) {
    pp *= S.ix.prime_of[x];
    if (pp >> 53) {
      ok = mul192(k0, k1, k2, pp);
      pp = 1;
    }
  }
  if (ok && pp > 1) ok = mul192(k0, k1, k2, pp);
  return ok && ck.w0 == k0 && ck.w1 == k1 && ck.w2 == k2;
}

// The exact stage for the first cnt <= 32 staged nodes, in two lane-parallel steps: (1) one node per
// lane finds its table slot by fingerprint; (2) the postings of all nodes are flattened (warp scan)
// and verified one posting per lane -- against the canonical-generation rules (each indexed anagram C
// is produced exactly once: from D = F meet C and the ascending insertion order) and against the
// anagram's own key.  Not inlined: one copy, called from every place that stages nodes.
__device__ __noinline__ void exact_stage(K1Shared& S, K1Warp& W, uint32_t cnt) {
  const uint32_t lane = lane_id();
  const DeviceIndex& ix = S.ix;
  uint32_t c_steps = 0, c_post = 0, c_ana = 0, c_inst = 0;
  uint32_t poff = 0, pcnt = 0, pinl = 0xFFFFFFFFu;
  if (lane < cnt) {
    const uint64_t fp = W.sq[lane].h;
    uint64_t idx = fp_index(fp, ix.table_mask);
    for (;;) {
      const Slot sl = ix.table[idx];
      ++c_steps;
      if (sl.post_cnt == 0) break;
      if (sl.fp == fp) {
        poff = sl.post_off;
        pcnt = sl.post_cnt;
        if (sl.pad & SLOT_INLINE) pinl = sl.pad & 0xFFu;
        break;
      }
      idx = (idx + 1) & ix.table_mask;
    }
    if (pcnt) {
      uint64_t np = 1;
      const uint32_t isz = W.sq[lane].isz;
      if (isz) {
        const uint64_t cls = __ldg(reinterpret_cast<const uint64_t*>(ix.mset + W.sq[lane].t) + 1);  // cls[6] | j | maxcls
        #pragma unroll 1
        for (uint32_t a = 0; a < isz; ++a) np = (np >> 53) ? 0 : np * ix.prime_of[(cls >> (8 * a)) & 0xFF];
        if (np >> 53) np = 0;
      }
      W.nprod[lane] = np;
    }
  }
  uint32_t incl = pcnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  const uint32_t total = __shfl_sync(FULL, incl, 31);
  W.pfx[lane] = incl - pcnt;
  W.poff[lane] = poff;
  W.pinl[lane] = pinl;
  __syncwarp();
  const uint32_t ka = W.ka;
  const int sd = ix.sd;
  const bool kF_ok = W.kF_ok != 0;
  for (uint32_t t0 = 0; t0 < total; t0 += 32) {
    const uint32_t t = t0 + lane;
    if (t < total) {
      uint32_t owner = 0;  // largest lane whose exclusive prefix is <= t
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1)
        if (W.pfx[owner + step] <= t) owner += step;
      const uint32_t p = W.poff[owner] + (t - W.pfx[owner]);
      const SEntry& s = W.sq[owner];
      const uint32_t inl = W.pinl[owner];
      const uint32_t r = inl != 0xFFFFFFFFu ? p : __ldg(ix.post_ana + p);
      const uint32_t x = inl != 0xFFFFFFFFu ? inl : __ldg(ix.post_cls + p);
      ++c_post;
      const uint64_t dd = s.dd;
      const uint32_t isz = s.isz;
      bool ok;
      if (x == POST_SELF) {
        ok = (sd == 0) || (isz == 0);
      } else {
        // budget left for the last insertion; ascending insertion order; never re-insert a deleted class
        ok = ka >= dd_count(dd) + isz + 1 && (isz == 0 || x >= s.imax) && !has_byte6(dd, x);
      }
      if (ok) {
        const AnaRec ar = ix.ana_rec[r];
        const Key192 ck = ar.key;
        const uint64_t np = W.nprod[owner];
        bool match;
        if (kF_ok && np != 0) {
          const uint64_t n = (x == POST_SELF) ? np : np * ix.prime_of[x];
          uint64_t a0, a1, a2, a3, b0, b1, b2, b3;
          mul192x64(ck.w0, ck.w1, ck.w2, s.dprod, a0, a1, a2, a3);
          mul192x64(W.kF[0], W.kF[1], W.kF[2], n, b0, b1, b2, b3);
          match = a0 == b0 && a1 == b1 && a2 == b2 && a3 == b3;
        } else {
          match = verify_general(S, W, s, x, ck);
        }
        if (match) {  // else: fingerprint collision
          const uint32_t io = ar.inst_off, n = ar.inst_cnt;
          ++c_ana;
          c_inst += n;
          const uint32_t pos = atomicAdd(&W.nhits, n);
          #pragma unroll 1
          for (uint32_t q = 0; q < n; ++q)
            if (pos + q < S.hit_cap) W.hits_q[pos + q] = io + q;
        }
      }
    }
  }
  {
    const uint32_t r0 = __reduce_add_sync(FULL, c_steps), r1 = __reduce_add_sync(FULL, c_post);
    const uint32_t r2 = __reduce_add_sync(FULL, c_ana), r3 = __reduce_add_sync(FULL, c_inst);
    if (lane == 0) {
      W.stat[0] += r0;
      W.stat[1] += r1;
      W.stat[2] += r2;
      W.stat[3] += r3;
    }
  }
  __syncwarp();
}

// Queue the nodes that passed the Bloom filter; whenever 32 are waiting, run the exact stage on them.
__device__ __forceinline__ void stage(K1Shared& S, K1Warp& W, uint32_t& sqn, bool pass, uint64_t h, uint64_t dprod, uint64_t dd,
                                      uint32_t t, uint32_t isz, uint32_t imax) {
  const uint32_t ballot = __ballot_sync(FULL, pass);
  if (ballot == 0) return;
  const uint32_t lane = lane_id();
  if (pass) {
    SEntry s;
    s.h = h;
    s.dprod = dprod;
    s.dd = dd;
    s.t = t;
    s.isz = (uint8_t)isz;
    s.imax = (uint8_t)imax;
    s.pad[0] = s.pad[1] = 0;
    W.sq[sqn + __popc(ballot & lanemask_lt())] = s;
  }
  sqn += __popc(ballot);
  __syncwarp();
  if (sqn >= 32) {
    exact_stage(S, W, 32);
    const uint32_t rest = sqn - 32;  // < 32: move the tail to the front
    SEntry tmp;
    if (lane < rest) tmp = W.sq[32 + lane];
    __syncwarp();
    if (lane < rest) W.sq[lane] = tmp;
    __syncwarp();
    sqn = rest;
  }
}

__device__ __forceinline__ bool bloom_pass(const K1Shared& S, uint64_t h) {
  return bloom_test(S.ix.bloom, S.ix.bloom_mask, h);
}

// The nodes X = D + I', |I'| >= 1, of the buffered deletion entries; lanes stride over the multiset table.
// Returns the new staging-queue length.
__device__ __noinline__ uint32_t insertion_nodes(K1Shared& S, K1Warp& W, uint32_t nD, uint32_t sqn) {
  const uint32_t lane = lane_id();
  const DeviceIndex& ix = S.ix;
  uint32_t probes = 0, passes = 0;
  for (uint32_t e = 0; e < nD; ++e) {
    const DEntry de = W.dch[e];  // warp-uniform broadcast
    const uint32_t d = dd_count(de.dd);
    const int jmax = (int)W.ka - (int)d - ix.sd;
    for (int j = 1; j <= jmax; ++j) {
      const uint32_t cx = W.L - d + j;
      const bool useful = (ix.sd == 0) ? ccbit(ix.charcount_mask, cx) : ccbit(ix.charcount_mask, cx + 1);
      if (!useful) continue;
      const uint32_t lo = ix.mset_end[j - 1], hi = ix.mset_end[j];
      for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t t = base + lane;
        bool active = t < hi;
        uint64_t h = de.h;
        uint32_t imax = 0;
        if (active) {
          const ulonglong2 me = __ldg(reinterpret_cast<const ulonglong2*>(ix.mset + t));  // {hsum, cls[6] | j | maxcls}
          h += me.x;
          imax = (uint32_t)(me.y >> 56);
          // canonical generation: never re-insert a deleted class
          #pragma unroll 1
          for (uint32_t b = 0; b < d; ++b) active = active && !has_byte6(me.y, (uint32_t)(de.dd >> (8 * b)) & 0xFF);
        }
        const bool pass = active && bloom_pass(S, h);
        probes += active;
        passes += pass;
        stage(S, W, sqn, pass, h, de.dprod, de.dd, t, (uint32_t)j, imax);
      }
    }
  }
  {
    const uint32_t r4 = __reduce_add_sync(FULL, probes), r5 = __reduce_add_sync(FULL, passes);
    if (lane == 0) {
      W.stat[4] += r4;
      W.stat[5] += r5;
    }
  }
  return sqn;
}

// General unranking of deletion set `rem` of size d (queries longer than COLEX_N symbols or more than 3
// deletions): combinadic positions p0 < ... < p(d-1), packed one per byte.  Cold path.
__device__ __noinline__ uint64_t unrank_general(const uint32_t* __restrict__ binom, uint32_t L, uint32_t d, uint32_t rem) {
  uint64_t pk = 0;
  int cpos = (int)L;
  #pragma unroll 1
  for (int i = (int)d; i >= 1; --i) {
    --cpos;
    while (__ldg(binom + cpos * 8 + i) > rem) --cpos;
    pk |= (uint64_t)(uint32_t)cpos << (8 * (i - 1));
    rem -= __ldg(binom + cpos * 8 + i);
  }
  return pk;
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(K1_WARPS * 32, MIN_CTAS)
probe_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
             const uint32_t* __restrict__ qlist, uint32_t nq, uint32_t* __restrict__ hits, uint32_t* __restrict__ hit_count,
             uint32_t* __restrict__ qflags, unsigned int* work, Counters* counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K1Shared& S = *reinterpret_cast<K1Shared*>(smem_raw);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(ix);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&S.ix);
    for (uint32_t i = threadIdx.x; i < sizeof(DeviceIndex) / 4; i += blockDim.x) dst[i] = src[i];
  }
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) S.rnd[i] = class_rnd(i);
  if (threadIdx.x == 0) S.hit_cap = bp.hit_cap;
  const uint32_t lane = lane_id();
  K1Warp& W = S.w[threadIdx.x >> 5];
  if (lane < 6) W.stat[lane] = 0;
  __syncthreads();

  uint32_t c_dkeys = 0, c_probes = 0, c_pass = 0;
  const uint32_t max_cc = S.ix.max_charcount;
  const int sd = S.ix.sd;
  const uint32_t* __restrict__ colex2 = S.ix.colex2;
  const uint32_t* __restrict__ colex3 = S.ix.colex3;
  const uint32_t* __restrict__ binom = S.ix.binom;

  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(work, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= nq) break;
    const uint32_t q = qlist ? qlist[qi] : qi;
    const uint8_t* qrow = queries + (size_t)q * bp.query_stride;
    const uint32_t L = qrow[0];
    uint32_t flags = 0;
    const uint32_t ka = L ? apply_threshold(bp.max_anagram, L) : 0;
    if (lane == 0) {
      W.nhits = 0;
      W.L = L;
      W.ka = ka;
      W.hits_q = hits + (size_t)qi * bp.hit_cap;
    }
    __syncwarp();
    if (L == 0) {
      flags = QF_EMPTY;
    } else if (ka > (uint32_t)ANL_MAX_K) {
      flags = QF_UNSUPPORTED;
    } else if (L <= max_cc + ka) {  // else every candidate would be longer than any indexed entry
      // sort the query symbols (rank sort) so equal symbols are adjacent; mhash(F) on the way
      uint64_t hF = 0;
      #pragma unroll 1
      for (uint32_t i = lane; i < L; i += 32) {
        const uint8_t v = qrow[2 + i];
        uint32_t r = 0;
        #pragma unroll 1
        for (uint32_t j = 0; j < L; ++j) {
          const uint8_t u = qrow[2 + j];
          r += (u < v) || (u == v && j < i);
        }
        W.sorted[r] = v;
        hF += S.rnd[v];
      }
      for (int o = 16; o > 0; o >>= 1) hF += __shfl_xor_sync(FULL, hF, o);
      const uint32_t dmax = min(ka, L - 1);
      if (lane <= dmax) W.binomL[lane] = __ldg(binom + L * 8 + lane);
      __syncwarp();
      {
        // exact key of the focus (every lane computes the same value; lane 0 publishes it)
        uint64_t k0 = 1, k1 = 0, k2 = 0, pp = 1;
        bool ok = true;
        #pragma unroll 1
        for (uint32_t p = 0; p < L && ok; ++p) {
          pp *= S.ix.prime_of[W.sorted[p]];
          if (pp >> 53) {
            ok = mul192(k0, k1, k2, pp);
            pp = 1;
          }
        }
        if (ok && pp > 1) ok = mul192(k0, k1, k2, pp);
        if (lane == 0) {
          W.kF[0] = k0;
          W.kF[1] = k1;
          W.kF[2] = k2;
          W.kF_ok = ok ? 1u : 0u;
        }
      }
      __syncwarp();
      uint32_t sqn = 0;
      bool done = false;
      if (bp.stop_at_exact) {
        // StopAtExactMatch (src/lib.rs:1164-1173): if the focus itself is indexed, it is the only result
        if (lane == 0) W.ka = 0;  // only the self posting of X = F is acceptable
        __syncwarp();
        const bool pass = lane == 0 && bloom_pass(S, hF);
        c_probes += lane == 0;
        c_pass += pass;
        stage(S, W, sqn, pass, hF, 1, 0x00FFFFFFFFFFFFFFULL, 0, 0, 0);
        if (sqn) exact_stage(S, W, sqn);
        sqn = 0;
        if (lane == 0) W.ka = ka;
        __syncwarp();
        done = W.nhits > 0;
      }
      if (!done) {
        // enumerate the deletion neighbourhood: all distinct non-empty sub-multisets of the
        // query reachable by d <= ka deletions (src/iterators.rs:153-187 yields the same set)
        uint64_t total64 = 0;
        #pragma unroll 1
        for (uint32_t d = 0; d <= dmax; ++d) total64 += W.binomL[d];
        if (total64 > 0x7FFFFFFFull) {
          flags = QF_UNSUPPORTED;
        } else {
          const uint32_t total = (uint32_t)total64;
          const bool fast = L <= (uint32_t)COLEX_N;
          uint32_t nD = 0;
          for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t t = base + lane;
            bool ok = t < total;
            uint32_t d = 0;
            uint64_t h = hF, dprod = 1, dd = 0x00FFFFFFFFFFFFFFULL;
            if (ok) {
              uint32_t rem = t;
              while (rem >= W.binomL[d]) {
                rem -= W.binomL[d];
                ++d;
              }
              // colex unranking: positions p0 < p1 < ... packed one per byte
              uint64_t pk = rem;  // d == 1: the position itself
              if (fast && d == 2) pk = __ldg(colex2 + rem);
              else if (fast && d == 3) pk = __ldg(colex3 + rem);
              else if (d >= 2) pk = unrank_general(binom, L, d, rem);
              uint32_t prev = 0xFFFFFFFFu;
              #pragma unroll 1
              for (uint32_t i = 0; i < d; ++i) {
                const uint32_t p = (uint32_t)(pk >> (8 * i)) & 0xFF;
                const uint32_t sym = W.sorted[p];
                // canonical: inside a run of equal symbols only a leading part may be deleted
                if (p > 0 && W.sorted[p - 1] == sym && prev + 1 != p) ok = false;
                prev = p;
                h -= S.rnd[sym];
                dprod *= S.ix.prime_of[sym];
                dd = (dd & ~(0xFFULL << (8 * i))) | ((uint64_t)sym << (8 * i));
              }
              dd |= (uint64_t)d << 56;
            }
            c_dkeys += ok;
            // the node X = D itself: useful iff C = D may exist, or (sd = 1) C = D + x may exist within the budget
            const uint32_t cx = L - d;
            const bool probe = ok && (ccbit(S.ix.charcount_mask, cx) || (sd == 1 && ka > d && ccbit(S.ix.charcount_mask, cx + 1)));
            const bool pass = probe && bloom_pass(S, h);
            c_probes += probe;
            c_pass += pass;
            stage(S, W, sqn, pass, h, dprod, dd, 0, 0, 0);
            // entries with budget for insertions beyond the table's own depth are buffered for the insertion pass
            const bool ins = ok && (int)ka - (int)d - sd >= 1;
            const uint32_t ballot = __ballot_sync(FULL, ins);
            if (ins) {
              DEntry de;
              de.h = h;
              de.dprod = dprod;
              de.dd = dd;
              W.dch[nD + __popc(ballot & lanemask_lt())] = de;
            }
            nD += __popc(ballot);
            __syncwarp();
            if (nD > DCH - 32 || (base + 32 >= total && nD)) {
              sqn = insertion_nodes(S, W, nD, sqn);
              nD = 0;
            }
          }
          if (sqn) exact_stage(S, W, sqn);
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      const uint32_t n = W.nhits;
      hit_count[qi] = n;
      if (n > bp.hit_cap) flags |= QF_HIT_OVERFLOW;
      qflags[qi] = flags;
    }
    __syncwarp();
  }

  // flush the work counters (one atomic per counter per warp)
  if (counters) {
    // lane-local u32 counts (a warp handles far fewer than 2^32 units per launch)
    const uint32_t dk = __reduce_add_sync(FULL, c_dkeys);
    const uint32_t pr = __reduce_add_sync(FULL, c_probes), pa = __reduce_add_sync(FULL, c_pass);
    __syncwarp();
    if (lane == 0) {
      atomicAdd(&counters->deletion_keys, (unsigned long long)dk);
      atomicAdd(&counters->probes, (unsigned long long)(pr + W.stat[4]));
      atomicAdd(&counters->filter_pass, (unsigned long long)(pa + W.stat[5]));
      atomicAdd(&counters->table_steps, (unsigned long long)W.stat[0]);
      atomicAdd(&counters->postings, (unsigned long long)W.stat[1]);
      atomicAdd(&counters->anagram_hits, (unsigned long long)W.stat[2]);
      atomicAdd(&counters->instance_pairs, (unsigned long long)W.stat[3]);
    }
  }
}

// ================================================================================================
// Kernel 1a (split probe path): the Bloom stage on its own
// ================================================================================================
// The same enumeration as probe_kernel (deletion sets in colex order, insertion multisets, one 64-bit add and
// one Bloom word per node), but every node that passes the filter is appended to a global queue for
// exact_kernel instead of being looked up by this warp.  Without the exact stage the kernel needs 2 KB of
// shared memory per warp and few registers, so twice as many warps are resident: the stage is bound by the
// latency of the Bloom word loads (ncu: long scoreboard), and more warps in flight is what hides it.
constexpr int KB_WARPS = 8;
#ifndef ANL_KB_MIN_CTAS
#define ANL_KB_MIN_CTAS 6
#endif
struct KBWarp {
  DEntry dch[DCH];
  uint32_t binomL[8];
  uint8_t sorted[256];
};
struct KBShared {
  uint64_t rnd[256];
  uint32_t prime_of[256];
  uint64_t charcount_mask[4];
  uint32_t mset_end[ANL_MAX_K + 1];
  KBWarp w[KB_WARPS];
};

// Queue space is reserved in chunks of QCHUNK entries per warp: one global atomic per chunk instead of one per
// push (with a single shared cursor the pushes ran at the same-address atomic rate of the L2 -- 24 % of the
// kernel's stall samples).  A push that does not fit the current chunk spills its tail into a fresh chunk, so
// chunks are filled completely; only a warp's last chunk keeps a hole, marked with QHOLE entries at exit.
constexpr uint32_t QCHUNK = 128;
constexpr uint32_t QHOLE = 0xFFFFFFFFu;  // QEntry.qi of an unused slot
struct QCursor {
  uint32_t pos, end;  // next free slot and end of this warp's current chunk (warp-uniform)
};
__device__ __forceinline__ void queue_push(QEntry* __restrict__ queue, uint32_t queue_cap, unsigned int* cursor, QCursor& qc, bool pass,
                                           uint64_t h, uint64_t dprod, uint64_t dd, uint32_t t, uint32_t isz, uint32_t imax,
                                           uint32_t qi) {
  const uint32_t ballot = __ballot_sync(FULL, pass);
  if (ballot == 0) return;
  const uint32_t need = __popc(ballot), room = qc.end - qc.pos;
  uint32_t fresh = 0;
  if (need > room) {  // warp-uniform
    if (lane_id() == 0) fresh = atomicAdd(cursor, QCHUNK);
    fresh = __shfl_sync(FULL, fresh, 0);
  }
  if (pass) {
    const uint32_t r = __popc(ballot & lanemask_lt());
    const uint32_t pos = r < room ? qc.pos + r : fresh + (r - room);
    if (pos < queue_cap) {
      QEntry e;
      e.h = h;
      e.dprod = dprod;
      e.dd = dd;
      e.t = t;
      e.qi = qi;
      e.isz = (uint8_t)isz;
      e.imax = (uint8_t)imax;
#pragma unroll
      for (int k = 0; k < 6; ++k) e.pad[k] = 0;
      queue[pos] = e;
    }
  }
  if (need > room) {
    qc.pos = fresh + (need - room);
    qc.end = fresh + QCHUNK;
  } else {
    qc.pos += need;
  }
}

constexpr uint32_t KB_HEAVY_LEN = 16;  // queries at least this long are enumerated in the first phase
__global__ void __launch_bounds__(KB_WARPS * 32, ANL_KB_MIN_CTAS)
bloom_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
             const uint32_t* __restrict__ qlist, uint32_t nq, uint32_t* __restrict__ qflags, unsigned int* work, Counters* counters,
             QEntry* __restrict__ queue, uint32_t queue_cap, QCtx* __restrict__ qctx) {
  __shared__ KBShared S;
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
    S.rnd[i] = class_rnd(i);
    S.prime_of[i] = ix->prime_of[i];
  }
  if (threadIdx.x < 4) S.charcount_mask[threadIdx.x] = ix->charcount_mask[threadIdx.x];
  if (threadIdx.x <= (unsigned)ANL_MAX_K) S.mset_end[threadIdx.x] = ix->mset_end[threadIdx.x];
  __syncthreads();
  const uint32_t lane = lane_id();
  KBWarp& W = S.w[threadIdx.x >> 5];
  const uint64_t* __restrict__ bloom = ix->bloom;
  const uint64_t bloom_wmask = ix->bloom_mask;
  const MsetEntry* __restrict__ mset = ix->mset;
  const uint32_t max_cc = ix->max_charcount;
  const int sd = ix->sd;
  const uint32_t* __restrict__ colex2 = ix->colex2;
  const uint32_t* __restrict__ colex3 = ix->colex3;
  const uint32_t* __restrict__ binom = ix->binom;
  unsigned int* cursor = work + 4;
  QCursor qc;
  qc.pos = qc.end = 0;
  uint32_t c_dkeys = 0, c_probes = 0, c_pass = 0;

  // Long queries first (their neighbourhood grows with C(L, k)): phase 0 walks the batch four queries per grab and
  // takes the long ones, phase 1 (its own counter, work[6]) the rest -- see score_kernel for the reason.
  uint32_t gbase = 0, gmask = 0, phase = 0;
  for (;;) {
    if (gmask == 0) {
      const uint32_t grab = phase == 0 ? 4u : 1u;  // (few per grab: a warp works its grab off serially)
      if (lane == 0) gbase = atomicAdd(phase == 0 ? work : work + 6, grab);
      gbase = __shfl_sync(FULL, gbase, 0);
      if (gbase >= nq) {
        if (phase == 0) {
          phase = 1;
          continue;
        }
        break;
      }
      bool mine = false;
      if (lane < grab && gbase + lane < nq) {
        const uint32_t q0 = qlist ? qlist[gbase + lane] : gbase + lane;
        mine = (queries[(size_t)q0 * bp.query_stride] >= KB_HEAVY_LEN) == (phase == 0);
      }
      gmask = __ballot_sync(FULL, mine);
      if (gmask == 0) continue;
    }
    const uint32_t qi = gbase + (uint32_t)__ffs(gmask) - 1;
    gmask &= gmask - 1;
    const uint32_t q = qlist ? qlist[qi] : qi;
    const uint8_t* qrow = queries + (size_t)q * bp.query_stride;
    const uint32_t L = qrow[0];
    uint32_t flags = 0;
    const uint32_t ka = L ? apply_threshold(bp.max_anagram, L) : 0;
    if (L == 0) {
      flags = QF_EMPTY;
    } else if (ka > (uint32_t)ANL_MAX_K) {
      flags = QF_UNSUPPORTED;
    } else if (L <= max_cc + ka) {  // else every candidate would be longer than any indexed entry
      // sort the query symbols (rank sort) so equal symbols are adjacent; mhash(F) on the way
      uint64_t hF = 0;
#pragma unroll 1
      for (uint32_t i = lane; i < L; i += 32) {
        const uint8_t v = qrow[2 + i];
        uint32_t r = 0;
#pragma unroll 1
        for (uint32_t j = 0; j < L; ++j) {
          const uint8_t u = qrow[2 + j];
          r += (u < v) || (u == v && j < i);
        }
        W.sorted[r] = v;
        hF += S.rnd[v];
      }
      for (int o = 16; o > 0; o >>= 1) hF += __shfl_xor_sync(FULL, hF, o);
      const uint32_t dmax = min(ka, L - 1);
      if (lane <= dmax) W.binomL[lane] = __ldg(binom + L * 8 + lane);
      __syncwarp();
      {
        // exact key of the focus for the exact stage (every lane computes the same value; lane 0 publishes it)
        uint64_t k0 = 1, k1 = 0, k2 = 0, pp = 1;
        bool ok = true;
#pragma unroll 1
        for (uint32_t p = 0; p < L && ok; ++p) {
          pp *= S.prime_of[W.sorted[p]];
          if (pp >> 53) {
            ok = mul192(k0, k1, k2, pp);
            pp = 1;
          }
        }
        if (ok && pp > 1) ok = mul192(k0, k1, k2, pp);
        if (lane == 0) {
          QCtx c;
          c.kF[0] = k0;
          c.kF[1] = k1;
          c.kF[2] = k2;
          c.L = L;
          c.ka_ok = ka | (ok ? 0x100u : 0u);
          qctx[qi] = c;
        }
      }
      uint64_t total64 = 0;
#pragma unroll 1
      for (uint32_t d = 0; d <= dmax; ++d) total64 += W.binomL[d];
      if (total64 > 0x7FFFFFFFull) {
        flags = QF_UNSUPPORTED;
      } else {
        const uint32_t total = (uint32_t)total64;
        const bool fast = L <= (uint32_t)COLEX_N;
        uint32_t nD = 0;
        for (uint32_t base = 0; base < total; base += 32) {
          const uint32_t t = base + lane;
          bool ok = t < total;
          uint32_t d = 0;
          uint64_t h = hF, dprod = 1, dd = 0x00FFFFFFFFFFFFFFULL;
          if (ok) {
            uint32_t rem = t;
            while (rem >= W.binomL[d]) {
              rem -= W.binomL[d];
              ++d;
            }
            uint64_t pk = rem;
            if (fast && d == 2) pk = __ldg(colex2 + rem);
            else if (fast && d == 3) pk = __ldg(colex3 + rem);
            else if (d >= 2) pk = unrank_general(binom, L, d, rem);
            uint32_t prev = 0xFFFFFFFFu;
#pragma unroll 1
            for (uint32_t i = 0; i < d; ++i) {
              const uint32_t p = (uint32_t)(pk >> (8 * i)) & 0xFF;
              const uint32_t sym = W.sorted[p];
              if (p > 0 && W.sorted[p - 1] == sym && prev + 1 != p) ok = false;  // canonical: leading part of a run only
              prev = p;
              h -= S.rnd[sym];
              dprod *= S.prime_of[sym];
              dd = (dd & ~(0xFFULL << (8 * i))) | ((uint64_t)sym << (8 * i));
            }
            dd |= (uint64_t)d << 56;
          }
          c_dkeys += ok;
          const uint32_t cx = L - d;
          const bool probe = ok && (ccbit(S.charcount_mask, cx) || (sd == 1 && ka > d && ccbit(S.charcount_mask, cx + 1)));
          bool pass = false;
          if (probe) {
            pass = bloom_test(bloom, bloom_wmask, h);
          }
          c_probes += probe;
          c_pass += pass;
          queue_push(queue, queue_cap, cursor, qc, pass, h, dprod, dd, 0, 0, 0, qi);
          // entries with budget for insertions beyond the table's own depth are buffered for the insertion pass
          const bool ins = ok && (int)ka - (int)d - sd >= 1;
          const uint32_t ballot = __ballot_sync(FULL, ins);
          if (ins) {
            DEntry de;
            de.h = h;
            de.dprod = dprod;
            de.dd = dd;
            W.dch[nD + __popc(ballot & lanemask_lt())] = de;
          }
          nD += __popc(ballot);
          __syncwarp();
          if (nD > DCH - 32 || (base + 32 >= total && nD)) {
            // the nodes X = D + I', |I'| >= 1, of the buffered entries; lanes stride over the multiset table
            for (uint32_t e = 0; e < nD; ++e) {
              const DEntry de = W.dch[e];  // warp-uniform broadcast
              const uint32_t dn = dd_count(de.dd);
              const int jmax = (int)ka - (int)dn - sd;
              for (int j = 1; j <= jmax; ++j) {
                const uint32_t cxj = L - dn + j;
                const bool useful = (sd == 0) ? ccbit(S.charcount_mask, cxj) : ccbit(S.charcount_mask, cxj + 1);
                if (!useful) continue;
                const uint32_t lo = S.mset_end[j - 1], hi = S.mset_end[j];
                for (uint32_t mb = lo; mb < hi; mb += 32) {
                  const uint32_t mt = mb + lane;
                  bool active = mt < hi;
                  uint64_t hx = de.h;
                  uint32_t imax = 0;
                  if (active) {
                    const ulonglong2 me = __ldg(reinterpret_cast<const ulonglong2*>(mset + mt));  // {hsum, cls[6] | j | maxcls}
                    hx += me.x;
                    imax = (uint32_t)(me.y >> 56);
#pragma unroll 1
                    for (uint32_t b = 0; b < dn; ++b) active = active && !has_byte6(me.y, (uint32_t)(de.dd >> (8 * b)) & 0xFF);
                  }
                  bool px = false;
                  if (active) {
                    px = bloom_test(bloom, bloom_wmask, hx);
                  }
                  c_probes += active;
                  c_pass += px;
                  queue_push(queue, queue_cap, cursor, qc, px, hx, de.dprod, de.dd, mt, (uint32_t)j, imax, qi);
                }
              }
            }
            nD = 0;
            __syncwarp();
          }
        }
      }
    }
    if (lane == 0 && flags) atomicOr(qflags + qi, flags);  // exact_kernel counts the hits and raises QF_HIT_OVERFLOW
    __syncwarp();
  }
  // mark the unused tail of this warp's last chunk
  for (uint32_t p = qc.pos + lane; p < qc.end; p += 32)
    if (p < queue_cap) {
      QEntry e;
      e.h = e.dprod = e.dd = 0;
      e.t = 0;
      e.qi = QHOLE;
      e.isz = e.imax = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) e.pad[k] = 0;
      queue[p] = e;
    }
  if (counters) {
    const uint32_t dk = __reduce_add_sync(FULL, c_dkeys);
    const uint32_t pr = __reduce_add_sync(FULL, c_probes), pa = __reduce_add_sync(FULL, c_pass);
    if (lane == 0) {
      atomicAdd(&counters->deletion_keys, (unsigned long long)dk);
      atomicAdd(&counters->probes, (unsigned long long)pr);
      atomicAdd(&counters->filter_pass, (unsigned long long)pa);
    }
  }
}

// ================================================================================================
// Kernel 1b (split probe path): the exact stage over the global queue of staged nodes
// ================================================================================================
// The same two steps as exact_stage -- slot lookup by fingerprint, then every posting verified against the
// canonical-generation rules and the anagram's exact key -- but over the staged nodes of ALL queries, one node
// per lane: every round has 32 busy lanes (inside probe_kernel a warp drains only its own query's nodes, often
// a handful), the Bloom stage no longer carries this code through its instruction cache, and both kernels run
// at the occupancy that suits them.  Hits are appended to the queries' hit lists with one atomic per anagram.
constexpr int KX_WARPS = 8;
struct KXWarp {
  QEntry e[32];
  uint64_t nprod[32];
  uint32_t pfx[33];
  uint32_t poff[32];
  uint32_t pinl[32];  // 0xFFFFFFFF, or the class of the slot's inline posting (poff is then the anagram rank)
};
// General exact verification without the warp's sorted copy of the query (cold, cf. verify_general).
__device__ __noinline__ bool verify_general_q(const DeviceIndex* __restrict__ ix, const uint8_t* __restrict__ qrow, uint32_t L,
                                              const QEntry& s, uint32_t x, const Key192& ck) {
  uint64_t k0 = 1, k1 = 0, k2 = 0, pp = 1;
  bool ok = true;
  const uint32_t d = dd_count(s.dd);
  uint32_t used = 0;  // deleted symbols already dropped
#pragma unroll 1
  for (uint32_t p = 0; p < L && ok; ++p) {
    const uint32_t sym = qrow[2 + p];
    bool dropped = false;
#pragma unroll 1
    for (uint32_t b = 0; b < d && !dropped; ++b)
      if (!((used >> b) & 1u) && ((s.dd >> (8 * b)) & 0xFF) == sym) {
        used |= 1u << b;
        dropped = true;
      }
    if (dropped) continue;
    pp *= ix->prime_of[sym];
    if (pp >> 53) {
      ok = mul192(k0, k1, k2, pp);
      pp = 1;
    }
  }
  if (ok && s.isz) {
    const uint64_t cls = __ldg(reinterpret_cast<const uint64_t*>(ix->mset + s.t) + 1);
#pragma unroll 1
    for (uint32_t a = 0; a < s.isz && ok; ++a) {
      pp *= ix->prime_of[(cls >> (8 * a)) & 0xFF];
      if (pp >> 53) {
        ok = mul192(k0, k1, k2, pp);
        pp = 1;
      }
    }
  }
  if (ok && x != POST_SELF) {
    pp *= ix->prime_of[x];
    if (pp >> 53) {
      ok = mul192(k0, k1, k2, pp);
      pp = 1;
    }
  }
  if (ok && pp > 1) ok = mul192(k0, k1, k2, pp);
  return ok && ck.w0 == k0 && ck.w1 == k1 && ck.w2 == k2;
}

#ifndef ANL_KX_MIN_CTAS
#define ANL_KX_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(KX_WARPS * 32, ANL_KX_MIN_CTAS)
exact_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
             const uint32_t* __restrict__ qlist, const QEntry* __restrict__ queue, uint32_t queue_cap,
             const QCtx* __restrict__ qctx, uint32_t* __restrict__ hits, uint32_t* __restrict__ hit_count,
             uint32_t* __restrict__ qflags, unsigned int* work, Counters* counters) {
  __shared__ KXWarp sm[KX_WARPS];
  KXWarp& W = sm[threadIdx.x >> 5];
  const uint32_t lane = lane_id();
  const uint32_t total = min(work[4], queue_cap);
  const Slot* __restrict__ table = ix->table;
  const uint64_t table_mask = ix->table_mask;
  const int sd = ix->sd;
  uint32_t c_steps = 0, c_post = 0, c_ana = 0, c_inst = 0;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(work + 5, 32u);
    base = __shfl_sync(FULL, base, 0);
    if (base >= total) break;
    const uint32_t i = base + lane;
    uint32_t poff = 0, pcnt = 0, pinl = 0xFFFFFFFFu;
    if (i < total && queue[i].qi != QHOLE) {
      const QEntry e = queue[i];
      W.e[lane] = e;
      const uint64_t fp = e.h;
      uint64_t idx = fp_index(fp, table_mask);
      for (;;) {
        const Slot sl = table[idx];
        ++c_steps;
        if (sl.post_cnt == 0) break;
        if (sl.fp == fp) {
          poff = sl.post_off;
          pcnt = sl.post_cnt;
          if (sl.pad & SLOT_INLINE) pinl = sl.pad & 0xFFu;
          break;
        }
        idx = (idx + 1) & table_mask;
      }
      if (pcnt) {
        uint64_t np = 1;
        if (e.isz) {
          const uint64_t cls = __ldg(reinterpret_cast<const uint64_t*>(ix->mset + e.t) + 1);  // cls[6] | j | maxcls
#pragma unroll 1
          for (uint32_t a = 0; a < e.isz; ++a) np = (np >> 53) ? 0 : np * ix->prime_of[(cls >> (8 * a)) & 0xFF];
          if (np >> 53) np = 0;
        }
        W.nprod[lane] = np;
      }
    }
    uint32_t incl = pcnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    const uint32_t npost = __shfl_sync(FULL, incl, 31);
    W.pfx[lane] = incl - pcnt;
    W.poff[lane] = poff;
    W.pinl[lane] = pinl;
    __syncwarp();
    for (uint32_t t0 = 0; t0 < npost; t0 += 32) {
      const uint32_t t = t0 + lane;
      if (t < npost) {
        uint32_t owner = 0;  // largest lane whose exclusive prefix is <= t
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1)
          if (W.pfx[owner + step] <= t) owner += step;
        const uint32_t p = W.poff[owner] + (t - W.pfx[owner]);
        const QEntry& s = W.e[owner];
        const uint32_t inl = W.pinl[owner];
        const uint32_t r = inl != 0xFFFFFFFFu ? p : __ldg(ix->post_ana + p);
        const uint32_t x = inl != 0xFFFFFFFFu ? inl : __ldg(ix->post_cls + p);
        ++c_post;
        const QCtx cx = qctx[s.qi];
        const uint32_t ka = cx.ka_ok & 0xFFu;
        const uint64_t dd = s.dd;
        const uint32_t isz = s.isz;
        bool ok;
        if (x == POST_SELF) {
          ok = (sd == 0) || (isz == 0);
        } else {
          // budget left for the last insertion; ascending insertion order; never re-insert a deleted class
          ok = ka >= dd_count(dd) + isz + 1 && (isz == 0 || x >= s.imax) && !has_byte6(dd, x);
        }
        if (ok) {
          const AnaRec ar = ix->ana_rec[r];
          const Key192 ck = ar.key;
          const uint64_t np = W.nprod[owner];
          bool match;
          if ((cx.ka_ok & 0x100u) && np != 0) {
            const uint64_t n = (x == POST_SELF) ? np : np * ix->prime_of[x];
            uint64_t a0, a1, a2, a3, b0, b1, b2, b3;
            mul192x64(ck.w0, ck.w1, ck.w2, s.dprod, a0, a1, a2, a3);
            mul192x64(cx.kF[0], cx.kF[1], cx.kF[2], n, b0, b1, b2, b3);
            match = a0 == b0 && a1 == b1 && a2 == b2 && a3 == b3;
          } else {
            const uint32_t q = qlist ? qlist[s.qi] : s.qi;
            match = verify_general_q(ix, queries + (size_t)q * bp.query_stride, cx.L, s, x, ck);
          }
          if (match) {  // else: fingerprint collision
            const uint32_t io = ar.inst_off, n = ar.inst_cnt;
            ++c_ana;
            c_inst += n;
            const uint32_t pos = atomicAdd(hit_count + s.qi, n);
            uint32_t* hq = hits + (size_t)s.qi * bp.hit_cap;
            for (uint32_t k = 0; k < n; ++k)
              if (pos + k < bp.hit_cap) hq[pos + k] = io + k;
            if (pos + n > bp.hit_cap) atomicOr(qflags + s.qi, QF_HIT_OVERFLOW);
          }
        }
      }
    }
    __syncwarp();
  }
  if (counters) {
    const uint32_t r0 = __reduce_add_sync(FULL, c_steps), r1 = __reduce_add_sync(FULL, c_post);
    const uint32_t r2 = __reduce_add_sync(FULL, c_ana), r3 = __reduce_add_sync(FULL, c_inst);
    if (lane == 0) {
      atomicAdd(&counters->table_steps, (unsigned long long)r0);
      atomicAdd(&counters->postings, (unsigned long long)r1);
      atomicAdd(&counters->anagram_hits, (unsigned long long)r2);
      atomicAdd(&counters->instance_pairs, (unsigned long long)r3);
    }
  }
}

// ================================================================================================
// Kernel 2: scoring + ranking
// ================================================================================================
constexpr int K2_WARPS = 4;

// Shared-memory accesses of the DP by 32-bit shared-window address.  (With generic pointers into the
// dynamic shared array, ptxas re-derived the CTA's shared window base -- S2UR SR_CgaCtaId + ULEA --
// inside the inner loop: 5 % of the kernel's stall samples.)
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

struct __align__(8) SurvRec {
  double dist;     // distance score
  double freq;     // absolute, then normalised frequency score
  double key;      // combined score (VariantResult::score), computed once per record
  uint32_t g;      // (global) gather id: the final tie-break key
  uint32_t raw;    // raw frequency
  uint32_t vocab;  // vocabulary id
  uint32_t pad;
};

// shared memory of one warp (dynamic; sized by the number of matrix columns `ML` a launch handles and the ring depth R)
//   q[256]                         query symbols
//   cell[(ML+1)][32] (uint32)      per column j, per lane: {t[j-1], lcs[j], lastrow[j], unused}
//   ring[R][(ML+1)][32] (uint8)    the last R rows of the DL matrix, per lane
__host__ __device__ inline size_t k2_warp_bytes(uint32_t ML, uint32_t R) {
  return 256 + (size_t)(ML + 1) * 32 * 4 + (size_t)R * (ML + 1) * 32;
}

__device__ __noinline__ double result_score(const BatchParams& bp, double dist, double freq) {
  // VariantResult::score (src/types.rs:335-341), evaluated without FMA contraction
  if (!bp.freq_weight_nonzero) return dist;
  return __ddiv_rn(__dadd_rn(dist, __dmul_rn(bp.freq_weight64, freq)), __dadd_rn(1.0, bp.freq_weight64));
}
// rank_cmp (src/types.rs:344-365) with the gather id as the final key (== stable sort of the
// reference's gather order)
// Branch-free: equal distance scores are the rule (the scores are small-integer quotients), so a short-circuit chain
// sent a handful of lanes down the tie path in nearly every iteration of the rank loop (ncu, eng k = 3: 4.4 of 32 lanes
// on those lines, a third of the kernel's issue slots).  All comparisons are evaluated and combined as predicates.
__device__ __forceinline__ bool ranks_before(const BatchParams& bp, bool gather_order, const SurvRec& a, const SurvRec& b) {
  const bool g_lt = a.g < b.g;
  if (gather_order) return g_lt;
  if (bp.freq_weight_positive) return (a.key > b.key) | ((a.key == b.key) & g_lt);
  return (a.dist > b.dist) | ((a.dist == b.dist) & ((a.freq > b.freq) | ((a.freq == b.freq) & g_lt)));
}

// Triage of one (input a, candidate b) pair for confusable rescoring.
//   CONF_SETTLED : the weight is known -- no pattern can match (necessary condition on the pair's "middle"), or the
//                  middles share no character and the patterns are simple, so the script is =[p]-[ma]+[mb]=[s]
//   CONF_QUEUE   : a pattern may match -> the confusable kernel computes the edit script
//   CONF_HOST    : text outside the BMP or longer than the device edit script handles -> host post-pass
constexpr int CONF_SETTLED = 0, CONF_QUEUE = 1, CONF_HOST = 2;
constexpr uint32_t CONF_FAST_MAX = 8;  // longest middle (characters) the on-the-spot path decodes

__device__ __forceinline__ bool u8_cont(uint32_t byte) { return (byte & 0xC0u) == 0x80u; }
// decodes the BMP characters of s[0, n) (valid UTF-8 of at most three bytes per character) into out[0, cap)
__device__ __forceinline__ uint32_t u8_decode_bmp(const uint8_t* __restrict__ s, uint32_t n, uint16_t* out, uint32_t cap) {
  uint32_t c = 0, i = 0;
  while (i < n && c < cap) {
    const uint32_t b0 = s[i];
    uint32_t cp = b0, l = 1;
    if (b0 >= 0xE0) {
      l = 3;
      cp = b0 & 0x0F;
    } else if (b0 >= 0xC0) {
      l = 2;
      cp = b0 & 0x1F;
    }
    for (uint32_t k = 1; k < l && i + k < n; ++k) cp = (cp << 6) | (s[i + k] & 0x3Fu);
    out[c++] = (uint16_t)cp;
    i += l;
  }
  return c;
}

__device__ __noinline__ int confusable_triage(const DeviceIndex* ix, const uint8_t* __restrict__ a, uint32_t na,
                                              const uint8_t* __restrict__ b, uint32_t nb, double* weight, uint32_t* cost,
                                              bool* wide) {
  *weight = 1.0;
  *cost = 0;
  // characters per string; a four-byte sequence (lead byte >= 0xF0) lies outside the BMP
  uint32_t ca = 0, cb = 0, big = 0, hibits = 0;
#pragma unroll 1
  for (uint32_t i = 0; i < na; ++i) {
    const uint32_t ch = a[i];
    ca += !u8_cont(ch);
    big |= ch >= 0xF0;
    hibits |= ch;
  }
#pragma unroll 1
  for (uint32_t i = 0; i < nb; ++i) {
    const uint32_t ch = b[i];
    cb += !u8_cont(ch);
    big |= ch >= 0xF0;
    hibits |= ch;
  }
  *wide = (hibits & 0x80) != 0;
  if (big) return CONF_HOST;
  // common prefix / suffix in bytes, moved back to character boundaries (the strings agree up to there, so a
  // boundary of one is a boundary of the other)
  uint32_t p = 0;
  const uint32_t m = min(na, nb);
  while (p < m && a[p] == b[p]) ++p;
  while (p > 0 && ((p < na && u8_cont(a[p])) || (p < nb && u8_cont(b[p])))) --p;
  uint32_t s = 0;
  while (s < m - p && a[na - 1 - s] == b[nb - 1 - s]) ++s;
  while (s > 0 && u8_cont(a[na - s])) --s;
  const uint32_t ea = na - s, eb = nb - s;  // the middles: a[p, ea), b[p, eb)
  uint64_t alo = 0, ahi = 0, blo = 0, bhi = 0;
  uint32_t la = 0, lb = 0, wide_a = 0, wide_b = 0;
#pragma unroll 1
  for (uint32_t i = p; i < ea; ++i) {
    const uint32_t ch = a[i];
    la += !u8_cont(ch);
    if (ch >= 0x80) wide_a = 1;
    else if (ch < 64) alo |= 1ull << ch;
    else ahi |= 1ull << (ch - 64);
  }
#pragma unroll 1
  for (uint32_t i = p; i < eb; ++i) {
    const uint32_t ch = b[i];
    lb += !u8_cont(ch);
    if (ch >= 0x80) wide_b = 1;
    else if (ch < 64) blo |= 1ull << ch;
    else bhi |= 1ull << (ch - 64);
  }
  bool possible = false;
#pragma unroll 1
  for (uint32_t k = 0; k < ix->n_conf_pats && !possible; ++k) {
    const ConfPat pat = ix->conf_pats[k];
    bool can = true;
    for (uint32_t q = 0; q < pat.n_instr && can; ++q) {
      const ConfInstr ins = ix->conf_instrs[pat.first_instr + q];
      if (ins.op == 0) continue;  // identities impose nothing on the middle
      const uint64_t slo = ins.op < 0 ? alo : blo, shi = ins.op < 0 ? ahi : bhi;
      const uint32_t swide = ins.op < 0 ? wide_a : wide_b;
      bool any = false;
      for (uint32_t o = 0; o < ins.n_opts && !any; ++o) {
        const ConfOpt opt = ix->conf_opts[ins.first_opt + o];
        any = ((opt.lo & ~slo) | (opt.hi & ~shi)) == 0 && (!opt.nonascii || swide);
      }
      can = any;
    }
    possible = can;
  }
  if (!possible) return CONF_SETTLED;
  if (ca > (uint32_t)esf::MAXLEN || cb > (uint32_t)esf::MAXLEN) return CONF_HOST;
  *cost = la + lb;
  // On the spot: middles without a common character (a lone insertion / deletion only when it is one character: a
  // longer one may be slid over equal neighbours by the clean-up passes, which rotates its text).
  if (!ix->conf_all_simple || la > CONF_FAST_MAX || lb > CONF_FAST_MAX || (la == 0 && lb != 1) || (lb == 0 && la != 1))
    return CONF_QUEUE;
  uint16_t ma[CONF_FAST_MAX], mb[CONF_FAST_MAX];
  u8_decode_bmp(a + p, ea - p, ma, CONF_FAST_MAX);
  u8_decode_bmp(b + p, eb - p, mb, CONF_FAST_MAX);
  for (uint32_t i = 0; i < la; ++i)
    for (uint32_t j = 0; j < lb; ++j)
      if (ma[i] == mb[j]) return CONF_QUEUE;
  // the script: [=prefix] [-ma] [+mb] [=suffix]; simple patterns never look at the text of an identity
  esf::View v[4];
  int nv = 0;
  if (p) v[nv++] = esf::View{esf::EQ, 1, 0, 0};
  if (la) v[nv++] = esf::View{esf::DEL, (uint8_t)la, 0, 0};
  if (lb) v[nv++] = esf::View{esf::INS, (uint8_t)lb, 0, 0};
  if (s) v[nv++] = esf::View{esf::EQ, 1, 0, 0};
  esf::PatTable T;
  T.pats = ix->conf_pats;
  T.instrs = ix->conf_instrs;
  T.opts = ix->conf_opts;
  T.text = ix->conf_text;
  T.n_pats = ix->n_conf_pats;
  double w = 1.0;
#pragma unroll 1
  for (uint32_t k = 0; k < T.n_pats; ++k) {
    const ConfPat pat = T.pats[k];
    if (esf::found_in(T, pat, ma, mb, v, nv)) w = __dmul_rn(w, pat.weight);
  }
  *weight = w;
  return CONF_SETTLED;
}

// What the shared tail does for one query.
constexpr uint32_t RCE_RANK_SCORE = 1;   // normalise frequencies, rank by rank_cmp (src/types.rs:344-365)
constexpr uint32_t RCE_RANK_GATHER = 2;  // keep the gather order (early confusables: ranking follows the rescoring)
constexpr uint32_t RCE_CROP = 4;         // crop at max_matches (src/lib.rs:1536-1589)
constexpr uint32_t RCE_CUTOFF = 8;       // cut-off (src/lib.rs:1598-1622)
constexpr uint32_t RCE_INPLACE = 16;     // write back over the query's own records instead of reserving from the pool
__host__ __device__ inline uint32_t rce_mode(int finish_mode) {
  switch (finish_mode) {
    case FINISH_FULL: return RCE_RANK_SCORE | RCE_CROP | RCE_CUTOFF;
    case FINISH_CROP: return RCE_RANK_SCORE | RCE_CROP;
    case FINISH_GATHER: return RCE_RANK_GATHER;
    default: return 0;  // FINISH_SHARD: survivors pass through unranked
  }
}
// the confusable stage of a launch (null when the model has no confusables or the mode has no post-pass): the
// emitted records remember their query, triage_kernel and confusable_kernel take it from there
struct ConfStage {
  uint32_t* rec_query = nullptr;  // per pool record: the query's row in the batch (indexes the raw-text offsets)
  uint32_t qrow = 0;
};

// The tail shared by score_kernel and merge_kernel: frequency normalisation, ranking, crop, cut-off and
// the packed emission of one query's survivors.  Returns the number of results written (lane 0 only).
__device__ __forceinline__ uint32_t rank_crop_emit(const BatchParams& bp, const uint32_t mode, SurvRec* surv, SurvRec* sorted,
                                                 uint32_t nsurv, double maxfreq, OutRec* __restrict__ out,
                                                 uint32_t* __restrict__ out_gid, OutHead* __restrict__ out_head, uint32_t qi,
                                                 uint32_t flags, uint32_t* __restrict__ qflags, unsigned int* pool_cursor,
                                                 uint32_t inplace_off, const ConfStage& cs) {
  const uint32_t lane = lane_id();
  // ---- normalise frequencies, rank (src/lib.rs:1521-1528) ------------------------------------------
  // (surv keeps the raw frequency in `raw`; `freq` becomes the normalised score, `key` the combined one)
  __threadfence_block();
  __syncwarp();
  if (!(mode & (RCE_RANK_SCORE | RCE_RANK_GATHER))) {
    // lexicon-sharded mode: ranking happens after the exchange (merge_kernel); pass the survivors through
    for (uint32_t i = lane; i < nsurv; i += 32) sorted[i] = surv[i];
  } else {
    const bool gather_order = (mode & RCE_RANK_GATHER) != 0;
    for (uint32_t i = lane; i < nsurv; i += 32) {
      SurvRec r = surv[i];
      if (maxfreq > 0.0) r.freq = __ddiv_rn(r.freq, maxfreq);
      r.key = result_score(bp, r.dist, r.freq);
      surv[i] = r;
    }
    __syncwarp();
    if (nsurv <= 32) {
      for (uint32_t i = lane; i < nsurv; i += 32) {
        const SurvRec a = surv[i];
        uint32_t rank = 0;
#pragma unroll 4
        for (uint32_t j = 0; j < nsurv; ++j) {  // (unrolled by 4: the loads from the scratch list overlap)
          const SurvRec b = surv[j];
          rank += ranks_before(bp, gather_order, b, a) ? 1u : 0u;  // (gather ids are unique: a record never ranks before itself)
        }
        sorted[rank] = a;
      }
    } else {
      // Long lists: counting every record against every other one is n * n / 32 comparisons per lane, and the few
      // queries with hundreds of survivors were most of the kernel.  Instead every block of 32 records is sorted by
      // itself (the same counting, inside the block), and a record's rank is its place in its own block plus, for
      // every other block, the number of records that come before it -- a 6-step binary search in that sorted block
      // (the order is strict and total: gather ids are unique).  n + 6 n / 32 comparisons per lane instead of n * n / 32.
      for (uint32_t b0 = 0; b0 < nsurv; b0 += 32) {
        const uint32_t bn = min(32u, nsurv - b0);
        if (lane < bn) {
          const SurvRec a = surv[b0 + lane];
          uint32_t r = 0;
#pragma unroll 4
          for (uint32_t j = 0; j < bn; ++j) r += ranks_before(bp, gather_order, surv[b0 + j], a) ? 1u : 0u;
          sorted[b0 + r] = a;
        }
      }
      __threadfence_block();
      __syncwarp();
      for (uint32_t i = lane; i < nsurv; i += 32) {
        const SurvRec a = sorted[i];
        const uint32_t mine = i & ~31u;
        uint32_t rank = i - mine;
        for (uint32_t b0 = 0; b0 < nsurv; b0 += 32) {
          if (b0 == mine) continue;
          const uint32_t bn = min(32u, nsurv - b0);
          uint32_t lo = 0;  // records of this block that rank before `a`
#pragma unroll
          for (uint32_t step = 32; step >= 1; step >>= 1) {
            const uint32_t p = lo + step;
            if (p <= bn && ranks_before(bp, gather_order, sorted[b0 + p - 1], a)) lo = p;
          }
          rank += lo;
        }
        surv[rank] = a;  // (the unsorted list is not read any more: it takes the final order)
      }
      SurvRec* t = surv;
      surv = sorted;
      sorted = t;
    }
  }
  __threadfence_block();
  __syncwarp();

  // ---- crop at max_matches with the reference's tie rules (src/lib.rs:1536-1589) ---------------------
  uint32_t n = nsurv;
  if ((mode & RCE_CROP) && bp.max_matches > 0 && n > bp.max_matches) {
    const SurvRec a = sorted[bp.max_matches - 1], b = sorted[bp.max_matches];
    const double last_score = result_score(bp, a.dist, a.freq);
    const double cropped = result_score(bp, b.dist, b.freq);
    if (cropped < last_score) {
      n = bp.max_matches;
    } else {
      // B = first i with dist_i < cropped; E = first i in [1, B) with dist_i == cropped
      uint32_t B = n, E = 0xFFFFFFFFu;
      for (uint32_t b0 = 0; b0 < n && B == n; b0 += 32) {
        const uint32_t i = b0 + lane;
        double dsc = 0.0;
        const bool in = i < n;
        if (in) dsc = sorted[i].dist;
        const uint32_t lt = __ballot_sync(FULL, in && dsc < cropped);
        uint32_t eq = __ballot_sync(FULL, in && i >= 1 && dsc == cropped);
        if (lt) {
          const uint32_t first = __ffs(lt) - 1;
          B = b0 + first;
          eq &= (first == 0) ? 0u : (0xFFFFFFFFu >> (32 - first));
        }
        if (eq && E == 0xFFFFFFFFu) E = b0 + __ffs(eq) - 1;
      }
      if (E != 0xFFFFFFFFu)
        n = E + 1;
      else if (B < n && B > 0)
        n = B + 1;
    }
  }
  // ---- cut-off (src/lib.rs:1598-1622); after the confusable rescoring when there is one --------------
  if ((mode & RCE_CUTOFF) && bp.cutoff_threshold >= 1.0 && n > 1) {
    const SurvRec a = sorted[0];
    const double lim = __ddiv_rn(result_score(bp, a.dist, a.freq), bp.cutoff_threshold);
    uint32_t cut = n;
    for (uint32_t b0 = 0; b0 < n && cut == n; b0 += 32) {
      const uint32_t i = b0 + lane;
      bool hit = false;
      if (i >= 1 && i < n) {
        const SurvRec r = sorted[i];
        hit = result_score(bp, r.dist, r.freq) <= lim;
      }
      const uint32_t m = __ballot_sync(FULL, hit);
      if (m) cut = b0 + __ffs(m) - 1;
    }
    n = cut;
  }
  // ---- emit into the packed pool (one atomic reservation per query) --------------------------------------
  uint32_t off = inplace_off;
  bool fits = true;
  if (!(mode & RCE_INPLACE)) {
    off = 0;
    if (lane == 0 && n > 0) off = atomicAdd(pool_cursor, n);
    off = __shfl_sync(FULL, off, 0);
    fits = (unsigned long long)off + n <= (unsigned long long)bp.pool_cap;
  }
  if (fits) {
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
      const uint32_t i = i0 + lane;
      if (i < n) {
        const SurvRec r = sorted[i];
        OutRec o;
        o.dist_score = r.dist;
        o.vocab_id = r.vocab;
        o.freq = r.raw;
        out[off + i] = o;
        if (out_gid) out_gid[off + i] = r.g;
        if (cs.rec_query) cs.rec_query[off + i] = cs.qrow;
      }
    }
  }
  if (lane == 0) {
    OutHead h;
    h.max_freq = maxfreq;
    h.offset = off;
    h.count = fits ? n : 0;
    out_head[qi] = h;
    const uint32_t nf = (flags & ~QF_OUT_OVERFLOW) | (fits ? 0u : QF_OUT_OVERFLOW);
    if (nf != flags) qflags[qi] = nf;
  }
  __syncwarp();
  return (lane == 0 && fits) ? n : 0;
}

// ================================================================================================
// Kernel 2a: prefilter -- bit-parallel restricted (OSA) Damerau-Levenshtein distance
// ================================================================================================
// Most candidates of the anagram neighbourhood are far beyond the edit-distance threshold.  Hyyro's
// bit-vector algorithm gives each lane the OSA distance of its candidate in ~30 instructions per candidate
// symbol (against ~30 per matrix CELL for the exact DP of the score kernel).  OSA and the true
// (Lowrance-Wagner) distance DL differ only through transpositions with a gap: such an operation costs
// c >= 2 in DL and c + 1 when replaced by plain edits, so OSA <= DL + floor(DL / 2).  Hence
// OSA > ke + floor(ke / 2) implies DL > ke: the candidate is dropped here (exact) and the hit list is
// compacted in place; everything else goes through the exact DP.  Only queries with more than one batch of
// candidates (it can save a whole batch of the DP) and at most 32 symbols (one 32-bit word) are filtered;
// they get QF_PREFILTERED.  A kernel of its own: a tiny loop and 1.3 KB of shared memory per warp, so it runs
// at full occupancy instead of sharing the instruction cache and the occupancy of the score kernel.
constexpr int KF_WARPS = 8;
__global__ void __launch_bounds__(KF_WARPS * 32)
prefilter_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
                 const uint32_t* __restrict__ qlist, uint32_t nq, uint32_t* hits, uint32_t* hit_count, uint32_t* qflags,
                 unsigned int* work, Counters* counters) {
  __shared__ uint32_t pm_s[KF_WARPS][256];  // per warp: bit j of pm[c] set iff query symbol j equals c
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t pm_a = (uint32_t)__cvta_generic_to_shared(&pm_s[warp][0]);
  for (uint32_t k = lane; k < 256; k += 32) sts_u32(pm_a + k * 4, 0);
  __syncwarp();
  const uint8_t* __restrict__ rows = ix->inst_rows;
  const uint32_t nstride = ix->norm_stride;
  unsigned long long c_pairs = 0, c_cells = 0;
  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(work, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= nq) break;
    const uint32_t flags = qflags[qi];
    if (flags & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED | QF_PREFILTERED)) continue;
    uint32_t nh = hit_count[qi];
    const uint32_t q = qlist ? qlist[qi] : qi;
    const uint8_t* qrow = queries + (size_t)q * bp.query_stride;
    const uint32_t Lq = qrow[0];
    if (Lq > 32 || nh <= 32) continue;
    const uint32_t ke = apply_threshold(bp.max_edit, Lq);
    uint32_t* hq = hits + (size_t)qi * bp.hit_cap;
    uint32_t mysym = 256u + lane;  // (lanes beyond the query get unique values: they match nobody)
    {
      if (lane < Lq) mysym = qrow[2 + lane];
      const uint32_t mm = __match_any_sync(FULL, mysym);  // the lanes (= query positions) holding the same symbol
      if (lane < Lq) sts_u32(pm_a + mysym * 4, mm);
      __syncwarp();
      const uint32_t top = 1u << (Lq - 1);
      const uint32_t osa_max = ke + ke / 2;
      uint32_t w = 0;
      for (uint32_t hb = 0; hb < nh; hb += 32) {
        const uint32_t hi = hb + lane;
        bool valid = hi < nh;
        uint32_t g = 0, Lc = 0;
        const uint8_t* row = rows;
        uint4 v0 = make_uint4(0, 0, 0, 0);
        if (valid) {
          g = hq[hi];
          row = rows + (size_t)g * nstride;
          v0 = __ldg(reinterpret_cast<const uint4*>(row));
          Lc = v0.x & 0xFF;
          const uint32_t diff = Lq > Lc ? Lq - Lc : Lc - Lq;
          valid = diff <= ke;  // length pre-check of damerau_levenshtein (src/distance.rs:109-130)
          if (valid) {
            c_pairs += 1;
            c_cells += (unsigned long long)Lq * Lc;
          } else {
            Lc = 0;
          }
        }
        const uint32_t Lcm = __reduce_max_sync(FULL, Lc);
        uint32_t D0 = 0, VP = 0xFFFFFFFFu, VN = 0, PMp = 0, sc = Lq;
        const uint32_t nbytes = Lcm ? Lcm + 2 : 0;  // row bytes to walk: len, flags, symbols
        for (uint32_t k0 = 0; k0 < nbytes; k0 += 16) {
          uint4 v = make_uint4(0, 0, 0, 0);
          if (k0 < Lc + 2 && Lc) v = (k0 == 0) ? v0 : __ldg(reinterpret_cast<const uint4*>(row + k0));
          uint32_t x = v.x, y = v.y, z = v.z, t = v.w;
          uint32_t pos = k0, steps = min(16u, nbytes - k0);
          if (k0 == 0) {  // skip the length and flag bytes
            x = __funnelshift_r(x, y, 16);
            y = __funnelshift_r(y, z, 16);
            z = __funnelshift_r(z, t, 16);
            t >>= 16;
            pos = 2;
            steps -= 2;
          }
          for (uint32_t st = 0; st < steps; ++st, ++pos) {
            const uint32_t c = x & 0xFFu;
            x = __funnelshift_r(x, y, 8);
            y = __funnelshift_r(y, z, 8);
            z = __funnelshift_r(z, t, 8);
            t >>= 8;
            if (pos < Lc + 2) {  // this lane's candidate still has symbols (never for dropped lanes: Lc = 0)
              const uint32_t PMj = lds_u32(pm_a + c * 4);
              const uint32_t TR = ((~D0 & PMj) << 1) & PMp;  // adjacent transposition
              D0 = TR | (((PMj & VP) + VP) ^ VP) | PMj | VN;
              const uint32_t HP = VN | ~(D0 | VP);
              const uint32_t HN = D0 & VP;
              sc += (HP & top) ? 1u : 0u;
              sc -= (HN & top) ? 1u : 0u;
              const uint32_t X = (HP << 1) | 1u;
              VP = (HN << 1) | ~(D0 | X);
              VN = X & D0;
              PMp = PMj;
            }
          }
        }
        const bool pass = valid && sc <= osa_max;
        const uint32_t pmask = __ballot_sync(FULL, pass);
        __syncwarp();  // every lane has read its entry of this batch: the compacted list may overwrite it
        if (pass) hq[w + __popc(pmask & lanemask_lt())] = g;
        w += __popc(pmask);
      }
      __syncwarp();
      if (lane < Lq) sts_u32(pm_a + mysym * 4, 0);  // leave the table clean for the next query
      nh = w;
      if (lane == 0) hit_count[qi] = w;  // a re-run of this kernel (pool overflow) must see the filtered list
      __syncwarp();
    }
    if (lane == 0) qflags[qi] = flags | QF_PREFILTERED;
  }
  if (counters) {
    for (int o = 16; o > 0; o >>= 1) {
      c_pairs += __shfl_xor_sync(FULL, c_pairs, o);
      c_cells += __shfl_xor_sync(FULL, c_cells, o);
    }
    if (lane == 0) {
      atomicAdd(&counters->dl_pairs, c_pairs);
      atomicAdd(&counters->dl_cells, c_cells);
    }
  }
}

constexpr uint32_t K2_HEAVY_HITS = 64;  // more candidates than two rounds of 32: scored in the first phase
#ifndef ANL_K2_MIN_CTAS
#define ANL_K2_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(K2_WARPS * 32, ANL_K2_MIN_CTAS)
score_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
             const uint32_t* __restrict__ qlist, uint32_t* __restrict__ rec_query, uint32_t nq, uint32_t* hits, uint32_t* hit_count,
             uint32_t* __restrict__ qflags, OutRec* __restrict__ out,
             uint32_t* __restrict__ out_gid, OutHead* __restrict__ out_head, SurvRec* __restrict__ scratch, unsigned int* work,
             unsigned int* work_light, unsigned int* pool_cursor, Counters* counters, uint32_t ML, uint32_t R, uint32_t need_min,
             uint32_t need_max, uint32_t scratch_cta0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  // per-warp shared memory by shared-window address: sq (query symbols), cell[(ML+1)][32] words, ring[R][(ML+1)][32] bytes
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + warp * (uint32_t)k2_warp_bytes(ML, R);
  const uint32_t sq_a = sbase;
  const uint32_t cell_a = sbase + 256 + lane * 4;                    // + j * 128
  const uint32_t ring_a = sbase + 256 + (ML + 1) * 32 * 4 + lane;    // + slot * rowbytes + j * 32
  const uint32_t rowbytes = (ML + 1) * 32;
  const uint32_t max_len = ix->max_len;

  const uint32_t gwarp = (blockIdx.x + scratch_cta0) * K2_WARPS + warp;  // (scratch_cta0: first scratch slot of this launch)
  SurvRec* surv = scratch + (size_t)gwarp * 2 * bp.hit_cap;  // survivors, then the sorted copy
  SurvRec* sorted = surv + bp.hit_cap;

  const uint8_t* __restrict__ rows = ix->inst_rows;
  const uint32_t nstride = ix->norm_stride;
  const int have_freq = ix->have_freq;
  const uint32_t* __restrict__ gid_of = ix->inst_gid;
  unsigned long long c_pairs = 0, c_cells = 0, c_surv = 0, c_res = 0, c_dpp = 0, c_dpc = 0;

  // Queries are split over launches by the number of matrix columns they can need (longest admissible
  // candidate = min(longest entry, query length + max edit distance)): the launch for short queries gets by
  // with a fraction of the shared memory, i.e. more resident warps.  The launch for the long ones takes 32
  // queries per counter increment and tests their class one per lane (it skips nearly all of them).
  // Heavy queries first: a query is one warp's job from start to end, so the kernel cannot end before its
  // heaviest query does; started last, a query with a thousand candidates is the whole tail of a 65 k-query
  // chunk.  Phase 0 walks the batch four queries per grab and takes those with more than two rounds of
  // candidates, phase 1 (its own counter) takes the rest one query per grab.
  uint32_t gbase = 0, gmask = 0, phase = 0;
  for (;;) {
    if (gmask == 0) {
      const uint32_t grab = need_min > 0 ? 32u : (phase == 0 ? 4u : 1u);  // (few per grab: a warp works its grab off serially)
      if (lane == 0) gbase = atomicAdd(phase == 0 ? work : work_light, grab);
      gbase = __shfl_sync(FULL, gbase, 0);
      if (gbase >= nq) {
        if (phase == 0) {
          phase = 1;
          continue;
        }
        break;
      }
      bool mine = false;
      if (lane < grab && gbase + lane < nq) {
        const uint32_t f0 = qflags[gbase + lane];
        const uint32_t q0 = qlist ? qlist[gbase + lane] : gbase + lane;
        const uint32_t L0 = queries[(size_t)q0 * bp.query_stride];
        const bool skip0 = (f0 & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED)) != 0;
        const uint32_t need = skip0 ? 0u : min(max_len, L0 + apply_threshold(bp.max_edit, L0));
        const bool heavy = !skip0 && hit_count[gbase + lane] > K2_HEAVY_HITS;
        mine = need >= need_min && need <= need_max && heavy == (phase == 0);
      }
      gmask = __ballot_sync(FULL, mine);
      if (gmask == 0) continue;
    }
    const uint32_t qi = gbase + (uint32_t)__ffs(gmask) - 1;
    gmask &= gmask - 1;
    const uint32_t flags = qflags[qi];
    const bool skip = (flags & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED)) != 0;
    if (skip) {
      if (lane == 0) {
        OutHead h;
        h.max_freq = 0.0;
        h.offset = 0;
        h.count = 0;
        out_head[qi] = h;
      }
      continue;
    }
    const uint32_t q = qlist ? qlist[qi] : qi;
    const uint8_t* qrow = queries + (size_t)q * bp.query_stride;
    const uint32_t Lq = qrow[0];
    const bool q_lower = (qrow[1] & Q_FIRST_LOWER) != 0;
    for (uint32_t i = lane; i < Lq; i += 32) sts_u8(sq_a + i, qrow[2 + i]);
    __syncwarp();
    const uint32_t ke = apply_threshold(bp.max_edit, Lq);
    const uint32_t nh = hit_count[qi];
    const uint32_t* hq = hits + (size_t)qi * bp.hit_cap;
    const double Ld = (double)Lq;

    // (candidates far beyond the edit distance were already dropped by prefilter_kernel when the query has more
    // than one batch of them; it also counted the pairs of those queries)
    const bool prefilter = (flags & QF_PREFILTERED) != 0;
    // every feature of the score is a small integer divided by the query length: lane v holds v / Ld once per
    // query and the per-candidate quotients are fetched by shuffle (same IEEE division, so the bits are the same)
    const double quot_lane = __ddiv_rn((double)lane, Ld);
    const bool quot_ok = Lq <= 31;

    uint32_t nsurv = 0;
    double maxfreq = 0.0;
#ifdef ANL_ROUND_STATS
    uint32_t nvalid_q = 0;
#endif

    for (uint32_t hb = 0; hb < nh; hb += 32) {
      const uint32_t hi = hb + lane;
      bool valid = hi < nh;
      uint32_t g = 0, Lc = 0;
      bool c_lower = false;
      if (valid) {
        g = hq[hi];
        const uint8_t* row = rows + (size_t)g * nstride;
        const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(row));
        Lc = v0.x & 0xFF;
        c_lower = ((v0.x >> 8) & ROW_FIRST_LOWER) != 0;
        // length pre-check of damerau_levenshtein (src/distance.rs:109-130)
        const uint32_t diff = Lq > Lc ? Lq - Lc : Lc - Lq;
        valid = diff <= ke;
        if (valid) {
          // stage the candidate's symbols: cell[j].t = t[j-1]
          for (uint32_t j0 = 0; j0 < Lc + 2; j0 += 16) {
            const uint4 v = (j0 == 0) ? v0 : __ldg(reinterpret_cast<const uint4*>(row + j0));
            // (a rolled loop over a shifting 128-bit window: staging runs once per batch, and the unrolled form
            // cost far more instruction-cache footprint than it saved in issue slots)
            uint32_t x = v.x, y = v.y, z = v.z, t = v.w;
            const uint32_t jend = min(j0 + 16, Lc + 2);
#pragma unroll 1
            for (uint32_t bytepos = j0; bytepos < jend; ++bytepos) {  // byte in the row; symbol index = bytepos - 2
              const uint32_t sym = x & 0xFFu;
              x = __funnelshift_r(x, y, 8);
              y = __funnelshift_r(y, z, 8);
              z = __funnelshift_r(z, t, 8);
              t >>= 8;
              // column j = bytepos - 1: {t, lcs = 0, lastrow = 0, D[0][j] = j}
              if (bytepos >= 2) sts_u32(cell_a + (bytepos - 1) * 128, sym | ((bytepos - 1) << 24));
            }
          }
          if (!prefilter) {  // (the prefilter pass counted them already)
            c_pairs += 1;
            c_cells += (unsigned long long)Lq * Lc;
          }
        }
      }
      const uint32_t vmask = __ballot_sync(FULL, valid);
      if (vmask == 0) continue;
      const uint32_t Lcm = __reduce_max_sync(FULL, valid ? Lc : 0u);
#ifdef ANL_ROUND_STATS  // experiment: dp_pairs = DP rounds executed, dp_cells = rounds the within-distance candidates alone would need
      c_dpp += lane == 0 ? 1 : 0;
#else
      c_dpp += valid ? 1 : 0;
      c_dpc += (unsigned long long)Lq * Lcm;  // per lane: x 32 lanes in the sum = warp-cells of this batch
#endif

      // ---- true Damerau-Levenshtein, all lanes in lock-step over (i, j) -----------------------
      // Row i of the matrix lives in ring slot (i mod R), R = max edit distance + 2: only the last
      // ke+2 rows are ever needed, because a transposition reaching further back costs more than
      // ke ("DP kernel").  Row and column numbers are shifted by S = ke + 2 so that
      // "no previous occurrence" (0) fails the reach test below without a separate check.
      const uint32_t S = ke + 2;
      if (!valid) Lc = 0;
      for (uint32_t j = Lc + 1; j <= Lcm; ++j) sts_u32(cell_a + j * 128, 0xFFu | (j << 24));  // sentinel symbol: never equal
      for (uint32_t j = 0; j <= Lcm; ++j) sts_u8(ring_a + j * 32, j);                        // row 0 in slot 0
      uint32_t lcs_best = 0;
      uint32_t slot = 0;  // ring slot of row i - 1
      for (uint32_t i = 1; i <= Lq; ++i) {
        const uint32_t sc = lds_u8(sq_a + i - 1);
        slot = slot + 1 == R ? 0 : slot + 1;
        const uint32_t cur_a = ring_a + slot * rowbytes;
        const uint32_t is = i + S;  // shifted row number
        uint32_t left = i, diag = i - 1, db = 0, lcs_diag = 0;
        sts_u8(cur_a, i);
        for (uint32_t j = 1; j <= Lcm; ++j) {
          const uint32_t cw = lds_u32(cell_a + j * 128);
          const uint32_t tc = cw & 0xFF, lcs_up = (cw >> 8) & 0xFF, last = (cw >> 16) & 0xFF, up = cw >> 24;
          const bool same = tc == sc;
          const uint32_t js = j + S;
          uint32_t v = min(min(left, up) + 1, diag + (same ? 0u : 1u));
          // transposition (src/distance.rs:160-165): mat[last][db] + (i-last-1) + 1 + (j-db-1); a term that
          // reaches back more than ke + 1 in total cannot be <= ke and is skipped (exact)
          const uint32_t reach = (is - last) + (js - db);
          if (reach <= ke + 1) {
            const uint32_t back = is - last + 1;  // rows between row i and row last-1
            const uint32_t ts = slot >= back ? slot - back : slot + R - back;
            const uint32_t tv = lds_u8(ring_a + ts * rowbytes + (db - S - 1) * 32) + reach - 1;
            v = min(v, tv);
          }
          sts_u8(cur_a + j * 32, v);
          const uint32_t lcs_new = same ? lcs_diag + 1 : 0;
          lcs_best = max(lcs_best, lcs_new);
          sts_u32(cell_a + j * 128, tc | (lcs_new << 8) | ((same ? is : last) << 16) | (v << 24));
          if (same) db = js;
          lcs_diag = lcs_up;
          diag = up;
          left = v;
        }
      }
      uint32_t ld = 255;
      if (valid) ld = lds_u8(ring_a + slot * rowbytes + Lc * 32);
      valid = valid && ld <= ke;

      // ---- prefix / suffix (src/distance.rs:208-231) ------------------------------------------------
      uint32_t pre = 0, suf = 0;
      {
        const uint32_t lim = min(Lq, Lcm);
        bool pgo = valid, sgo = valid;
        for (uint32_t i = 0; i < lim; ++i) {
          if (valid && i < Lc) {
            const uint32_t a = lds_u32(cell_a + (i + 1) * 128) & 0xFF;
            pgo = pgo && (a == lds_u8(sq_a + i));
            pre += pgo;
            const uint32_t b = lds_u32(cell_a + (Lc - i) * 128) & 0xFF;
            sgo = sgo && (b == lds_u8(sq_a + Lq - 1 - i));
            suf += sgo;
          }
        }
      }
      // features are skipped (0 / true) when their weight is <= 0 (src/lib.rs:1352-1377)
      const uint32_t f_lcs = bp.w_lcs > 0.0 ? lcs_best : 0;
      const uint32_t f_pre = bp.w_prefix > 0.0 ? pre : 0;
      const uint32_t f_suf = bp.w_suffix > 0.0 ? suf : 0;
      const bool samecase = bp.w_case > 0.0 ? (c_lower == q_lower) : true;

      // ---- f64 score, left to right, no FMA (src/lib.rs:1433-1452) -----------------------------------
      double q_ld, q_lcs, q_pre, q_suf;
      if (quot_ok) {  // warp-uniform; all four indices are <= Lq <= 31 (ld is clamped: it only matters when <= Lq)
        q_ld = __shfl_sync(FULL, quot_lane, min(ld, 31u));
        q_lcs = __shfl_sync(FULL, quot_lane, f_lcs);
        q_pre = __shfl_sync(FULL, quot_lane, f_pre);
        q_suf = __shfl_sync(FULL, quot_lane, f_suf);
      } else {
        q_ld = __ddiv_rn((double)ld, Ld);
        q_lcs = __ddiv_rn((double)f_lcs, Ld);
        q_pre = __ddiv_rn((double)f_pre, Ld);
        q_suf = __ddiv_rn((double)f_suf, Ld);
      }
      const double ds = ld > Lq ? 0.0 : __dsub_rn(1.0, q_ld);
      double acc = __dmul_rn(bp.w_ld, ds);
      acc = __dadd_rn(acc, __dmul_rn(bp.w_lcs, q_lcs));
      acc = __dadd_rn(acc, __dmul_rn(bp.w_prefix, q_pre));
      acc = __dadd_rn(acc, __dmul_rn(bp.w_suffix, q_suf));
      acc = __dadd_rn(acc, samecase ? bp.w_case : 0.0);
      const double score = __ddiv_rn(acc, bp.w_sum);
      double freq = 1.0;
      if (valid && have_freq) freq = (double)__ldg(ix->inst_freq + g);
      // max_freq is taken over every instance within the edit distance, before the score threshold
      double mf = valid ? freq : 0.0;
      for (int o = 16; o > 0; o >>= 1) mf = fmax(mf, __shfl_xor_sync(FULL, mf, o));
      maxfreq = fmax(maxfreq, mf);
      c_surv += valid ? 1 : 0;
#ifdef ANL_ROUND_STATS
      nvalid_q += __popc(__ballot_sync(FULL, valid));
#endif
      const bool keep = valid && score >= bp.score_threshold;
      const uint32_t kmask = __ballot_sync(FULL, keep);
      if (keep) {
        SurvRec r;
        r.dist = score;
        r.freq = freq;
        r.key = 0.0;
        r.g = gid_of ? __ldg(gid_of + g) : g;  // sharded index: global gather id
        r.raw = (uint32_t)freq;               // raw frequency (exact: u32 or 1.0)
        r.vocab = __ldg(ix->inst_vocab + g);
        r.pad = 0;
        surv[nsurv + __popc(kmask & lanemask_lt())] = r;
      }
      nsurv += __popc(kmask);
      __syncwarp();
    }

#ifdef ANL_ROUND_STATS
    c_dpc += lane == 0 ? (nvalid_q + 31) / 32 : 0;
#endif
    ConfStage cs;
    if (rec_query && (bp.finish_mode == FINISH_CROP || bp.finish_mode == FINISH_GATHER)) {
      cs.rec_query = rec_query;
      cs.qrow = q;
    }
    c_res += rank_crop_emit(bp, rce_mode(bp.finish_mode), surv, sorted, nsurv, maxfreq, out, out_gid, out_head, qi, flags,
                            qflags, pool_cursor, 0, cs);
  }

  if (counters) {
    unsigned long long v[6] = {c_pairs, c_cells, c_surv, c_res, c_dpp, c_dpc};
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
      unsigned long long x = v[k];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
      v[k] = x;
    }
    if (lane == 0) {
      atomicAdd(&counters->dl_pairs, v[0]);
      atomicAdd(&counters->dl_cells, v[1]);
      atomicAdd(&counters->survivors, v[2]);
      atomicAdd(&counters->results, v[3]);
      atomicAdd(&counters->dp_pairs, v[4]);
      atomicAdd(&counters->dp_cells, v[5]);
    }
  }
}


// ================================================================================================
// Kernel 3 (lexicon-sharded mode only): merge of the per-shard survivor lists after the exchange
// ================================================================================================
// Every shard scored the whole query batch against its part of the index and exported, per query,
// its survivors (distance score, raw frequency, vocabulary id, GLOBAL gather id) and its local
// max_freq.  After the all-gather each rank holds all G exports; one warp per query concatenates the
// G lists, takes the global max_freq (frequency normalisation is global, src/lib.rs:1460,1521-1525)
// and runs the same rank / crop / cut-off tail as the unsharded kernel.
__global__ void __launch_bounds__(K2_WARPS * 32)
merge_kernel(const BatchParams bp, uint32_t nq, uint32_t n_shards, const OutHead* __restrict__ heads_all, uint32_t head_stride,
             const OutRec* __restrict__ recs_all, const uint32_t* __restrict__ gids_all, uint32_t rec_stride,
             const uint32_t* __restrict__ qflags_in, uint32_t* __restrict__ qflags, OutRec* __restrict__ out,
             OutHead* __restrict__ out_head, SurvRec* __restrict__ scratch, uint32_t scratch_cap, unsigned int* work,
             unsigned int* pool_cursor) {
  const uint32_t lane = lane_id();
  const uint32_t gwarp = blockIdx.x * K2_WARPS + (threadIdx.x >> 5);
  SurvRec* surv = scratch + (size_t)gwarp * 2 * scratch_cap;
  SurvRec* sorted = surv + scratch_cap;
  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(work, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= nq) break;
    uint32_t nsurv = 0;
    double maxfreq = 0.0;
    const uint32_t flags = qflags_in[qi];
    // the shards' headers in one go (lane r holds shard r's; 32 shards per round), then the lists one after the other
    for (uint32_t r0 = 0; r0 < n_shards; r0 += 32) {
      OutHead mine;
      mine.max_freq = 0.0;
      mine.offset = 0;
      mine.count = 0;
      if (r0 + lane < n_shards) mine = heads_all[(size_t)(r0 + lane) * head_stride + qi];
      double mf = mine.max_freq;
      for (int o = 16; o > 0; o >>= 1) mf = fmax(mf, __shfl_xor_sync(FULL, mf, o));
      maxfreq = fmax(maxfreq, mf);
      const uint32_t in_round = min(32u, n_shards - r0);
      for (uint32_t k = 0; k < in_round; ++k) {
        const uint32_t cnt = __shfl_sync(FULL, mine.count, k), off = __shfl_sync(FULL, mine.offset, k);
        const OutRec* recs = recs_all + (size_t)(r0 + k) * rec_stride + off;
        const uint32_t* gids = gids_all + (size_t)(r0 + k) * rec_stride + off;
        for (uint32_t i = lane; i < cnt; i += 32) {
          if (nsurv + i < scratch_cap) {
            const OutRec o = recs[i];
            SurvRec s;
            s.dist = o.dist_score;
            s.freq = (double)o.freq;
            s.key = 0.0;
            s.g = gids[i];
            s.raw = o.freq;
            s.vocab = o.vocab_id;
            s.pad = 0;
            surv[nsurv + i] = s;
          }
        }
        nsurv += cnt;
      }
    }
    __syncwarp();
    if (nsurv > scratch_cap) {  // cannot happen: the host sizes scratch_cap from the gathered counts
      if (lane == 0) {
        OutHead h;
        h.max_freq = maxfreq;
        h.offset = 0;
        h.count = 0;
        out_head[qi] = h;
        qflags[qi] = flags | QF_OUT_OVERFLOW;
      }
      continue;
    }
    rank_crop_emit(bp, rce_mode(bp.finish_mode), surv, sorted, nsurv, maxfreq, out, nullptr, out_head, qi, flags, qflags,
                   pool_cursor, 0, ConfStage());
  }
}

// ================================================================================================
// Kernels 4 + 5 (only with confusables): device-side rescoring of the ranked lists
// ================================================================================================
// triage_kernel: one pool record per thread (the score stage left each record's query in rec_query).  Settles the
// pairs whose confusable weight is known without an edit script and queues the rest: pure-ASCII pairs from the front
// of the work list (confusable_kernel, over bytes), pairs with other BMP characters from its back
// (confusable_wide_kernel, over UTF-16 code units).  A kernel of its own: inside the score kernel the triage ran on
// the few lanes that hold a query's results (6 of 32 on cfg 2) and walked the candidate's text byte by byte at that
// occupancy; here every lane has a record.
__global__ void __launch_bounds__(256)
triage_kernel(const DeviceIndex* __restrict__ ix, const uint8_t* __restrict__ qblob, const uint32_t* __restrict__ qboff,
              const uint32_t* __restrict__ rec_query, OutRec* __restrict__ out, const unsigned int* __restrict__ pool_cursor,
              uint32_t pool_cap, ConfWork* __restrict__ worklist, unsigned int* work_cursor, unsigned int* wide_cursor) {
  const uint32_t used = *pool_cursor;
  if (used > pool_cap) return;  // pool overflow: the score stage runs again with a larger pool
  const uint8_t* __restrict__ vtext = ix->vocab_text;
  const uint32_t* __restrict__ voff = ix->vocab_text_off;
  const uint32_t lane = lane_id();
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < used; base += gridDim.x * blockDim.x) {
    const uint32_t rec = base + lane;
    bool queue = false, wide = false;
    uint32_t cost = 0, query = 0;
    if (rec < used) {
      OutRec r = out[rec];
      query = rec_query[rec];
      const uint32_t vocab = r.vocab_id & ~OUT_SKIP_CONFUSABLES;
      const uint32_t a0 = qboff[query], a1 = qboff[query + 1];
      const uint32_t b0 = __ldg(voff + vocab), b1 = __ldg(voff + vocab + 1);
      double w;
      const int tri = confusable_triage(ix, qblob + a0, a1 - a0, vtext + b0, b1 - b0, &w, &cost, &wide);
      // Without a work list the host runs the post-pass (variant lists, sharded mode): it only wants to know which
      // records provably keep their score -- a weight found on the spot is left for it to apply.
      if (tri == CONF_SETTLED && (w == 1.0 || worklist)) {
        r.vocab_id |= OUT_SKIP_CONFUSABLES;
        if (w != 1.0) r.dist_score = __dmul_rn(r.dist_score, w);  // src/lib.rs:1660
        out[rec] = r;
      }
      queue = tri == CONF_QUEUE;
    }
    if (worklist) {
      // the two queues share the array (capacity = pool capacity >= records emitted): bytes from the front, wide from the back
      const uint32_t qa = __ballot_sync(FULL, queue && !wide), qw = __ballot_sync(FULL, queue && wide);
      if (qa | qw) {
        uint32_t abase = 0, wbase = 0;
        if (lane == 0) {
          if (qa) abase = atomicAdd(work_cursor, (unsigned int)__popc(qa));
          if (qw) wbase = atomicAdd(wide_cursor, (unsigned int)__popc(qw));
        }
        abase = __shfl_sync(FULL, abase, 0);
        wbase = __shfl_sync(FULL, wbase, 0);
        if (queue) {
          ConfWork w;
          w.rec = rec;
          w.query = query;
          w.cost = cost;
          w.pad = 0;
          const uint32_t pos = wide ? pool_cap - 1 - (wbase + __popc(qw & lanemask_lt())) : abase + __popc(qa & lanemask_lt());
          worklist[pos] = w;
        }
      }
    }
  }
}

// confusable_kernel: one queued (input, candidate) pair per thread.  Computes the edit script of the raw
// strings and the product of the weights of all patterns found in it (rescore_confusables /
// compute_confusable_weight, src/lib.rs:1656-1663, 1733-1756), multiplies the record's distance score and marks it
// settled.  Pairs outside the limits of editscript_fixed.h stay unsettled (host post-pass).
__global__ void __launch_bounds__(64)
confusable_kernel(const DeviceIndex* __restrict__ ix, const uint8_t* __restrict__ qblob, const uint32_t* __restrict__ qboff,
                  const ConfWork* __restrict__ worklist, const unsigned int* __restrict__ work_count, uint32_t work_cap,
                  OutRec* __restrict__ out) {
  const uint32_t total = min(*work_count, work_cap);
  esf::PatTable T;
  T.pats = ix->conf_pats;
  T.instrs = ix->conf_instrs;
  T.opts = ix->conf_opts;
  T.text = ix->conf_text;
  T.n_pats = ix->n_conf_pats;
  const uint8_t* __restrict__ vtext = ix->vocab_text;
  const uint32_t* __restrict__ voff = ix->vocab_text_off;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
    const ConfWork it = worklist[w];
    const OutRec r = out[it.rec];
    const uint32_t vocab = r.vocab_id & ~OUT_SKIP_CONFUSABLES;
    const uint32_t a0 = qboff[it.query], a1 = qboff[it.query + 1];
    const uint32_t b0 = __ldg(voff + vocab), b1 = __ldg(voff + vocab + 1);
    // private copies: the diff touches every character many times
    uint8_t a[esf::MAXLEN], b[esf::MAXLEN];
    const int na = (int)(a1 - a0), nb = (int)(b1 - b0);
    if (na > esf::MAXLEN || nb > esf::MAXLEN) continue;
    for (int i = 0; i < na; ++i) a[i] = qblob[a0 + i];
    for (int i = 0; i < nb; ++i) b[i] = vtext[b0 + i];
    esf::View v[esf::MAXSEG];
    const int nv = esf::shortest_edit_script(a, na, b, nb, v);
    if (nv < 0) continue;
    double weight = 1.0;
    for (uint32_t k = 0; k < T.n_pats; ++k) {
      const ConfPat pat = T.pats[k];
      if (esf::found_in(T, pat, a, b, v, nv)) weight = __dmul_rn(weight, pat.weight);
    }
    OutRec o = r;
    if (weight != 1.0) o.dist_score = __dmul_rn(r.dist_score, weight);
    o.vocab_id = r.vocab_id | OUT_SKIP_CONFUSABLES;
    out[it.rec] = o;
  }
}

// The same for the pairs with characters beyond ASCII (the back of the work list): strings decoded to UTF-16 code
// units -- every character of the Basic Multilingual Plane is one unit -- and the Unicode Alphabetic ranges for the
// boundary scores of the diff clean-up (cf. UnicodeClass in editscript.cpp).
__constant__ uint32_t c_alpha_ranges[2 * 800];
__constant__ uint32_t c_n_alpha_ranges;
struct DeviceCharClass {
  static __device__ __forceinline__ bool alnum(uint32_t c) {
    if (c < 0x80) return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
    uint32_t lo = 0, hi = c_n_alpha_ranges;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (c > c_alpha_ranges[2 * mid + 1]) lo = mid + 1; else hi = mid;
    }
    return lo < c_n_alpha_ranges && c >= c_alpha_ranges[2 * lo];
  }
  static __device__ __forceinline__ bool space(uint32_t c) {
    return c == ' ' || (c >= 9 && c <= 13) || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
           c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
  }
};
cudaError_t upload_alphabetic_ranges(const uint32_t* ranges, uint32_t n) {
  if (n > 800) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemcpyToSymbol(c_alpha_ranges, ranges, (size_t)n * 2 * sizeof(uint32_t));
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(c_n_alpha_ranges, &n, sizeof n);
}
__global__ void __launch_bounds__(64)
confusable_wide_kernel(const DeviceIndex* __restrict__ ix, const uint8_t* __restrict__ qblob, const uint32_t* __restrict__ qboff,
                       const ConfWork* __restrict__ worklist, const unsigned int* __restrict__ wide_count, uint32_t work_cap,
                       OutRec* __restrict__ out) {
  const uint32_t total = min(*wide_count, work_cap);
  esf::PatTable T;
  T.pats = ix->conf_pats;
  T.instrs = ix->conf_instrs;
  T.opts = ix->conf_opts;
  T.text = ix->conf_text;
  T.n_pats = ix->n_conf_pats;
  const uint8_t* __restrict__ vtext = ix->vocab_text;
  const uint32_t* __restrict__ voff = ix->vocab_text_off;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
    const ConfWork it = worklist[work_cap - 1 - w];
    const OutRec r = out[it.rec];
    const uint32_t vocab = r.vocab_id & ~OUT_SKIP_CONFUSABLES;
    const uint32_t a0 = qboff[it.query], a1 = qboff[it.query + 1];
    const uint32_t b0 = __ldg(voff + vocab), b1 = __ldg(voff + vocab + 1);
    if (a1 - a0 > 3u * esf::MAXLEN || b1 - b0 > 3u * esf::MAXLEN) continue;  // (the triage checked the character counts)
    uint16_t a[esf::MAXLEN], b[esf::MAXLEN];
    const int na = (int)u8_decode_bmp(qblob + a0, a1 - a0, a, esf::MAXLEN);
    const int nb = (int)u8_decode_bmp(vtext + b0, b1 - b0, b, esf::MAXLEN);
    esf::View v[esf::MAXSEG];
    const int nv = esf::shortest_edit_script_t<DeviceCharClass, uint16_t>(a, na, b, nb, v);
    if (nv < 0) continue;
    double weight = 1.0;
    for (uint32_t k = 0; k < T.n_pats; ++k) {
      const ConfPat pat = T.pats[k];
      if (esf::found_in(T, pat, a, b, v, nv)) weight = __dmul_rn(weight, pat.weight);
    }
    OutRec o = r;
    if (weight != 1.0) o.dist_score = __dmul_rn(r.dist_score, weight);
    o.vocab_id = r.vocab_id | OUT_SKIP_CONFUSABLES;
    out[it.rec] = o;
  }
}

// finish_kernel: one warp per query.  After the rescoring: re-rank (stable: the previous position is the
// last key), crop when the confusables ran before pruning, cut-off -- written back over the query's own
// records.  A query with an unsettled record is flagged for the host instead and left untouched.
__global__ void __launch_bounds__(K2_WARPS * 32)
finish_kernel(const BatchParams bp, uint32_t nq, OutRec* __restrict__ out, OutHead* __restrict__ out_head,
              SurvRec* __restrict__ scratch, uint32_t scratch_cap, unsigned int* work) {
  const uint32_t lane = lane_id();
  const uint32_t gwarp = blockIdx.x * K2_WARPS + (threadIdx.x >> 5);
  SurvRec* surv = scratch + (size_t)gwarp * 2 * scratch_cap;
  SurvRec* sorted = surv + scratch_cap;
  const uint32_t mode = (bp.finish_mode == FINISH_GATHER ? (RCE_RANK_SCORE | RCE_CROP | RCE_CUTOFF) : (RCE_RANK_SCORE | RCE_CUTOFF)) |
                        RCE_INPLACE;
  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(work, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= nq) break;
    const OutHead h = out_head[qi];
    const uint32_t n = h.count;
    if (n == 0) continue;
    OutRec* recs = out + h.offset;
    bool unsettled = n > scratch_cap;
    for (uint32_t i0 = 0; i0 < n && !unsettled; i0 += 32) {
      const uint32_t i = i0 + lane;
      bool u = false;
      if (i < n) {
        const OutRec o = recs[i];
        u = !(o.vocab_id & OUT_SKIP_CONFUSABLES);
        SurvRec s;
        s.dist = o.dist_score;
        s.freq = (double)o.freq;
        s.key = 0.0;
        s.g = i;
        s.raw = o.freq;
        s.vocab = o.vocab_id;
        s.pad = 0;
        surv[i] = s;
      }
      unsettled = __any_sync(FULL, u);
    }
    if (unsettled) {
      if (lane == 0) out_head[qi].count = n | HEAD_HOST_FINISH;
      continue;
    }
    rank_crop_emit(bp, mode, surv, sorted, n, h.max_freq, out, nullptr, out_head, qi, 0, nullptr, nullptr, h.offset, ConfStage());
  }
}

// Lexicon-sharded mode: a hit-list overflow on any shard cannot be repaired after the exchange.  res[0] = 1 if some
// (shard, query) carries QF_HIT_OVERFLOW without QF_UNSUPPORTED, res[1] = the smallest such query (pre-set to ~0).
__global__ void shard_flagcheck_kernel(const uint32_t* __restrict__ flags_all, uint32_t n, uint32_t n_shards, uint64_t stride,
                                       unsigned int* res) {
  const uint64_t total = (uint64_t)n * n_shards;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(i / n), q = (uint32_t)(i % n);
    const uint32_t f = flags_all[(size_t)r * stride + q];
    if ((f & (QF_HIT_OVERFLOW | QF_UNSUPPORTED)) == QF_HIT_OVERFLOW) {
      atomicOr(res, 1u);
      atomicMin(res + 1, q);
    }
  }
}
cudaError_t launch_shard_flagcheck(const uint32_t* flags_all, uint32_t n, uint32_t n_shards, uint64_t stride, unsigned int* res,
                                   int sm_count, cudaStream_t stream) {
  static const unsigned int init[2] = {0u, 0xFFFFFFFFu};
  cudaError_t e = cudaMemcpyAsync(res, init, sizeof init, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess || n == 0) return e;
  shard_flagcheck_kernel<<<(unsigned)sm_count * 8, 256, 0, stream>>>(flags_all, n, n_shards, stride, res);
  ++g_kernel_launches;
  return cudaGetLastError();
}

cudaError_t launch_merge(const BatchParams& bp, uint32_t n, uint32_t n_shards, const OutHead* heads_all, uint32_t head_stride,
                         const OutRec* recs_all, const uint32_t* gids_all, uint32_t rec_stride, const uint32_t* qflags_in, uint32_t* qflags,
                         OutRec* out, OutHead* out_head, void* scratch, uint32_t scratch_cap, unsigned int* work, int sm_count,
                         cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(work, 0, 4 * sizeof(unsigned int), stream);
  if (e != cudaSuccess) return e;
  long long grid = merge_grid(sm_count, n);
  merge_kernel<<<(unsigned)grid, K2_WARPS * 32, 0, stream>>>(bp, n, n_shards, heads_all, head_stride, recs_all, gids_all, rec_stride,
                                                             qflags_in, qflags, out, out_head,
                                                             reinterpret_cast<SurvRec*>(scratch), scratch_cap, work + 1,
                                                             work + 2);
  ++g_kernel_launches;
  return cudaGetLastError();
}
long long merge_grid(int sm_count, uint32_t n) {
  long long grid = (long long)sm_count * 8;
  const long long want = ((long long)n + K2_WARPS - 1) / K2_WARPS;
  if (grid > want) grid = want;
  return grid < 1 ? 1 : grid;
}
size_t merge_scratch_bytes(int sm_count, uint32_t n, uint32_t scratch_cap) {
  return (size_t)merge_grid(sm_count, n) * K2_WARPS * 2 * scratch_cap * sizeof(SurvRec);
}

// ================================================================================================
// Kernels 2p: scoring over a global, shape-sorted list of (query, candidate) pairs
// ================================================================================================
// score_kernel keeps a query's candidates together: one warp, rounds of 32 lanes, every lane walking a matrix as
// large as the longest candidate of its round.  On cfg 2 that ran 19.6 useful lanes per round and 2.4 matrix cells
// per useful cell (ncu, round 1) -- and one query with thousands of candidates was one warp's job.  Here the pairs of
// ALL queries are sorted by matrix shape (query length, candidate length) first, so that the 32 lanes of a warp hold
// 32 pairs of the same shape whatever queries they belong to:
//   pairfilter_kernel : per query (one warp): length check + bit-parallel OSA rejection (as prefilter_kernel), hit
//                       list compacted in place, a dense slot range reserved for the survivors, shape histogram
//   pairscan_kernel   : exclusive scan of the histogram -> first position of every shape; long shapes go last,
//                       starting at a tile boundary, so every tile of 32 positions belongs to one launch class
//   pairscatter_kernel: per query: every surviving candidate takes the next position of its shape
//   dp_kernel         : per tile of 32 pairs (one warp): true Damerau-Levenshtein + LCS + prefix + suffix in lock step,
//                       no padding lanes, no padding columns; the packed features go to the pair's dense slot
//   rank_kernel       : per query (one warp): features -> f64 score, max frequency, threshold, rank, crop, cut-off
constexpr uint32_t PAIR_SHORT_MAX = 24;     // longest side of a "short" shape (the DP's small shared-memory class)
constexpr uint32_t PAIR_SHORT_KEYS = 640;   // short shapes: Lq * 25 + Lc (625 used); long shapes follow:
constexpr uint32_t PAIR_BUCKETS = PAIR_TABLE;  // min(Lq, 63) << 6 | min(Lc, 63) behind the short ones (4736 used of 5120)
constexpr uint32_t PAIR_HOLE = 0xFFFFFFFFu;  // pair_q of an unused position (padding before the long class)
constexpr uint32_t RES_REJECT = 0xFFFFFFFFu;  // packed features of a candidate beyond the edit distance
__device__ __forceinline__ uint32_t pair_bucket(uint32_t Lq, uint32_t Lc) {
  if (max(Lq, Lc) <= PAIR_SHORT_MAX) return Lq * (PAIR_SHORT_MAX + 1) + Lc;
  return PAIR_SHORT_KEYS + ((min(Lq, 63u) << 6) | min(Lc, 63u));
}
static_assert(PAIR_SHORT_KEYS >= (PAIR_SHORT_MAX + 1) * (PAIR_SHORT_MAX + 1) && PAIR_SHORT_KEYS + 4096 <= PAIR_TABLE, "shape table");
static_assert(PAIR_TABLE % 1024 == 0 && PAIR_SHORT_KEYS % (PAIR_TABLE / 1024) == 0, "pairscan_kernel splits the table over 1024 threads");
// work[] slots of the pair path (the launchers zero them)
constexpr int PW_TOTAL = 8;      // dense slots reserved = pairs that passed the filter
constexpr int PW_TILES_A = 9;    // tiles of the short class
constexpr int PW_TILES_ALL = 10; // all tiles
constexpr int PW_NPOS = 17;      // positions of the sorted pair list (pairs + padding)
constexpr int PW_DP_A = 11;      // work counters: dp short class, dp long class, rank (two phases), scatter, filter
constexpr int PW_DP_B = 12;
constexpr int PW_RANK = 13;
constexpr int PW_SCATTER = 15;
constexpr int PW_FILTER = 16;

constexpr int PF_WARPS = 8;
__global__ void __launch_bounds__(PF_WARPS * 32)
pairfilter_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
                  const uint32_t* __restrict__ qlist, uint32_t nq, uint32_t* hits, uint32_t* hit_count, uint32_t* qflags,
                  uint32_t* __restrict__ qbase, uint32_t* __restrict__ hist, unsigned int* work, Counters* counters) {
  __shared__ uint32_t pm_s[PF_WARPS][256];  // per warp: bit j of pm[c] set iff query symbol j equals c
  __shared__ uint32_t hist_s[PAIR_BUCKETS];
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t pm_a = (uint32_t)__cvta_generic_to_shared(&pm_s[warp][0]);
  for (uint32_t k = lane; k < 256; k += 32) sts_u32(pm_a + k * 4, 0);
  for (uint32_t k = threadIdx.x; k < PAIR_BUCKETS; k += blockDim.x) hist_s[k] = 0;
  __syncthreads();
  const uint8_t* __restrict__ rows = ix->inst_rows;
  const uint32_t nstride = ix->norm_stride;
  unsigned long long c_pairs = 0, c_cells = 0;
  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(work + PW_FILTER, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= nq) break;
    const uint32_t flags = qflags[qi];
    if (flags & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED)) {
      if (lane == 0) qbase[qi] = 0;
      continue;
    }
    const uint32_t nh = hit_count[qi];
    const uint32_t q = qlist ? qlist[qi] : qi;
    const uint8_t* qrow = queries + (size_t)q * bp.query_stride;
    const uint32_t Lq = qrow[0];
    uint32_t* hq = hits + (size_t)qi * bp.hit_cap;
    uint32_t w = 0;
    if (flags & QF_PREFILTERED) {
      // a re-run of the score stage (pool or pair-list overflow): the list is filtered already, only the shapes
      // are counted again
      for (uint32_t hb = 0; hb < nh; hb += 32) {
        const uint32_t hi = hb + lane;
        if (hi < nh) atomicAdd(&hist_s[pair_bucket(Lq, __ldg(rows + (size_t)hq[hi] * nstride))], 1u);
      }
      w = nh;
    } else {
      const uint32_t ke = apply_threshold(bp.max_edit, Lq);
      const bool osa = Lq <= 32;  // the bit-parallel distance holds the query in one 32-bit word
      uint32_t mysym = 256u + lane;  // (lanes beyond the query get unique values: they match nobody)
      if (osa) {
        if (lane < Lq) mysym = qrow[2 + lane];
        const uint32_t mm = __match_any_sync(FULL, mysym);  // the lanes (= query positions) holding the same symbol
        if (lane < Lq) sts_u32(pm_a + mysym * 4, mm);
        __syncwarp();
      }
      const uint32_t top = 1u << ((Lq - 1) & 31);
      const uint32_t osa_max = ke + ke / 2;
      for (uint32_t hb = 0; hb < nh; hb += 32) {
        const uint32_t hi = hb + lane;
        bool valid = hi < nh;
        uint32_t g = 0, Lc = 0;
        const uint8_t* row = rows;
        uint4 v0 = make_uint4(0, 0, 0, 0);
        if (valid) {
          g = hq[hi];
          row = rows + (size_t)g * nstride;
          v0 = __ldg(reinterpret_cast<const uint4*>(row));
          Lc = v0.x & 0xFF;
          const uint32_t diff = Lq > Lc ? Lq - Lc : Lc - Lq;
          valid = diff <= ke;  // length pre-check of damerau_levenshtein (src/distance.rs:109-130)
          if (valid) {
            c_pairs += 1;
            c_cells += (unsigned long long)Lq * Lc;
          }
        }
        const uint32_t Lreal = Lc;
        if (!valid) Lc = 0;
        bool pass = valid;
        if (osa) {
          const uint32_t Lcm = __reduce_max_sync(FULL, Lc);
          uint32_t D0 = 0, VP = 0xFFFFFFFFu, VN = 0, PMp = 0, sc = Lq;
          const uint32_t nbytes = Lcm ? Lcm + 2 : 0;  // row bytes to walk: len, flags, symbols
          for (uint32_t k0 = 0; k0 < nbytes; k0 += 16) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (k0 < Lc + 2 && Lc) v = (k0 == 0) ? v0 : __ldg(reinterpret_cast<const uint4*>(row + k0));
            uint32_t x = v.x, y = v.y, z = v.z, t = v.w;
            uint32_t pos = k0, steps = min(16u, nbytes - k0);
            if (k0 == 0) {  // skip the length and flag bytes
              x = __funnelshift_r(x, y, 16);
              y = __funnelshift_r(y, z, 16);
              z = __funnelshift_r(z, t, 16);
              t >>= 16;
              pos = 2;
              steps -= 2;
            }
            for (uint32_t st = 0; st < steps; ++st, ++pos) {
              const uint32_t c = x & 0xFFu;
              x = __funnelshift_r(x, y, 8);
              y = __funnelshift_r(y, z, 8);
              z = __funnelshift_r(z, t, 8);
              t >>= 8;
              if (pos < Lc + 2) {  // this lane's candidate still has symbols (never for dropped lanes: Lc = 0)
                const uint32_t PMj = lds_u32(pm_a + c * 4);
                const uint32_t TR = ((~D0 & PMj) << 1) & PMp;  // adjacent transposition
                D0 = TR | (((PMj & VP) + VP) ^ VP) | PMj | VN;
                const uint32_t HP = VN | ~(D0 | VP);
                const uint32_t HN = D0 & VP;
                sc += (HP & top) ? 1u : 0u;
                sc -= (HN & top) ? 1u : 0u;
                const uint32_t X = (HP << 1) | 1u;
                VP = (HN << 1) | ~(D0 | X);
                VN = X & D0;
                PMp = PMj;
              }
            }
          }
          pass = valid && sc <= osa_max;
        }
        const uint32_t pmask = __ballot_sync(FULL, pass);
        __syncwarp();  // every lane has read its entry of this batch: the compacted list may overwrite it
        if (pass) {
          hq[w + __popc(pmask & lanemask_lt())] = g;
          atomicAdd(&hist_s[pair_bucket(Lq, Lreal)], 1u);
        }
        w += __popc(pmask);
      }
      __syncwarp();
      if (osa && lane < Lq) sts_u32(pm_a + mysym * 4, 0);  // leave the table clean for the next query
    }
    if (lane == 0) {
      hit_count[qi] = w;
      qflags[qi] = flags | QF_PREFILTERED;
      qbase[qi] = w ? atomicAdd(work + PW_TOTAL, w) : 0;  // the query's dense slot range
    }
    __syncwarp();
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < PAIR_BUCKETS; k += blockDim.x) {
    const uint32_t v = hist_s[k];
    if (v) atomicAdd(hist + k, v);
  }
  if (counters) {
    for (int o = 16; o > 0; o >>= 1) {
      c_pairs += __shfl_xor_sync(FULL, c_pairs, o);
      c_cells += __shfl_xor_sync(FULL, c_cells, o);
    }
    if (lane == 0) {
      atomicAdd(&counters->dl_pairs, c_pairs);
      atomicAdd(&counters->dl_cells, c_cells);
    }
  }
}

// hist[PAIR_BUCKETS] -> first[PAIR_BUCKETS] (exclusive scan; the long class starts at a multiple of 32), cursors zeroed,
// tile counts published.  One CTA.
__global__ void __launch_bounds__(1024)
pairscan_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ first, uint32_t* __restrict__ cursor,
                uint4* __restrict__ pairs, uint32_t pair_cap, unsigned int* work) {
  __shared__ uint32_t s_part[1024];
  __shared__ uint32_t s_short;
  constexpr uint32_t PER = PAIR_BUCKETS / 1024;
  const uint32_t t = threadIdx.x;
  uint32_t v[PER], sum = 0;
#pragma unroll
  for (uint32_t k = 0; k < PER; ++k) {
    v[k] = hist[t * PER + k];
    sum += v[k];
    cursor[t * PER + k] = 0;
  }
  s_part[t] = sum;
  __syncthreads();
  // inclusive scan of the 1024 partial sums (Hillis-Steele in shared memory)
  for (uint32_t o = 1; o < 1024; o <<= 1) {
    const uint32_t add = t >= o ? s_part[t - o] : 0;
    __syncthreads();
    s_part[t] += add;
    __syncthreads();
  }
  if (t == PAIR_SHORT_KEYS / PER - 1) s_short = s_part[t];  // pairs of the short class
  __syncthreads();
  const uint32_t n_short = s_short, pad = ((n_short + 31) & ~31u) - n_short;
  uint32_t run = s_part[t] - sum + (t * PER >= PAIR_SHORT_KEYS ? pad : 0);
#pragma unroll
  for (uint32_t k = 0; k < PER; ++k) {
    first[t * PER + k] = run;
    run += v[k];
  }
  if (t < pad && n_short + t < pair_cap) pairs[n_short + t] = make_uint4(PAIR_HOLE, 0, 0, 0);  // the padding positions hold no pair
  if (t == 1023) {
    const uint32_t n_all = s_part[1023] + pad;  // positions incl. the padding between the classes
    work[PW_TILES_A] = (n_short + 31) / 32;
    work[PW_TILES_ALL] = (n_all + 31) / 32;
    work[PW_NPOS] = n_all;
  }
}

// every surviving candidate takes the next position of its shape
constexpr int PS_WARPS = 8;
__global__ void __launch_bounds__(PS_WARPS * 32)
pairscatter_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
                   const uint32_t* __restrict__ qlist, uint32_t nq, const uint32_t* __restrict__ hits,
                   const uint32_t* __restrict__ hit_count, const uint32_t* __restrict__ qflags, const uint32_t* __restrict__ qbase,
                   const uint32_t* __restrict__ first, uint32_t* __restrict__ cursor, uint4* __restrict__ pairs,
                   uint32_t pair_cap, unsigned int* work) {
  const uint32_t lane = lane_id();
  const uint8_t* __restrict__ rows = ix->inst_rows;
  const uint32_t nstride = ix->norm_stride;
  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(work + PW_SCATTER, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= nq) break;
    if (qflags[qi] & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED)) continue;
    const uint32_t nh = hit_count[qi];
    if (nh == 0) continue;
    const uint32_t q = qlist ? qlist[qi] : qi;
    const uint32_t Lq = queries[(size_t)q * bp.query_stride];
    const uint32_t* hq = hits + (size_t)qi * bp.hit_cap;
    const uint32_t base = qbase[qi];
    for (uint32_t hb = 0; hb < nh; hb += 32) {
      const uint32_t k = hb + lane;
      const bool valid = k < nh;
      uint32_t g = 0, bucket = PAIR_BUCKETS + lane;  // (idle lanes: unique keys, they match nobody)
      if (valid) {
        g = hq[k];
        bucket = pair_bucket(Lq, __ldg(rows + (size_t)g * nstride));
      }
      const uint32_t peers = __match_any_sync(FULL, bucket);
      if (valid) {
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t pos = 0;
        if (lane == leader) pos = atomicAdd(cursor + bucket, (uint32_t)__popc(peers));
        pos = __shfl_sync(peers, pos, leader) + __popc(peers & lanemask_lt()) + first[bucket];
        const uint32_t d = base + k;
        if (pos < pair_cap && d < pair_cap) pairs[pos] = make_uint4(qi, g, d, 0);  // one 16-byte record per pair
      }
    }
  }
}

// ---- the DP over tiles of 32 same-shape pairs -------------------------------------------------------------------------
// shared memory of one warp (dynamic; sized by the launch class: MQ query rows, MC candidate columns, ring depth R)
//   qs[MQ][32] (uint8)             query symbols, per lane
//   cell[(MC+1)][32] (uint32)      per column j, per lane: {t[j-1], lcs[j], lastrow[j], D[i-1][j]}
//   ring[R][(MC+1)][32] (uint8)    the last R rows of the DL matrix, per lane
__host__ __device__ inline size_t dp_warp_bytes(uint32_t MQ, uint32_t MC, uint32_t R) {
  return (size_t)MQ * 32 + (size_t)(MC + 1) * 32 * 4 + (size_t)R * (MC + 1) * 32;
}
constexpr int DP_WARPS = 4;
#ifndef ANL_DP_MIN_CTAS
#define ANL_DP_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(DP_WARPS * 32, ANL_DP_MIN_CTAS)
dp_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
          const uint32_t* __restrict__ qlist, const uint4* __restrict__ pairs, uint32_t pair_cap, uint32_t* __restrict__ res,
          unsigned int* work, int cls, Counters* counters, uint32_t MQ, uint32_t MC, uint32_t R) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + warp * (uint32_t)dp_warp_bytes(MQ, MC, R);
  const uint32_t qs_a = sbase + lane;                                        // + i * 32
  const uint32_t cell_a = sbase + MQ * 32 + lane * 4;                        // + j * 128
  const uint32_t ring_a = sbase + MQ * 32 + (MC + 1) * 32 * 4 + lane;        // + slot * rowbytes + j * 32
  const uint32_t rowbytes = (MC + 1) * 32;
  const uint8_t* __restrict__ rows = ix->inst_rows;
  const uint32_t nstride = ix->norm_stride;
  // tiles [t_lo, t_hi) of this launch class; positions beyond n_pos hold nothing
  const uint32_t tiles_a = work[PW_TILES_A], tiles_all = work[PW_TILES_ALL];
  const uint32_t t_lo = cls == 0 ? 0u : tiles_a, t_hi = cls == 0 ? tiles_a : tiles_all;
  const uint32_t n_pos = min(work[PW_NPOS], pair_cap);
  unsigned int* counter = work + (cls == 0 ? PW_DP_A : PW_DP_B);
  const uint32_t S = R;  // shift of row / column numbers: "no previous occurrence" (0) fails the reach test
  unsigned long long c_dpp = 0, c_dpc = 0;
  for (;;) {
    uint32_t tile = 0;
    // (the pair list is sorted by shape, largest matrices last: walk it from the back so that the costly tiles start first)
    if (lane == 0) tile = atomicAdd(counter, 1u);
    tile = __shfl_sync(FULL, tile, 0);
    if (tile >= t_hi - t_lo) break;
    tile = t_hi - 1 - tile;
    const uint32_t p = tile * 32 + lane;
    uint32_t qi = PAIR_HOLE, g = 0, d = 0;
    if (p < n_pos) {
      const uint4 pr = pairs[p];
      qi = pr.x;
      g = pr.y;
      d = pr.z;
    }
    const bool valid = qi != PAIR_HOLE;
    uint32_t Lq = 0, Lc = 0, ke = 0;
    if (valid) {
      // stage the query's and the candidate's symbols, one per shared-memory row / column of this lane
      const uint32_t q = qlist ? qlist[qi] : qi;
      const uint8_t* qrow = queries + (size_t)q * bp.query_stride;
      const uint8_t* row = rows + (size_t)g * nstride;
      const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(qrow));
      const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(row));
      Lq = min(q0.x & 0xFF, MQ);  // (the class bounds hold by construction of the tiles; the clamp guards the arrays)
      Lc = min(c0.x & 0xFF, MC);
      ke = apply_threshold(bp.max_edit, q0.x & 0xFF);
      for (uint32_t j0 = 0; j0 < Lq + 2; j0 += 16) {
        const uint4 v = (j0 == 0) ? q0 : __ldg(reinterpret_cast<const uint4*>(qrow + j0));
        uint32_t x = v.x, y = v.y, z = v.z, t = v.w;
        const uint32_t jend = min(j0 + 16, Lq + 2);
#pragma unroll 1
        for (uint32_t bytepos = j0; bytepos < jend; ++bytepos) {
          const uint32_t sym = x & 0xFFu;
          x = __funnelshift_r(x, y, 8);
          y = __funnelshift_r(y, z, 8);
          z = __funnelshift_r(z, t, 8);
          t >>= 8;
          if (bytepos >= 2) sts_u8(qs_a + (bytepos - 2) * 32, sym);
        }
      }
      for (uint32_t j0 = 0; j0 < Lc + 2; j0 += 16) {
        const uint4 v = (j0 == 0) ? c0 : __ldg(reinterpret_cast<const uint4*>(row + j0));
        uint32_t x = v.x, y = v.y, z = v.z, t = v.w;
        const uint32_t jend = min(j0 + 16, Lc + 2);
#pragma unroll 1
        for (uint32_t bytepos = j0; bytepos < jend; ++bytepos) {
          const uint32_t sym = x & 0xFFu;
          x = __funnelshift_r(x, y, 8);
          y = __funnelshift_r(y, z, 8);
          z = __funnelshift_r(z, t, 8);
          t >>= 8;
          // column j = bytepos - 1: {t, lcs = 0, lastrow = 0, D[0][j] = j}
          if (bytepos >= 2) sts_u32(cell_a + (bytepos - 1) * 128, sym | ((bytepos - 1) << 24));
        }
      }
    }
    const uint32_t Lqm = __reduce_max_sync(FULL, Lq), Lcm = __reduce_max_sync(FULL, Lc);
    // (tiles are same-shape except where two shapes meet: shorter strings are padded with symbols that match nothing)
    for (uint32_t i = Lq; i < Lqm; ++i) sts_u8(qs_a + i * 32, 0xFEu);
    for (uint32_t j = Lc + 1; j <= Lcm; ++j) sts_u32(cell_a + j * 128, 0xFFu | (j << 24));
    for (uint32_t j = 0; j <= Lcm; ++j) sts_u8(ring_a + j * 32, j);  // row 0 in slot 0
    c_dpp += valid ? 1 : 0;
    c_dpc += (unsigned long long)Lqm * Lcm;  // per lane: x 32 lanes in the sum = warp-cells of this tile

    // ---- true Damerau-Levenshtein, all lanes in lock-step over (i, j); see score_kernel for the recurrence -------
    uint32_t lcs_best = 0, ld = 255;
    uint32_t slot = 0;  // ring slot of row i - 1
    for (uint32_t i = 1; i <= Lqm; ++i) {
      const uint32_t sc = lds_u8(qs_a + (i - 1) * 32);
      slot = slot + 1 == R ? 0 : slot + 1;
      const uint32_t cur_a = ring_a + slot * rowbytes;
      const uint32_t is = i + S;  // shifted row number
      uint32_t left = i, diag = i - 1, db = 0, lcs_diag = 0;
      sts_u8(cur_a, i);
      for (uint32_t j = 1; j <= Lcm; ++j) {
        const uint32_t cw = lds_u32(cell_a + j * 128);
        const uint32_t tc = cw & 0xFF, lcs_up = (cw >> 8) & 0xFF, last = (cw >> 16) & 0xFF, up = cw >> 24;
        const bool same = tc == sc;
        const uint32_t js = j + S;
        uint32_t v = min(min(left, up) + 1, diag + (same ? 0u : 1u));
        // transposition (src/distance.rs:160-165): mat[last][db] + (i-last-1) + 1 + (j-db-1); a term that
        // reaches back more than ke + 1 in total cannot be <= ke and is skipped (exact)
        const uint32_t reach = (is - last) + (js - db);
        if (reach <= ke + 1) {
          const uint32_t back = is - last + 1;  // rows between row i and row last-1
          const uint32_t ts = slot >= back ? slot - back : slot + R - back;
          const uint32_t tv = lds_u8(ring_a + ts * rowbytes + (db - S - 1) * 32) + reach - 1;
          v = min(v, tv);
        }
        sts_u8(cur_a + j * 32, v);  // (a distance never exceeds the longer side: it fits the byte)
        const uint32_t lcs_new = same ? lcs_diag + 1 : 0;
        lcs_best = max(lcs_best, lcs_new);
        sts_u32(cell_a + j * 128, tc | (lcs_new << 8) | ((same ? is : last) << 16) | (v << 24));
        if (same) db = js;
        lcs_diag = lcs_up;
        diag = up;
        left = v;
      }
      if (i == Lq) ld = lds_u8(cur_a + Lc * 32);  // this lane's own matrix ends here
    }
    // ---- prefix / suffix (src/distance.rs:208-231) ------------------------------------------------
    uint32_t pre = 0, suf = 0;
    const bool within = valid && ld <= ke;
    {
      const uint32_t lim = min(Lqm, Lcm);
      bool pgo = within, sgo = within;
      for (uint32_t i = 0; i < lim; ++i) {
        if (within && i < Lc && i < Lq) {
          const uint32_t a = lds_u32(cell_a + (i + 1) * 128) & 0xFF;
          pgo = pgo && (a == lds_u8(qs_a + i * 32));
          pre += pgo;
          const uint32_t b = lds_u32(cell_a + (Lc - i) * 128) & 0xFF;
          sgo = sgo && (b == lds_u8(qs_a + (Lq - 1 - i) * 32));
          suf += sgo;
        }
      }
    }
    if (valid && d < pair_cap) res[d] = within ? (ld | (lcs_best << 8) | (pre << 16) | (suf << 24)) : RES_REJECT;
    __syncwarp();
  }
  if (counters) {
    for (int o = 16; o > 0; o >>= 1) {
      c_dpp += __shfl_xor_sync(FULL, c_dpp, o);
      c_dpc += __shfl_xor_sync(FULL, c_dpc, o);
    }
    if (lane == 0) {
      atomicAdd(&counters->dp_pairs, c_dpp);
      atomicAdd(&counters->dp_cells, c_dpc);
    }
  }
}

// ---- the same DP with the rows of the NEXT tile brought in by the TMA engine while this tile is computed ----------------
// dp_kernel loads a tile's 32 query rows and 32 candidate rows with ordinary vector loads and only then starts the
// matrix: every tile begins with a global-memory round trip.  Here a lane issues two 1-D bulk copies
// (cp.async.bulk global -> shared, completion on an mbarrier of the warp) for the rows of tile t + 1 as soon as the
// staging buffer of tile t has been unpacked, and the pair records of tile t + 2 are already on their way into
// registers: the loads leave the issue stream, and their latency hides behind the ~3000 instructions of a tile's DP.
// Rows are fixed-stride (norm_stride / query_stride, multiples of 16 bytes, 16-byte aligned): exactly what a 1-D bulk
// copy needs.  Staging layout: [lane][row bytes], rows 48 bytes apart (not 32: a 128-bit shared load serves 8 lanes
// per pass, and 8 rows 48 bytes apart fall into 8 different bank groups).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__host__ __device__ inline uint32_t dp_stage_stride(uint32_t row_bytes) { return (row_bytes % 32 == 0) ? row_bytes + 16 : row_bytes; }
// per warp: the arrays of dp_kernel + staging for 32 query rows and 32 candidate rows + one mbarrier
__host__ __device__ inline size_t dp_tma_warp_bytes(uint32_t MQ, uint32_t MC, uint32_t R, uint32_t qb, uint32_t cb) {
  const size_t dp = (dp_warp_bytes(MQ, MC, R) + 15) & ~(size_t)15;
  return dp + 32 * (size_t)(dp_stage_stride(qb) + dp_stage_stride(cb)) + 16;
}
__global__ void __launch_bounds__(DP_WARPS * 32, ANL_DP_MIN_CTAS)
dp_tma_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
              const uint32_t* __restrict__ qlist, const uint4* __restrict__ pairs, uint32_t pair_cap, uint32_t* __restrict__ res,
              unsigned int* work, int cls, Counters* counters, uint32_t MQ, uint32_t MC, uint32_t R, uint32_t qb, uint32_t cb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t wbytes = (uint32_t)dp_tma_warp_bytes(MQ, MC, R, qb, cb);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + warp * wbytes;
  const uint32_t qs_a = sbase + lane;                                        // + i * 32
  const uint32_t cell_a = sbase + MQ * 32 + lane * 4;                        // + j * 128
  const uint32_t ring_a = sbase + MQ * 32 + (MC + 1) * 32 * 4 + lane;        // + slot * rowbytes + j * 32
  const uint32_t rowbytes = (MC + 1) * 32;
  const uint32_t qstride = dp_stage_stride(qb), cstride = dp_stage_stride(cb);
  const uint32_t stage_q = sbase + (((uint32_t)dp_warp_bytes(MQ, MC, R) + 15) & ~15u) + lane * qstride;
  const uint32_t stage_c = sbase + (((uint32_t)dp_warp_bytes(MQ, MC, R) + 15) & ~15u) + 32 * qstride + lane * cstride;
  const uint32_t bar = sbase + wbytes - 16;
  const uint8_t* __restrict__ rows = ix->inst_rows;
  const uint32_t nstride = ix->norm_stride;
  const uint32_t tiles_a = work[PW_TILES_A], tiles_all = work[PW_TILES_ALL];
  const uint32_t t_lo = cls == 0 ? 0u : tiles_a, t_hi = cls == 0 ? tiles_a : tiles_all;
  const uint32_t n_pos = min(work[PW_NPOS], pair_cap);
  unsigned int* counter = work + (cls == 0 ? PW_DP_A : PW_DP_B);
  const uint32_t S = R;
  unsigned long long c_dpp = 0, c_dpc = 0;
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t parity = 0;
  auto grab = [&]() {
    uint32_t t = 0;
    if (lane == 0) t = t_lo + atomicAdd(counter, 1u);
    return __shfl_sync(FULL, t, 0);
  };
  struct Pair {
    uint32_t qi, g, d;
  };
  auto load_pair = [&](uint32_t tile) {
    Pair pr{PAIR_HOLE, 0, 0};
    const uint32_t p = tile * 32 + lane;
    if (tile < t_hi && p < n_pos) {
      const uint4 v = pairs[p];
      pr.qi = v.x;
      pr.g = v.y;
      pr.d = v.z;
    }
    return pr;
  };
  // rows of a tile -> the staging buffer (the previous tile's rows have been unpacked: the buffer is free)
  auto issue_rows = [&](const Pair& pr) {
    const bool v = pr.qi != PAIR_HOLE;
    const uint32_t nv = __popc(__ballot_sync(FULL, v));
    fence_proxy_async_smem();  // the generic-proxy reads of the buffer precede the async-proxy writes
    if (lane == 0) mbar_arrive_expect_tx(bar, nv * (qb + cb));
    __syncwarp();
    if (v) {
      const uint32_t q = qlist ? qlist[pr.qi] : pr.qi;
      bulk_copy_g2s(stage_q, queries + (size_t)q * bp.query_stride, qb, bar);
      bulk_copy_g2s(stage_c, rows + (size_t)pr.g * nstride, cb, bar);
    }
  };
  uint32_t tile = grab();
  Pair cur = load_pair(tile);
  uint32_t tile1 = grab();
  Pair nxt = load_pair(tile1);
  if (tile < t_hi) issue_rows(cur);
  while (tile < t_hi) {
    // ---- wait for this tile's rows, unpack them into the DP's column layout --------------------------------------
    {
      uint32_t spins = 0;
      while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();  // (a lost copy would otherwise hang the device)
      }
      parity ^= 1;
    }
    const bool valid = cur.qi != PAIR_HOLE;
    uint32_t Lq = 0, Lc = 0, ke = 0;
    if (valid) {
      const uint4 q0 = lds_u128(stage_q), c0 = lds_u128(stage_c);
      Lq = min(q0.x & 0xFF, min(MQ, qb - 2));
      Lc = min(c0.x & 0xFF, min(MC, cb - 2));
      ke = apply_threshold(bp.max_edit, q0.x & 0xFF);
      for (uint32_t j0 = 0; j0 < Lq + 2; j0 += 16) {
        const uint4 v = (j0 == 0) ? q0 : lds_u128(stage_q + j0);
        uint32_t x = v.x, y = v.y, z = v.z, t = v.w;
        const uint32_t jend = min(j0 + 16, Lq + 2);
#pragma unroll 1
        for (uint32_t bytepos = j0; bytepos < jend; ++bytepos) {
          const uint32_t sym = x & 0xFFu;
          x = __funnelshift_r(x, y, 8);
          y = __funnelshift_r(y, z, 8);
          z = __funnelshift_r(z, t, 8);
          t >>= 8;
          if (bytepos >= 2) sts_u8(qs_a + (bytepos - 2) * 32, sym);
        }
      }
      for (uint32_t j0 = 0; j0 < Lc + 2; j0 += 16) {
        const uint4 v = (j0 == 0) ? c0 : lds_u128(stage_c + j0);
        uint32_t x = v.x, y = v.y, z = v.z, t = v.w;
        const uint32_t jend = min(j0 + 16, Lc + 2);
#pragma unroll 1
        for (uint32_t bytepos = j0; bytepos < jend; ++bytepos) {
          const uint32_t sym = x & 0xFFu;
          x = __funnelshift_r(x, y, 8);
          y = __funnelshift_r(y, z, 8);
          z = __funnelshift_r(z, t, 8);
          t >>= 8;
          if (bytepos >= 2) sts_u32(cell_a + (bytepos - 1) * 128, sym | ((bytepos - 1) << 24));
        }
      }
    }
    __syncwarp();
    // ---- the staging buffer is free: the next tile's rows start their trip, the tile after that its pair records ----
    const uint32_t tile2 = grab();
    if (tile1 < t_hi) issue_rows(nxt);
    const Pair nxt2 = load_pair(tile2);

    const uint32_t Lqm = __reduce_max_sync(FULL, Lq), Lcm = __reduce_max_sync(FULL, Lc);
    for (uint32_t i = Lq; i < Lqm; ++i) sts_u8(qs_a + i * 32, 0xFEu);
    for (uint32_t j = Lc + 1; j <= Lcm; ++j) sts_u32(cell_a + j * 128, 0xFFu | (j << 24));
    for (uint32_t j = 0; j <= Lcm; ++j) sts_u8(ring_a + j * 32, j);  // row 0 in slot 0
    c_dpp += valid ? 1 : 0;
    c_dpc += (unsigned long long)Lqm * Lcm;
    uint32_t lcs_best = 0, ld = 255;
    uint32_t slot = 0;
    for (uint32_t i = 1; i <= Lqm; ++i) {
      const uint32_t sc = lds_u8(qs_a + (i - 1) * 32);
      slot = slot + 1 == R ? 0 : slot + 1;
      const uint32_t cur_a = ring_a + slot * rowbytes;
      const uint32_t is = i + S;
      uint32_t left = i, diag = i - 1, db = 0, lcs_diag = 0;
      sts_u8(cur_a, i);
      for (uint32_t j = 1; j <= Lcm; ++j) {
        const uint32_t cw = lds_u32(cell_a + j * 128);
        const uint32_t tc = cw & 0xFF, lcs_up = (cw >> 8) & 0xFF, last = (cw >> 16) & 0xFF, up = cw >> 24;
        const bool same = tc == sc;
        const uint32_t js = j + S;
        uint32_t v = min(min(left, up) + 1, diag + (same ? 0u : 1u));
        const uint32_t reach = (is - last) + (js - db);
        if (reach <= ke + 1) {
          const uint32_t back = is - last + 1;
          const uint32_t ts = slot >= back ? slot - back : slot + R - back;
          const uint32_t tv = lds_u8(ring_a + ts * rowbytes + (db - S - 1) * 32) + reach - 1;
          v = min(v, tv);
        }
        sts_u8(cur_a + j * 32, v);
        const uint32_t lcs_new = same ? lcs_diag + 1 : 0;
        lcs_best = max(lcs_best, lcs_new);
        sts_u32(cell_a + j * 128, tc | (lcs_new << 8) | ((same ? is : last) << 16) | (v << 24));
        if (same) db = js;
        lcs_diag = lcs_up;
        diag = up;
        left = v;
      }
      if (i == Lq) ld = lds_u8(cur_a + Lc * 32);
    }
    uint32_t pre = 0, suf = 0;
    const bool within = valid && ld <= ke;
    {
      const uint32_t lim = min(Lqm, Lcm);
      bool pgo = within, sgo = within;
      for (uint32_t i = 0; i < lim; ++i) {
        if (within && i < Lc && i < Lq) {
          const uint32_t a = lds_u32(cell_a + (i + 1) * 128) & 0xFF;
          pgo = pgo && (a == lds_u8(qs_a + i * 32));
          pre += pgo;
          const uint32_t b = lds_u32(cell_a + (Lc - i) * 128) & 0xFF;
          sgo = sgo && (b == lds_u8(qs_a + (Lq - 1 - i) * 32));
          suf += sgo;
        }
      }
    }
    if (valid && cur.d < pair_cap) res[cur.d] = within ? (ld | (lcs_best << 8) | (pre << 16) | (suf << 24)) : RES_REJECT;
    __syncwarp();
    tile = tile1;
    cur = nxt;
    tile1 = tile2;
    nxt = nxt2;
  }
  if (counters) {
    for (int o = 16; o > 0; o >>= 1) {
      c_dpp += __shfl_xor_sync(FULL, c_dpp, o);
      c_dpc += __shfl_xor_sync(FULL, c_dpc, o);
    }
    if (lane == 0) {
      atomicAdd(&counters->dp_pairs, c_dpp);
      atomicAdd(&counters->dp_cells, c_dpc);
    }
  }
}

// queries per counter increment of rank_kernel: at least ~24 grabs per warp, so that the warps end together
static uint32_t rank_grab(uint32_t n, long long grid) {
  const long long warps = grid * K2_WARPS;
  long long g = (long long)n / (warps * 24);
  if (g < 1) g = 1;
  if (g > 8) g = 8;
  return (uint32_t)g;
}

// ---- per query: features -> score -> rank / crop / cut-off -------------------------------------------------------------
// One warp per query, but the per-query latency chain is kept short: a warp takes 32 queries per counter increment and
// its lanes fetch their headers (flags, candidate count, dense base, length, case flag) side by side; queries with at
// most 32 candidates (nearly all) keep their survivor lists in shared memory instead of the global scratch.
constexpr uint32_t RK_SMEM_SURV = 32;
constexpr uint32_t RK_GRAB = 8;  // most queries per counter increment (a warp works its grab off serially; the launcher
                                 // passes fewer for small batches: at 131 072 queries a grab of 8 left the SMs idle 45 % of the kernel)
__global__ void __launch_bounds__(K2_WARPS * 32)
rank_kernel(const DeviceIndex* __restrict__ ix, const BatchParams bp, const uint8_t* __restrict__ queries,
            const uint32_t* __restrict__ qlist, uint32_t* __restrict__ rec_query, uint32_t nq, const uint32_t* __restrict__ hits,
            const uint32_t* __restrict__ hit_count, uint32_t* __restrict__ qflags, const uint32_t* __restrict__ qbase,
            const uint32_t* __restrict__ res, uint32_t pair_cap, OutRec* __restrict__ out, uint32_t* __restrict__ out_gid,
            OutHead* __restrict__ out_head, SurvRec* __restrict__ scratch, unsigned int* work, Counters* counters, uint32_t grab) {
  __shared__ SurvRec s_scr[K2_WARPS][2 * RK_SMEM_SURV];
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t gwarp = blockIdx.x * K2_WARPS + warp;
  SurvRec* g_surv = scratch + (size_t)gwarp * 2 * bp.hit_cap;  // survivors, then the sorted copy (long lists)
  const uint8_t* __restrict__ rows = ix->inst_rows;
  const uint32_t nstride = ix->norm_stride;
  const int have_freq = ix->have_freq;
  const uint32_t* __restrict__ gid_of = ix->inst_gid;
  unsigned long long c_surv = 0, c_res = 0;
  for (;;) {
    uint32_t q0 = 0;
    if (lane == 0) q0 = atomicAdd(work + PW_RANK, grab);
    q0 = __shfl_sync(FULL, q0, 0);
    if (q0 >= nq) break;
    // headers of the grabbed queries, one per lane
    uint32_t m_flags = QF_EMPTY, m_nh = 0, m_base = 0, m_q = 0, m_len = 0;
    if (lane < grab && q0 + lane < nq) {
      const uint32_t qi = q0 + lane;
      m_flags = qflags[qi];
      m_nh = hit_count[qi];
      m_base = qbase[qi];
      m_q = qlist ? qlist[qi] : qi;
      m_len = *reinterpret_cast<const uint16_t*>(queries + (size_t)m_q * bp.query_stride);  // length | flags << 8
    }
    const uint32_t nblock = min(grab, nq - q0);
    for (uint32_t t = 0; t < nblock; ++t) {
      const uint32_t qi = q0 + t;
      const uint32_t flags = __shfl_sync(FULL, m_flags, t);
      const uint32_t nh = __shfl_sync(FULL, m_nh, t);
      const uint32_t base = __shfl_sync(FULL, m_base, t);
      const uint32_t q = __shfl_sync(FULL, m_q, t);
      const uint32_t len = __shfl_sync(FULL, m_len, t);
      if ((flags & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED)) || (unsigned long long)base + nh > pair_cap) {
        // (a query beyond the pair-list capacity: the host runs the score stage again with a larger list)
        if (lane == 0) {
          OutHead h;
          h.max_freq = 0.0;
          h.offset = 0;
          h.count = 0;
          out_head[qi] = h;
        }
        continue;
      }
      const uint32_t Lq = len & 0xFF;
      const bool q_lower = ((len >> 8) & Q_FIRST_LOWER) != 0;
      const uint32_t* hq = hits + (size_t)qi * bp.hit_cap;
      SurvRec* surv = nh <= RK_SMEM_SURV ? &s_scr[warp][0] : g_surv;
      SurvRec* sorted = nh <= RK_SMEM_SURV ? &s_scr[warp][RK_SMEM_SURV] : g_surv + bp.hit_cap;
      const double Ld = (double)Lq;
      // every feature of the score is a small integer divided by the query length: lane v holds v / Ld once per
      // query and the per-candidate quotients are fetched by shuffle (same IEEE division, so the bits are the same)
      const double quot_lane = __ddiv_rn((double)lane, Ld);
      const bool quot_ok = Lq <= 31;
      uint32_t nsurv = 0;
      double maxfreq = 0.0;
      for (uint32_t hb = 0; hb < nh; hb += 32) {
        const uint32_t k = hb + lane;
        uint32_t r = RES_REJECT, g = 0;
        if (k < nh) {
          r = res[base + k];
          g = hq[k];
        }
        const bool valid = r != RES_REJECT;
        const uint32_t ld = r & 0xFF;
        bool c_lower = false;
        uint32_t vocab = 0;
        double freq = 1.0;
        if (valid) {
          c_lower = (__ldg(rows + (size_t)g * nstride + 1) & ROW_FIRST_LOWER) != 0;
          vocab = __ldg(ix->inst_vocab + g);
          if (have_freq) freq = (double)__ldg(ix->inst_freq + g);
        }
        // features are skipped (0 / true) when their weight is <= 0 (src/lib.rs:1352-1377)
        const uint32_t f_lcs = (valid && bp.w_lcs > 0.0) ? (r >> 8) & 0xFF : 0;
        const uint32_t f_pre = (valid && bp.w_prefix > 0.0) ? (r >> 16) & 0xFF : 0;
        const uint32_t f_suf = (valid && bp.w_suffix > 0.0) ? r >> 24 : 0;
        const bool samecase = bp.w_case > 0.0 ? (c_lower == q_lower) : true;
        // ---- f64 score, left to right, no FMA (src/lib.rs:1433-1452) -----------------------------------
        double q_ld, q_lcs, q_pre, q_suf;
        if (quot_ok) {  // warp-uniform; all four indices are <= Lq <= 31 (ld is clamped: it only matters when <= Lq)
          q_ld = __shfl_sync(FULL, quot_lane, min(ld, 31u));
          q_lcs = __shfl_sync(FULL, quot_lane, f_lcs);
          q_pre = __shfl_sync(FULL, quot_lane, f_pre);
          q_suf = __shfl_sync(FULL, quot_lane, f_suf);
        } else {
          q_ld = __ddiv_rn((double)ld, Ld);
          q_lcs = __ddiv_rn((double)f_lcs, Ld);
          q_pre = __ddiv_rn((double)f_pre, Ld);
          q_suf = __ddiv_rn((double)f_suf, Ld);
        }
        const double ds = ld > Lq ? 0.0 : __dsub_rn(1.0, q_ld);
        double acc = __dmul_rn(bp.w_ld, ds);
        acc = __dadd_rn(acc, __dmul_rn(bp.w_lcs, q_lcs));
        acc = __dadd_rn(acc, __dmul_rn(bp.w_prefix, q_pre));
        acc = __dadd_rn(acc, __dmul_rn(bp.w_suffix, q_suf));
        acc = __dadd_rn(acc, samecase ? bp.w_case : 0.0);
        const double score = __ddiv_rn(acc, bp.w_sum);
        // max_freq is taken over every instance within the edit distance, before the score threshold
        double mf = valid ? freq : 0.0;
        for (int o = 16; o > 0; o >>= 1) mf = fmax(mf, __shfl_xor_sync(FULL, mf, o));
        maxfreq = fmax(maxfreq, mf);
        c_surv += valid ? 1 : 0;
        const bool keep = valid && score >= bp.score_threshold;
        const uint32_t kmask = __ballot_sync(FULL, keep);
        if (keep) {
          SurvRec sr;
          sr.dist = score;
          sr.freq = freq;
          sr.key = 0.0;
          sr.g = gid_of ? __ldg(gid_of + g) : g;  // sharded index: global gather id
          sr.raw = (uint32_t)freq;               // raw frequency (exact: u32 or 1.0)
          sr.vocab = vocab;
          sr.pad = 0;
          surv[nsurv + __popc(kmask & lanemask_lt())] = sr;
        }
        nsurv += __popc(kmask);
        __syncwarp();
      }
      ConfStage cs;
      if (rec_query && (bp.finish_mode == FINISH_CROP || bp.finish_mode == FINISH_GATHER)) {
        cs.rec_query = rec_query;
        cs.qrow = q;
      }
      c_res += rank_crop_emit(bp, rce_mode(bp.finish_mode), surv, sorted, nsurv, maxfreq, out, out_gid, out_head, qi, flags,
                              qflags, work + 2, 0, cs);
    }
  }
  if (counters) {
    for (int o = 16; o > 0; o >>= 1) {
      c_surv += __shfl_xor_sync(FULL, c_surv, o);
      c_res += __shfl_xor_sync(FULL, c_res, o);
    }
    if (lane == 0) {
      atomicAdd(&counters->survivors, c_surv);
      atomicAdd(&counters->results, c_res);
    }
  }
}

// ================================================================================================
// launchers
// ================================================================================================
static int g_k1_ctas_per_sm = 0;
static int g_k1_variant = 4;  // resident CTAs per SM the probe kernel is compiled for (register budget); ANL_K1_CTAS=3|4
typedef void (*ProbeFn)(const DeviceIndex*, const BatchParams, const uint8_t*, const uint32_t*, uint32_t, uint32_t*, uint32_t*,
                        uint32_t*, unsigned int*, Counters*);
static ProbeFn probe_fn() { return g_k1_variant == 3 ? probe_kernel<3> : probe_kernel<ANL_K1_MIN_CTAS>; }

static uint32_t ring_depth(const BatchParams& bp) {
  // rows needed = max edit distance + 2; thresholds are capped at 255 but anything beyond 14 is
  // rejected by the host (ANL_ERR_UNSUPPORTED)
  uint32_t kmax = bp.max_edit.kind == 0 ? 12u : (bp.max_edit.value & 0xFFu);
  if (kmax > 14) kmax = 14;
  return kmax + 2;
}

cudaError_t configure_kernels() {
  if (const char* v = getenv("ANL_K1_CTAS")) g_k1_variant = atoi(v) == 3 ? 3 : 4;
  cudaError_t e = cudaFuncSetAttribute(probe_fn(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Shared));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(dp_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_k1_ctas_per_sm, probe_fn(), K1_WARPS * 32, sizeof(K1Shared));
  return e;
}

cudaError_t launch_encode(const DeviceIndex* d_ix, const BatchParams& bp, const uint8_t* qblob, const uint32_t* qboff, uint32_t n,
                          uint8_t* rows, uint8_t* status, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  encode_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_ix, bp, qblob, qboff, n, rows, status);
  ++g_kernel_launches;
  return cudaGetLastError();
}

cudaError_t launch_probe(const DeviceIndex* d_ix, const DeviceIndex& h_ix, const BatchParams& bp, const LaunchBuffers& lb,
                         int sm_count, cudaStream_t stream) {
  (void)h_ix;
  if (lb.n == 0) return cudaSuccess;
  // the split path does not serve StopAtExactMatch (the enumeration depends on the exact stage's answer)
  const bool split = lb.queue != nullptr && lb.queue_cap > 0 && lb.qctx != nullptr && !bp.stop_at_exact;
  cudaError_t e = cudaMemsetAsync(lb.work, 0, sizeof(unsigned int), stream);
  if (e != cudaSuccess) return e;
  if (split) {
    e = cudaMemsetAsync(lb.work + 4, 0, 3 * sizeof(unsigned int), stream);  // queue length, exact-stage work counter, Bloom phase 1
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(lb.hit_count, 0, (size_t)lb.n * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(lb.qflags, 0, (size_t)lb.n * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
  }
  if (split) {
    static int kb_ctas = 0;
    if (kb_ctas == 0) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&kb_ctas, bloom_kernel, KB_WARPS * 32, 0) != cudaSuccess || kb_ctas < 1)
        kb_ctas = 1;
    }
    long long want = ((long long)lb.n + KB_WARPS - 1) / KB_WARPS;
    long long grid = (long long)sm_count * kb_ctas;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    bloom_kernel<<<(unsigned)grid, KB_WARPS * 32, 0, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.n, lb.qflags, lb.work, lb.counters,
                                                               lb.queue, lb.queue_cap, lb.qctx);
    ++g_kernel_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (lb.ev_bloom_done && (e = cudaEventRecord(lb.ev_bloom_done, stream)) != cudaSuccess) return e;
  } else {
    if (lb.ev_bloom_done && (e = cudaEventRecord(lb.ev_bloom_done, stream)) != cudaSuccess) return e;
    int per_sm = g_k1_ctas_per_sm > 0 ? g_k1_ctas_per_sm : 1;
    // persistent grid: a whole number of CTAs per SM, no more warps than queries
    long long want = ((long long)lb.n + K1_WARPS - 1) / K1_WARPS;
    long long grid = (long long)sm_count * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    probe_fn()<<<(unsigned)grid, K1_WARPS * 32, sizeof(K1Shared), stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.n, lb.hits,
                                                                            lb.hit_count, lb.qflags, lb.work, lb.counters);
    ++g_kernel_launches;
    return cudaGetLastError();
  }
  exact_kernel<<<(unsigned)sm_count * 8, KX_WARPS * 32, 0, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.queue, lb.queue_cap, lb.qctx,
                                                                     lb.hits, lb.hit_count, lb.qflags, lb.work, lb.counters);
  ++g_kernel_launches;
  return cudaGetLastError();
}

static int k2_ctas_per_sm(uint32_t R, size_t smem) {
  (void)R;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, score_kernel, K2_WARPS * 32, smem);
  if (e != cudaSuccess || n < 1) n = 1;
  return n;
}

size_t score_scratch_bytes(const BatchParams& bp, int sm_count, uint32_t n_queries) {
  // one survivor list + one sorted copy per resident warp.  Upper bound on resident warps: 16 CTAs/SM
  // exceeds any smem-limited occupancy here; the launcher never starts more warps than queries.
  size_t ctas = (size_t)sm_count * 16;
  const size_t want = ((size_t)n_queries + K2_WARPS - 1) / K2_WARPS;
  if (want < ctas) ctas = std::max<size_t>(want, 1);
  return ctas * K2_WARPS * 2 * bp.hit_cap * sizeof(SurvRec);
}

cudaError_t launch_prefilter(const DeviceIndex* d_ix, const BatchParams& bp, const LaunchBuffers& lb, int sm_count,
                             cudaStream_t stream) {
  if (lb.n <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(lb.work, 0, sizeof(unsigned int), stream);  // the probe kernel's counter is free again
  if (e != cudaSuccess) return e;
  long long grid = (long long)sm_count * 8;  // 64 warps per SM
  const long long want = ((long long)lb.n + KF_WARPS - 1) / KF_WARPS;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  prefilter_kernel<<<(unsigned)grid, KF_WARPS * 32, 0, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.n, lb.hits, lb.hit_count, lb.qflags,
                                                                 lb.work, lb.counters);
  ++g_kernel_launches;
  return cudaGetLastError();
}

static long long score_class_grid(const BatchParams& bp, const LaunchBuffers& lb, int sm_count, uint32_t cols) {
  const uint32_t R = ring_depth(bp);
  const size_t smem = k2_warp_bytes(cols, R) * K2_WARPS;
  if (smem > 200 * 1024) return -1;
  long long grid = (long long)sm_count * k2_ctas_per_sm(R, smem);
  const long long cap = (long long)sm_count * 16;
  if (grid > cap) grid = cap;
  long long want = ((long long)lb.n + K2_WARPS - 1) / K2_WARPS;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  return grid;
}

static cudaError_t launch_score_class(const DeviceIndex* d_ix, const BatchParams& bp, const LaunchBuffers& lb, int sm_count,
                                      cudaStream_t stream, uint32_t cols, uint32_t need_min, uint32_t need_max, unsigned int* work,
                                      unsigned int* work_light, long long grid, uint32_t scratch_cta0) {
  const uint32_t R = ring_depth(bp);
  const size_t smem = k2_warp_bytes(cols, R) * K2_WARPS;
  if (grid < 1) return cudaErrorInvalidConfiguration;
  (void)sm_count;
  SurvRec* scratch = reinterpret_cast<SurvRec*>(lb.scratch);
  score_kernel<<<(unsigned)grid, K2_WARPS * 32, smem, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.rec_query,
                                                                lb.n, lb.hits, lb.hit_count, lb.qflags, lb.out, lb.out_gid,
                                                                lb.out_head, scratch, work, work_light, lb.work + 2, lb.counters,
                                                                cols, R, need_min, need_max, scratch_cta0);
  ++g_kernel_launches;
  return cudaGetLastError();
}

// Queries whose matrices fit K2_SHORT_COLS columns (nearly all of them) run in a launch with a small
// shared-memory footprint and therefore more resident warps; the rest in a second launch sized by the
// longest indexed entry.  Both launches walk the whole batch and skip the queries of the other class.
// With a side stream the second launch (few queries, long dependent chains) runs beside the first on its own
// scratch slots instead of after it.
constexpr uint32_t K2_SHORT_COLS = 24;

cudaError_t launch_score(const DeviceIndex* d_ix, const DeviceIndex& h_ix, const BatchParams& bp, const LaunchBuffers& lb,
                         int sm_count, cudaStream_t stream) {
  if (lb.n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(lb.work + 1, 0, 3 * sizeof(unsigned int), stream);  // work counter, pool cursor, confusable queue
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(lb.work + 6, 0, 2 * sizeof(unsigned int), stream);  // second-phase work counters of the two classes
  if (e != cudaSuccess) return e;
  const uint32_t ML = h_ix.max_len;
  if (ML <= K2_SHORT_COLS + 4)
    return launch_score_class(d_ix, bp, lb, sm_count, stream, ML, 0, ML, lb.work + 1, lb.work + 6, score_class_grid(bp, lb, sm_count, ML), 0);
  const long long grid_a = score_class_grid(bp, lb, sm_count, K2_SHORT_COLS);
  // the longest query of the batch bounds what the second class can need (a symbol takes at least one byte)
  const uint32_t longest = bp.query_stride - 2, kmax = ring_depth(bp) - 2;
  const bool second = std::min<uint32_t>(ML, longest + kmax) > K2_SHORT_COLS;
  long long grid_b = second ? score_class_grid(bp, lb, sm_count, ML) : 0;
  if (second && grid_b < 1) return cudaErrorInvalidConfiguration;
  // side by side only when the scratch has slots for both grids (the long class then gets by with one CTA per SM)
  bool beside = second && lb.aux_stream && lb.ev_fork && lb.ev_join && grid_a > 0;
  if (beside) {
    if (grid_b > sm_count) grid_b = sm_count;
    const size_t slot = (size_t)K2_WARPS * 2 * bp.hit_cap * sizeof(SurvRec);
    beside = (size_t)(grid_a + grid_b) * slot <= lb.scratch_bytes;
    if (!beside) grid_b = score_class_grid(bp, lb, sm_count, ML);
  }
  if (beside) {
    // fork: the memset above (pool cursor, queue length) precedes both launches
    if ((e = cudaEventRecord(lb.ev_fork, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(lb.aux_stream, lb.ev_fork, 0)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(lb.work, 0, sizeof(unsigned int), lb.aux_stream)) != cudaSuccess) return e;
    e = launch_score_class(d_ix, bp, lb, sm_count, lb.aux_stream, ML, K2_SHORT_COLS + 1, ML, lb.work, lb.work + 7, grid_b, (uint32_t)grid_a);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventRecord(lb.ev_join, lb.aux_stream)) != cudaSuccess) return e;
  }
  e = launch_score_class(d_ix, bp, lb, sm_count, stream, K2_SHORT_COLS, 0, K2_SHORT_COLS, lb.work + 1, lb.work + 6, grid_a, 0);
  if (e != cudaSuccess) return e;
  if (beside) return cudaStreamWaitEvent(stream, lb.ev_join, 0);
  if (!second) return cudaSuccess;
  e = cudaMemsetAsync(lb.work, 0, sizeof(unsigned int), stream);  // the probe kernel's counter is free again
  if (e != cudaSuccess) return e;
  return launch_score_class(d_ix, bp, lb, sm_count, stream, ML, K2_SHORT_COLS + 1, ML, lb.work, lb.work + 7, grid_b, 0);
}

// The pair-list score stage: filter + shape histogram, scan, scatter, DP per shape class, rank.  work[8..17] are its
// counters (zeroed here); the caller checks work[PW_TOTAL] <= lb.pair_cap afterwards (else: run the stage again with a
// larger pair list).
cudaError_t launch_score_pairs(const DeviceIndex* d_ix, const DeviceIndex& h_ix, const BatchParams& bp, const LaunchBuffers& lb,
                               int sm_count, cudaStream_t stream, cudaEvent_t ev_filter_done) {
  if (lb.n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(lb.work + 1, 0, 3 * sizeof(unsigned int), stream);  // (unused), pool cursor, confusable queue
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(lb.work + 8, 0, 10 * sizeof(unsigned int), stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(lb.pair_hist, 0, PAIR_BUCKETS * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  auto grid_for = [&](int warps_per_cta, int ctas_per_sm) {
    long long grid = (long long)sm_count * ctas_per_sm;
    const long long want = ((long long)lb.n + warps_per_cta - 1) / warps_per_cta;
    if (grid > want) grid = want;
    return (unsigned)(grid < 1 ? 1 : grid);
  };
  pairfilter_kernel<<<grid_for(PF_WARPS, 6), PF_WARPS * 32, 0, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.n, lb.hits, lb.hit_count,
                                                                        lb.qflags, lb.qbase, lb.pair_hist, lb.work, lb.counters);
  pairscan_kernel<<<1, 1024, 0, stream>>>(lb.pair_hist, lb.pair_first, lb.pair_cursor, lb.pairs, lb.pair_cap, lb.work);
  pairscatter_kernel<<<grid_for(PS_WARPS, 8), PS_WARPS * 32, 0, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.n, lb.hits, lb.hit_count,
                                                                          lb.qflags, lb.qbase, lb.pair_first, lb.pair_cursor,
                                                                          lb.pairs, lb.pair_cap, lb.work);
  g_kernel_launches += 3;
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (ev_filter_done && (e = cudaEventRecord(ev_filter_done, stream)) != cudaSuccess) return e;
  const uint32_t R = ring_depth(bp);
  static int use_tma = -1;
  if (use_tma < 0) {
    const char* v = getenv("ANL_DP_TMA");
    use_tma = v ? (atoi(v) != 0) : 0;  // (measured, profiles/r02d: the staged variant is 30 % slower -- see DESIGN.md)
  }
  // short shapes: both sides <= PAIR_SHORT_MAX (nearly all pairs); long shapes: sized by the longest entry and by the
  // longest query that can still have a candidate (length check: |Lq - Lc| <= max edit distance)
  const uint32_t longest_q = std::min<uint32_t>(std::min<uint32_t>(bp.query_stride - 2, (uint32_t)ANL_MAX_SYMBOLS), h_ix.max_len + (R - 2));
  for (int cls = 0; cls < 2; ++cls) {
    if (cls == 1 && std::max(longest_q, h_ix.max_len) <= PAIR_SHORT_MAX) break;
    const uint32_t MQ = cls == 0 ? std::min(PAIR_SHORT_MAX, longest_q) : longest_q;
    const uint32_t MC = cls == 0 ? std::min(PAIR_SHORT_MAX, h_ix.max_len) : h_ix.max_len;
    // bytes of a row the DP can need: length + flags + symbols, in 16-byte units (what a bulk copy moves)
    const uint32_t qb = std::min<uint32_t>(bp.query_stride, (MQ + 2 + 15) & ~15u), cb = std::min<uint32_t>(h_ix.norm_stride, (MC + 2 + 15) & ~15u);
    const size_t per_warp = use_tma ? dp_tma_warp_bytes(MQ, MC, R, qb, cb) : dp_warp_bytes(MQ, MC, R);
    int warps = DP_WARPS;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    const size_t smem = per_warp * warps;
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    int ctas = 0;
    if (use_tma)
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, dp_tma_kernel, warps * 32, smem);
    else
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, dp_kernel, warps * 32, smem);
    if (e != cudaSuccess || ctas < 1) ctas = 1;
    if (cls == 1) ctas = std::min(ctas, 2);  // (few tiles, long dependent chains)
    const unsigned grid = (unsigned)std::min<long long>((long long)sm_count * ctas, (long long)sm_count * 16);
    if (use_tma)
      dp_tma_kernel<<<grid, warps * 32, smem, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.pairs, lb.pair_cap, lb.pair_res, lb.work, cls,
                                                        lb.counters, MQ, MC, R, qb, cb);
    else
      dp_kernel<<<grid, warps * 32, smem, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.pairs, lb.pair_cap, lb.pair_res, lb.work, cls,
                                                    lb.counters, MQ, MC, R);
    ++g_kernel_launches;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  {
    long long grid = (long long)sm_count * 8;
    const long long want = ((long long)lb.n + K2_WARPS - 1) / K2_WARPS;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    rank_kernel<<<(unsigned)grid, K2_WARPS * 32, 0, stream>>>(d_ix, bp, lb.queries, lb.qlist, lb.rec_query, lb.n, lb.hits, lb.hit_count,
                                                              lb.qflags, lb.qbase, lb.pair_res, lb.pair_cap, lb.out, lb.out_gid,
                                                              lb.out_head, reinterpret_cast<SurvRec*>(lb.scratch), lb.work,
                                                              lb.counters, rank_grab(lb.n, grid));
    ++g_kernel_launches;
  }
  return cudaGetLastError();
}

// The device confusable stage: triage of every emitted record, edit scripts of the queued pairs; launch_finish then
// re-ranks / crops / cuts off per query.  Only launched when the score kernel recorded the records' queries
// (lb.rec_query); without lb.conf_work the triage only marks the settled records for the host post-pass.
cudaError_t launch_confusables(const DeviceIndex* d_ix, const BatchParams& bp, const LaunchBuffers& lb, int sm_count,
                               cudaStream_t stream) {
  if (lb.n == 0 || !lb.rec_query || !lb.qblob || !(bp.finish_mode == FINISH_CROP || bp.finish_mode == FINISH_GATHER))
    return cudaSuccess;
  {
    unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)bp.pool_cap + 255) / 256, (uint64_t)sm_count * 8);
    if (blocks < 1) blocks = 1;
    cudaError_t e = cudaMemsetAsync(lb.work + 7, 0, sizeof(unsigned int), stream);  // length of the wide queue
    if (e != cudaSuccess) return e;
    triage_kernel<<<blocks, 256, 0, stream>>>(d_ix, lb.qblob, lb.qboff, lb.rec_query, lb.out, lb.work + 2, bp.pool_cap, lb.conf_work,
                                              lb.work + 3, lb.work + 7);
    ++g_kernel_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess || !lb.conf_work) return e;
  }
  // one thread per possible work item (the queue length is only known on the device): threads beyond the
  // queue exit at once, and the long, divergent per-pair work is balanced by the block scheduler
  unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)bp.pool_cap + 63) / 64, (uint64_t)sm_count * 1024);
  if (blocks < 1) blocks = 1;
  // Few pairs hold characters beyond ASCII, but a single diff is a long dependent chain (0.1-0.3 ms for one thread): on
  // its own the wide kernel took 0.3 ms whatever the batch size -- per 131 072-query chunk as much as the byte kernel.
  // It runs beside the byte kernel on the side stream (the two queues and the records they settle are disjoint).
  const bool beside = lb.aux_stream && lb.ev_fork && lb.ev_join;
  cudaStream_t wide_stream = stream;
  cudaError_t e;
  if (beside) {
    if ((e = cudaEventRecord(lb.ev_fork, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(lb.aux_stream, lb.ev_fork, 0)) != cudaSuccess) return e;
    wide_stream = lb.aux_stream;
  }
  confusable_wide_kernel<<<(unsigned)sm_count * 8, 64, 0, wide_stream>>>(d_ix, lb.qblob, lb.qboff, lb.conf_work, lb.work + 7, bp.pool_cap,
                                                                         lb.out);
  if (beside && (e = cudaEventRecord(lb.ev_join, lb.aux_stream)) != cudaSuccess) return e;
  confusable_kernel<<<blocks, 64, 0, stream>>>(d_ix, lb.qblob, lb.qboff, lb.conf_work, lb.work + 3, bp.pool_cap, lb.out);
  g_kernel_launches += 2;
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (beside) return cudaStreamWaitEvent(stream, lb.ev_join, 0);
  return cudaSuccess;
}
cudaError_t launch_finish(const BatchParams& bp, const LaunchBuffers& lb, int sm_count, cudaStream_t stream) {
  if (lb.n == 0 || !lb.conf_work) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(lb.work, 0, sizeof(unsigned int), stream);  // the probe kernel's counter is free again
  if (e != cudaSuccess) return e;
  // the scratch is the score kernel's (score_scratch_bytes): at most sm_count * 16 CTAs of K2_WARPS warps
  long long grid = (long long)sm_count * 8;
  const long long want = ((long long)lb.n + K2_WARPS - 1) / K2_WARPS;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  finish_kernel<<<(unsigned)grid, K2_WARPS * 32, 0, stream>>>(bp, lb.n, lb.out, lb.out_head, reinterpret_cast<SurvRec*>(lb.scratch),
                                                              bp.hit_cap, lb.work);
  ++g_kernel_launches;
  return cudaGetLastError();
}

}  // namespace anl
