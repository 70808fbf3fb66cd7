// engine.cu -- device residency of the index and the batched lookup pipeline.
//
// Replaces the reference's per-query call chain find_variants -> find_nearest_anahashes ->
// gather_instances -> score_and_rank (src/lib.rs:972-1027) with: one H2D copy of a batch's raw text, the
// kernels of kernels.cu (encode, Bloom stage, exact stage, prefilter, score/rank, confusables, finish), one
// packed D2H copy, and a thin host pass that assembles the result arrays and finishes the few queries the
// device declines (confusable pairs with non-ASCII or over-long text, src/lib.rs:1592-1622).  There is no CPU
// fallback: every failure of the CUDA path is an error.
#include "engine.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <type_traits>

#include "hostpool.h"
#include "unicode_tables.h"

namespace anl {

void confusable_stats(uint64_t* v);

// ---- recycler of large host blocks (see engine.h) -------------------------------------------------------
namespace {
struct ParkedBlock {
  void* p;
  size_t bytes;
};
std::mutex g_park_m;
ParkedBlock g_parked[6] = {{nullptr, 0}, {nullptr, 0}, {nullptr, 0}, {nullptr, 0}, {nullptr, 0}, {nullptr, 0}};
const size_t kParkMin = 4u << 20, kParkMax = 2048ull << 20;
}  // namespace
void* big_block_take(size_t min_bytes, size_t* got_bytes) {
  if (min_bytes < kParkMin) return nullptr;
  std::lock_guard<std::mutex> lk(g_park_m);
  for (ParkedBlock& b : g_parked) {
    if (b.p && b.bytes >= min_bytes && b.bytes / 4 <= min_bytes) {
      void* r = b.p;
      *got_bytes = b.bytes;
      b.p = nullptr;
      b.bytes = 0;
      return r;
    }
  }
  return nullptr;
}
void big_block_give(void* p, size_t bytes) {
  if (bytes >= kParkMin && bytes <= kParkMax) {
    std::lock_guard<std::mutex> lk(g_park_m);
    ParkedBlock* slot = nullptr;
    for (ParkedBlock& b : g_parked)
      if (!b.p) slot = &b;
    if (!slot) {  // replace the smallest parked block if this one is bigger
      slot = &g_parked[0];
      for (ParkedBlock& b : g_parked)
        if (b.bytes < slot->bytes) slot = &b;
      if (slot->bytes >= bytes) slot = nullptr;
    }
    if (slot) {
      void* old = slot->p;
      slot->p = p;
      slot->bytes = bytes;
      p = old;
    }
  }
  free(p);
}

// ---- DMA-able host blocks for the result arrays (hostpool.h) -------------------------------------------------
namespace {
struct DmaParked {
  void* p;
  size_t bytes;
  bool pinned;
};
std::mutex g_dma_m;
DmaParked g_dma_parked[12];
}  // namespace
void* dma_block_take(size_t min_bytes, size_t* got_bytes, bool* pinned) {
  min_bytes = std::max<size_t>(min_bytes, 256);
  {
    std::lock_guard<std::mutex> lk(g_dma_m);
    DmaParked* best = nullptr;
    for (DmaParked& b : g_dma_parked)
      if (b.p && b.bytes >= min_bytes && b.bytes <= std::max(4 * min_bytes, min_bytes + (4u << 20)) && (!best || b.bytes < best->bytes))
        best = &b;
    if (best) {
      void* r = best->p;
      *got_bytes = best->bytes;
      *pinned = best->pinned;
      best->p = nullptr;
      best->bytes = 0;
      return r;
    }
  }
  const size_t two_mb = (size_t)2 << 20;
  const size_t want = min_bytes >= two_mb ? (min_bytes + two_mb - 1) & ~(two_mb - 1) : (min_bytes + 4095) & ~(size_t)4095;
  void* p = nullptr;
  if (cudaHostAlloc(&p, want, cudaHostAllocPortable) == cudaSuccess && p) {
    *pinned = true;
    *got_bytes = want;
    return p;
  }
  cudaGetLastError();  // (no device / pinning limit reached: plain memory, the copies are then staged by the driver)
  p = malloc(want);
  *pinned = false;
  *got_bytes = p ? want : 0;
  return p;
}
void dma_block_give(void* p, size_t bytes, bool pinned) {
  if (!p) return;
  if (bytes <= kParkMax) {
    std::lock_guard<std::mutex> lk(g_dma_m);
    DmaParked* slot = nullptr;
    for (DmaParked& b : g_dma_parked)
      if (!b.p) slot = &b;
    if (!slot) {  // replace the smallest parked block if this one is bigger
      slot = &g_dma_parked[0];
      for (DmaParked& b : g_dma_parked)
        if (b.bytes < slot->bytes) slot = &b;
      if (slot->bytes >= bytes) slot = nullptr;
    }
    if (slot) {
      std::swap(p, slot->p);
      std::swap(bytes, slot->bytes);
      std::swap(pinned, slot->pinned);
    }
  }
  if (!p) return;
  if (pinned)
    cudaFreeHost(p);
  else
    free(p);
}

// DistanceThreshold on the host (src/lib.rs:982-1012); the kernels carry their own copy.
static uint32_t host_threshold(const anl_distance_threshold& t, size_t len) {
  auto sat = [](double v) -> uint32_t { return !(v == v) || v <= 0 ? 0u : (v >= 255 ? 255u : (uint32_t)v); };
  if (t.kind == ANL_THRESHOLD_ABSOLUTE) return std::min<uint32_t>(t.value & 0xFF, sat((double)(len / 2)));
  const uint32_t v = sat(std::floor((float)len * t.ratio));
  return std::min<uint32_t>(v, t.kind == ANL_THRESHOLD_RATIO ? 12u : (t.value & 0xFF));
}

#define CU_TRY(expr)                                                                       \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      *err = std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr;          \
      return false;                                                                        \
    }                                                                                      \
  } while (0)

template <class T>
static bool dev_realloc(T** p, size_t count, std::string* err) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  void* q = nullptr;
  CU_TRY(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  *p = reinterpret_cast<T*>(q);
  return true;
}
template <class T>
static bool pinned_realloc(T** p, size_t count, std::string* err) {
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  void* q = nullptr;
  CU_TRY(cudaMallocHost(&q, std::max<size_t>(count, 1) * sizeof(T)));
  *p = reinterpret_cast<T*>(q);
  return true;
}

Engine::~Engine() {
  shard_comm_free();
  for (DeviceBatch* b : cache_) destroy_batch(b);
  cache_.clear();
  release_index();
  if (stream_) cudaStreamDestroy(stream_);
}

void Engine::release_index() {
  for (void* p : index_allocs_) cudaFree(p);
  index_allocs_.clear();
  for (void* p : conf_allocs_) cudaFree(p);
  conf_allocs_.clear();
  conf_uploaded_ = (size_t)-1;
  if (d_mset_) cudaFree(d_mset_);
  d_mset_ = nullptr;
  if (d_ix_) cudaFree(d_ix_);
  d_ix_ = nullptr;
}

template <class V, class T>
static bool upload_vec(const V& v, const T** out, std::vector<void*>* allocs, std::string* err) {
  static_assert(std::is_same<typename V::value_type, T>::value, "element type mismatch");
  void* p = nullptr;
  CU_TRY(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
  allocs->push_back(p);
  if (!v.empty()) CU_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const T*>(p);
  return true;
}

bool Engine::upload(int device, std::string* err) {
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0) {
    *err = std::string("no CUDA device available (") + cudaGetErrorString(ce) +
           "); the variant-lookup path has no CPU fallback";
    return false;
  }
  if (device >= 0) CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaGetDevice(&device_));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device_));
  sm_count_ = prop.multiProcessorCount;
  if (prop.major < 10) {
    *err = std::string("device ") + prop.name + " is not sm_100 class; this library is built for sm_100a only";
    return false;
  }
  if (!stream_) CU_TRY(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  CU_TRY(configure_kernels());
  for (DeviceBatch* b : cache_) destroy_batch(b);
  cache_.clear();
  release_index();

  const HostIndex& hx = hm_->index;
  DeviceIndex ix;
  memset(&ix, 0, sizeof ix);
  if (!upload_vec(hx.table, &ix.table, &index_allocs_, err)) return false;
  ix.table_mask = hx.table.size() - 1;
  if (!upload_vec(hx.bloom, &ix.bloom, &index_allocs_, err)) return false;
  ix.bloom_mask = hx.bloom.size() - 1;
  if (!upload_vec(hx.post_ana, &ix.post_ana, &index_allocs_, err)) return false;
  if (!upload_vec(hx.post_cls, &ix.post_cls, &index_allocs_, err)) return false;
  ix.sd = hx.sd;
  {
    // keys + instance offsets -> 32-byte anagram records, single postings -> their slots (device-only forms)
    const Key192* d_key = nullptr;
    const uint32_t* d_off = nullptr;
    std::vector<void*> tmp;
    bool ok = upload_vec(hx.ana_key, &d_key, &tmp, err) && upload_vec(hx.ana_inst_off, &d_off, &tmp, err);
    void* rec = nullptr;
    if (ok && cudaMalloc(&rec, std::max<size_t>(hx.ana_key.size(), 1) * sizeof(AnaRec)) != cudaSuccess) {
      *err = "out of device memory for the anagram records";
      ok = false;
    }
    if (ok) {
      index_allocs_.push_back(rec);
      ix.ana_rec = reinterpret_cast<const AnaRec*>(rec);
      const cudaError_t ce2 = finish_device_index(d_key, d_off, (uint32_t)hx.ana_key.size(), reinterpret_cast<AnaRec*>(rec),
                                                  const_cast<Slot*>(ix.table), hx.table.size(), ix.post_ana, ix.post_cls, stream_);
      if (ce2 != cudaSuccess || cudaStreamSynchronize(stream_) != cudaSuccess) {
        *err = std::string("index finishing kernels failed: ") + cudaGetErrorString(ce2 != cudaSuccess ? ce2 : cudaGetLastError());
        ok = false;
      }
    }
    for (void* p : tmp) cudaFree(p);
    if (!ok) return false;
  }
  if (!upload_vec(hx.inst_rows, &ix.inst_rows, &index_allocs_, err)) return false;
  ix.norm_stride = hx.norm_stride;
  if (!upload_vec(hx.inst_vocab, &ix.inst_vocab, &index_allocs_, err)) return false;
  if (!upload_vec(hx.inst_freq, &ix.inst_freq, &index_allocs_, err)) return false;
  ix.inst_gid = nullptr;
  if (!hx.inst_gid.empty() && !upload_vec(hx.inst_gid, &ix.inst_gid, &index_allocs_, err)) return false;
  // saturating binomials C(n, k), n < 256, k < 8
  std::vector<uint32_t> binom(256 * 8, 0);
  for (int n = 0; n < 256; ++n) {
    binom[n * 8 + 0] = 1;
    for (int k = 1; k < 8; ++k) {
      uint64_t v = n == 0 ? 0 : (uint64_t)binom[(n - 1) * 8 + k - 1] + binom[(n - 1) * 8 + k];
      binom[n * 8 + k] = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
    }
  }
  if (!upload_vec(binom, &ix.binom, &index_allocs_, err)) return false;
  // colex unranking tables: nested loops visit the subsets in colex (combinadic) rank order
  std::vector<uint32_t> colex2, colex3;
  for (uint32_t p1 = 1; p1 < (uint32_t)COLEX_N; ++p1)
    for (uint32_t p0 = 0; p0 < p1; ++p0) colex2.push_back(p0 | (p1 << 8));
  for (uint32_t p2 = 2; p2 < (uint32_t)COLEX_N; ++p2)
    for (uint32_t p1 = 1; p1 < p2; ++p1)
      for (uint32_t p0 = 0; p0 < p1; ++p0) colex3.push_back(p0 | (p1 << 8) | (p2 << 16));
  if (!upload_vec(colex2, &ix.colex2, &index_allocs_, err)) return false;
  if (!upload_vec(colex3, &ix.colex3, &index_allocs_, err)) return false;
  // alphabet tables for the encode kernel (queries are normalised on the device unless a member does not fit)
  {
    std::vector<AlphaMember> members;
    std::vector<AlphaFirst> first;
    const char* off = getenv("ANL_HOST_ENCODE");
    if (!(off && atoi(off)) && hm_->alphabet.export_tables(&members, &first)) {
      if (members.empty()) members.push_back(AlphaMember{});
      std::vector<uint32_t> lower(2 * (size_t)anl_unicode::kLowercaseRanges_len);
      for (unsigned i = 0; i < anl_unicode::kLowercaseRanges_len; ++i) {
        lower[2 * i] = anl_unicode::kLowercaseRanges[i][0];
        lower[2 * i + 1] = anl_unicode::kLowercaseRanges[i][1];
      }
      if (!upload_vec(members, &ix.alpha_members, &index_allocs_, err)) return false;
      if (!upload_vec(first, &ix.alpha_first, &index_allocs_, err)) return false;
      if (!upload_vec(lower, &ix.lower_ranges, &index_allocs_, err)) return false;
      ix.n_lower_ranges = anl_unicode::kLowercaseRanges_len;
      ix.unk_symbol = hm_->alphabet.unk_symbol();
      ix.device_encode = 1;
    }
  }
  memcpy(ix.prime_of, hx.prime_of, sizeof ix.prime_of);
  memcpy(ix.charcount_mask, hx.charcount_mask, sizeof ix.charcount_mask);
  ix.max_charcount = hx.max_charcount;
  ix.max_len = hx.max_len;
  ix.n_anagrams = (uint32_t)hx.ana_key.size();
  ix.n_instances = (uint32_t)hx.inst_vocab.size();
  ix.have_freq = hm_->have_freq ? 1 : 0;
  ix.mset = nullptr;
  h_ix_ = ix;
  void* dp = nullptr;
  CU_TRY(cudaMalloc(&dp, sizeof(DeviceIndex)));
  d_ix_ = reinterpret_cast<DeviceIndex*>(dp);
  CU_TRY(cudaMemcpy(d_ix_, &h_ix_, sizeof h_ix_, cudaMemcpyHostToDevice));
  hm_->index.mset_built_j = 0;
  return true;
}

bool Engine::ensure_msets(uint32_t J, std::string* err) {
  if (J == 0 || (hm_->index.mset_built_j >= J && d_mset_)) return true;
  if (!hm_->ensure_msets(J, err)) return false;
  const HostIndex& hx = hm_->index;
  CU_TRY(cudaDeviceSynchronize());
  if (d_mset_) cudaFree(d_mset_);
  d_mset_ = nullptr;
  CU_TRY(cudaMalloc(&d_mset_, std::max<size_t>(hx.mset.size(), 1) * sizeof(MsetEntry)));
  CU_TRY(cudaMemcpy(d_mset_, hx.mset.data(), hx.mset.size() * sizeof(MsetEntry), cudaMemcpyHostToDevice));
  h_ix_.mset = reinterpret_cast<const MsetEntry*>(d_mset_);
  memcpy(h_ix_.mset_end, hx.mset_end, sizeof h_ix_.mset_end);
  CU_TRY(cudaMemcpy(d_ix_, &h_ix_, sizeof h_ix_, cudaMemcpyHostToDevice));
  return true;
}

// Device copy of what the score kernel needs to prefilter confusable checks: the raw text of every
// vocabulary entry and, per pattern an ASCII pair could satisfy, the character sets its deletion /
// insertion options require.  Rebuilt when confusables or vocabulary changed since the last upload.
bool Engine::ensure_confusable_table(std::string* err) {
  const std::vector<Confusable>& cf = hm_->confusables;
  if (conf_uploaded_ == cf.size() && conf_vocab_ == hm_->decoder.size()) return true;
  CU_TRY(cudaDeviceSynchronize());
  for (void* p : conf_allocs_) cudaFree(p);
  conf_allocs_.clear();
  h_ix_.vocab_text = nullptr;
  h_ix_.vocab_text_off = nullptr;
  h_ix_.conf_pats = nullptr;
  h_ix_.conf_instrs = nullptr;
  h_ix_.conf_opts = nullptr;
  h_ix_.conf_text = nullptr;
  h_ix_.n_conf_pats = 0;
  h_ix_.conf_prefilter = 0;
  h_ix_.conf_all_simple = 0;
  if (!cf.empty()) {
    std::vector<ConfPat> pats;
    std::vector<ConfInstr> instrs;
    std::vector<ConfOpt> opts;
    std::vector<uint16_t> ctext;  // option texts as UTF-16 code units (BMP only: one unit per character)
    bool fits = true;
    for (const Confusable& c : cf) {
      ConfPat pat;
      memset(&pat, 0, sizeof pat);
      pat.weight = c.weight;
      pat.first_instr = (uint16_t)instrs.size();
      pat.strictbegin = c.strictbegin ? 1 : 0;
      pat.strictend = c.strictend ? 1 : 0;
      const size_t instr_mark = instrs.size(), opt_mark = opts.size(), text_mark = ctext.size();
      bool viable = true;
      for (const ConfusableInstr& ins : c.script) {
        ConfInstr ci{(int8_t)ins.op, 0, (uint16_t)opts.size()};
        for (const std::u32string& o : ins.options32) {
          ConfOpt m{0, 0, (uint32_t)ctext.size(), (uint32_t)o.size(), 0, 0};
          bool bmp = true;
          for (char32_t ch : o) {
            if (ch > 0xFFFF) {
              bmp = false;
              break;
            }
            if (ch >= 0x80) m.nonascii = 1;
            else if (ch < 64) m.lo |= 1ull << ch;
            else m.hi |= 1ull << (ch - 64);
          }
          if (!bmp) continue;  // cannot match the text of a pair inside the BMP
          if (ci.n_opts < 255) {
            opts.push_back(m);
            for (char32_t ch : o) ctext.push_back((uint16_t)ch);
            ++ci.n_opts;
          } else {
            fits = false;
          }
        }
        if (ci.n_opts == 0) {
          viable = false;  // every option needs a character outside the BMP: impossible for a BMP pair
          break;
        }
        instrs.push_back(ci);
        ++pat.n_instr;
      }
      if (!viable || pat.n_instr == 0) {
        instrs.resize(instr_mark);
        opts.resize(opt_mark);
        ctext.resize(text_mark);
        continue;
      }
      pats.push_back(pat);
      if (instrs.size() > 60000 || opts.size() > 60000) fits = false;
    }
    ctext.push_back(0);
    if (fits && hm_->decoder.size() < 0x7FFFFFFFull) {
      std::vector<uint8_t> text;
      std::vector<uint32_t> off;
      off.reserve(hm_->decoder.size() + 1);
      size_t total = 0;
      for (const VocabEntry& v : hm_->decoder) total += v.text.size();
      if (total < 0xFFFFFFF0ull) {
        text.reserve(total);
        for (const VocabEntry& v : hm_->decoder) {
          off.push_back((uint32_t)text.size());
          text.insert(text.end(), v.text.begin(), v.text.end());
        }
        off.push_back((uint32_t)text.size());
        if (!upload_vec(text, &h_ix_.vocab_text, &conf_allocs_, err)) return false;
        if (!upload_vec(off, &h_ix_.vocab_text_off, &conf_allocs_, err)) return false;
        if (!upload_vec(pats, &h_ix_.conf_pats, &conf_allocs_, err)) return false;
        if (!upload_vec(instrs, &h_ix_.conf_instrs, &conf_allocs_, err)) return false;
        if (!upload_vec(opts, &h_ix_.conf_opts, &conf_allocs_, err)) return false;
        if (!upload_vec(ctext, &h_ix_.conf_text, &conf_allocs_, err)) return false;
        h_ix_.n_conf_pats = (uint32_t)pats.size();
        h_ix_.conf_prefilter = 1;
        h_ix_.conf_all_simple = hm_->all_confusables_simple ? 1 : 0;
        {
          // character classes of the device edit script (lossless shift): the Alphabetic ranges, per device
          std::vector<uint32_t> alpha(2 * (size_t)anl_unicode::kAlphabeticRanges_len);
          for (unsigned i = 0; i < anl_unicode::kAlphabeticRanges_len; ++i) {
            alpha[2 * i] = anl_unicode::kAlphabeticRanges[i][0];
            alpha[2 * i + 1] = anl_unicode::kAlphabeticRanges[i][1];
          }
          CU_TRY(upload_alphabetic_ranges(alpha.data(), anl_unicode::kAlphabeticRanges_len));
        }
      }
    }
  }
  CU_TRY(cudaMemcpy(d_ix_, &h_ix_, sizeof h_ix_, cudaMemcpyHostToDevice));
  conf_uploaded_ = cf.size();
  conf_vocab_ = hm_->decoder.size();
  return true;
}

static uint32_t threshold_cap(const anl_distance_threshold& t) {
  // largest value the thresholded distance can take for any input (src/lib.rs:982-1012)
  if (t.kind == ANL_THRESHOLD_RATIO) return 12;
  return t.value & 0xFF;
}

bool Engine::make_batch_params(const anl_search_params& p, BatchParams* bp, uint32_t* needed_j, std::string* err) const {
  memset(bp, 0, sizeof *bp);
  for (const anl_distance_threshold* t : {&p.max_anagram_distance, &p.max_edit_distance}) {
    if (t->kind < 0 || t->kind > 2) {
      *err = "invalid distance threshold kind";
      return false;
    }
  }
  bp->max_anagram = Threshold{p.max_anagram_distance.kind, p.max_anagram_distance.ratio, p.max_anagram_distance.value};
  bp->max_edit = Threshold{p.max_edit_distance.kind, p.max_edit_distance.ratio, p.max_edit_distance.value};
  if (threshold_cap(p.max_edit_distance) > 14) {
    *err = "max_edit_distance above 14 is outside the supported range of the GPU path";
    return false;
  }
  bp->max_matches = p.max_matches > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)p.max_matches;
  bp->score_threshold = p.score_threshold;
  bp->cutoff_threshold = p.cutoff_threshold;
  const Weights& w = hm_->weights;
  bp->w_ld = w.ld;
  bp->w_lcs = w.lcs;
  bp->w_prefix = w.prefix;
  bp->w_suffix = w.suffix;
  bp->w_case = w.case_;
  bp->w_sum = w.sum();
  bp->freq_weight64 = (double)p.freq_weight;
  bp->freq_weight_nonzero = p.freq_weight != 0.0f;
  bp->freq_weight_positive = p.freq_weight > 0.0f;
  bp->stop_at_exact = p.stop_criterion == ANL_STOP_AT_EXACT_MATCH;
  if (hm_->confusables.empty())
    bp->finish_mode = FINISH_FULL;
  else
    bp->finish_mode = hm_->confusables_before_pruning ? FINISH_GATHER : FINISH_CROP;
  // variant lists: every survivor goes to the host in gather order; expansion, ranking, crop and cut-off follow there
  if (hm_->any_variants) bp->finish_mode = FINISH_GATHER;
  if (hm_->index.n_shards > 1) bp->finish_mode = FINISH_SHARD;  // ranking happens after the exchange (shard_merge)
  uint32_t hit_cap = 4096;
  if (const char* e = getenv("ANL_HIT_CAP")) hit_cap = (uint32_t)std::max(1, atoi(e));
  bp->hit_cap = hit_cap;
  const uint32_t kcap = std::min<uint32_t>(threshold_cap(p.max_anagram_distance), ANL_MAX_K);
  const uint32_t sd = (uint32_t)hm_->index.sd;
  *needed_j = kcap > sd ? kcap - sd : 0;
  return true;
}

// ---- batches -------------------------------------------------------------------------------------------
void Engine::destroy_batch(DeviceBatch* b) {
  if (!b) return;
  for (void* p : {(void*)b->h_rows, (void*)b->h_head, (void*)b->h_flags, (void*)b->h_hitcnt, (void*)b->h_out,
                  (void*)b->h_work, (void*)b->h_qboff, (void*)b->h_qblob, (void*)b->h_summary})
    if (p) cudaFreeHost(p);
  for (void* p : {(void*)b->d_final, (void*)b->d_loff, (void*)b->d_oflags, (void*)b->d_off64, (void*)b->d_tile_sum,
                  (void*)b->d_summary, (void*)b->rr_conf_work, (void*)b->rr_work, (void*)b->d_rec_query, (void*)b->rr_rec_query,
                  (void*)b->d_qbase, (void*)b->d_pairs, (void*)b->d_pair_res,
                  (void*)b->d_pair_tab})
    if (p) cudaFree(p);
  if (b->d_qblob) cudaFree(b->d_qblob);
  if (b->d_qboff) cudaFree(b->d_qboff);
  if (b->d_conf_work) cudaFree(b->d_conf_work);
  if (b->d_enc_status) cudaFree(b->d_enc_status);
  if (b->d_queue) cudaFree(b->d_queue);
  if (b->d_qctx) cudaFree(b->d_qctx);
  if (b->h_enc_status) cudaFreeHost(b->h_enc_status);
  for (void* p : {(void*)b->d_rows, (void*)b->d_hits, (void*)b->d_hit_count, (void*)b->d_qflags, (void*)b->d_out,
                  (void*)b->d_gid, (void*)b->d_head, b->d_scratch, (void*)b->d_work, (void*)b->d_counters, (void*)b->rr_qlist,
                  (void*)b->rr_hits, (void*)b->rr_hit_count, (void*)b->rr_qflags, (void*)b->rr_head, (void*)b->rr_out,
                  (void*)b->rr_scratch})
    if (p) cudaFree(p);
  for (auto& ev : b->events)
    if (ev) cudaEventDestroy(ev);
  if (b->uploaded) cudaEventDestroy(b->uploaded);
  if (b->ev_fork) cudaEventDestroy(b->ev_fork);
  if (b->ev_done) cudaEventDestroy(b->ev_done);
  if (b->ev_join) cudaEventDestroy(b->ev_join);
  if (b->aux) cudaStreamDestroy(b->aux);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

void Engine::free_batch(DeviceBatch* b) {
  if (!b) return;
  std::unique_lock<std::mutex> lk(cache_m_);
  if (cache_.size() < 8) {
    b->owned_blob.clear();
    b->owned_blob.shrink_to_fit();
    b->blob = nullptr;
    b->ran = false;
    b->runs_recorded = 0;
    cache_.push_back(b);
  } else {
    lk.unlock();
    destroy_batch(b);
  }
}

bool Engine::grow_pool(DeviceBatch* b, uint32_t pool_cap, bool keep_records, std::string* err) {
  const bool need_gid = hm_->index.n_shards > 1;
  const bool need_cw = h_ix_.conf_prefilter && !need_gid;  // the confusable queue can hold every pool record
  const bool need_rq = h_ix_.conf_prefilter != 0;  // the triage runs whenever the model has confusables
  if (pool_cap <= b->cap_pool && b->d_out && b->d_final && (!need_gid || b->d_gid) && (!need_cw || b->d_conf_work) &&
      (!need_rq || b->d_rec_query))
    return true;
  pool_cap = std::max(pool_cap, b->cap_pool);
  if (keep_records && b->d_out && b->cap_pool) {
    // (hit-overflow patch: the records already in the pool stay valid)
    OutRec* bigger = nullptr;
    if (!dev_realloc(&bigger, pool_cap, err)) return false;
    CU_TRY(cudaMemcpy(bigger, b->d_out, (size_t)b->cap_pool * sizeof(OutRec), cudaMemcpyDeviceToDevice));
    cudaFree(b->d_out);
    b->d_out = bigger;
  } else if (!dev_realloc(&b->d_out, pool_cap, err)) {
    return false;
  }
  if (!dev_realloc(&b->d_final, pool_cap, err)) return false;
  if (need_gid && !dev_realloc(&b->d_gid, pool_cap, err)) return false;
  if (need_cw && !dev_realloc(&b->d_conf_work, pool_cap, err)) return false;
  if (need_rq && !dev_realloc(&b->d_rec_query, pool_cap, err)) return false;
  if (!pinned_realloc(&b->h_out, pool_cap, err)) return false;
  b->cap_pool = pool_cap;
  return true;
}

bool Engine::ensure_capacity(DeviceBatch* b, uint32_t n, uint32_t stride, uint32_t hit_cap, uint32_t pool_cap, size_t scratch,
                             std::string* err) {
  if ((size_t)n * stride > b->cap_rows_bytes || !b->d_rows) {
    if (!dev_realloc(&b->d_rows, (size_t)n * stride, err)) return false;
    if (!pinned_realloc(&b->h_rows, (size_t)n * stride, err)) return false;
    b->cap_rows_bytes = (size_t)n * stride;
  }
  if ((size_t)n * hit_cap > b->cap_hits || !b->d_hits) {
    if (!dev_realloc(&b->d_hits, (size_t)n * hit_cap, err)) return false;
    b->cap_hits = (size_t)n * hit_cap;
  }
  if (n > b->cap_n || !b->d_head) {
    if (!dev_realloc(&b->d_hit_count, n, err)) return false;
    if (!dev_realloc(&b->d_qflags, n, err)) return false;
    if (!dev_realloc(&b->d_head, n, err)) return false;
    if (!pinned_realloc(&b->h_head, n, err)) return false;
    if (!pinned_realloc(&b->h_flags, n, err)) return false;
    if (!pinned_realloc(&b->h_hitcnt, n, err)) return false;
    if (!dev_realloc(&b->d_enc_status, (size_t)n + 1, err) || !pinned_realloc(&b->h_enc_status, (size_t)n + 1, err)) return false;
    if (!dev_realloc(&b->d_loff, (size_t)n + 1, err) || !dev_realloc(&b->d_oflags, n, err) || !dev_realloc(&b->d_off64, n, err) ||
        !dev_realloc(&b->d_tile_sum, (size_t)export_tiles(n) + 1, err) || !dev_realloc(&b->d_qbase, n, err))
      return false;
    b->cap_n = n;
  }
  if (!grow_pool(b, pool_cap, false, err)) return false;
  if (scratch > b->cap_scratch || !b->d_scratch) {
    if (!dev_realloc(reinterpret_cast<uint8_t**>(&b->d_scratch), scratch, err)) return false;
    b->cap_scratch = scratch;
  }
  if (!b->d_work) {
    if (!dev_realloc(&b->d_work, WORK_SLOTS, err)) return false;
    if (!pinned_realloc(&b->h_work, WORK_SLOTS, err)) return false;
    CU_TRY(cudaMemset(b->d_work, 0, WORK_SLOTS * sizeof(unsigned int)));
    if (!dev_realloc(&b->d_counters, 1, err)) return false;
    if (!dev_realloc(&b->d_summary, 1, err) || !pinned_realloc(&b->h_summary, 1, err)) return false;
    CU_TRY(cudaMemset(b->d_summary, 0, sizeof(ExportSummary)));
    if (!dev_realloc(&b->d_pair_tab, 3 * (size_t)PAIR_TABLE, err)) return false;
  }
  return true;
}

bool Engine::grow_pairs(DeviceBatch* b, size_t pair_cap, std::string* err) {
  if (pair_cap <= b->cap_pairs && b->d_pairs) return true;
  pair_cap = std::min<size_t>(std::max(pair_cap, b->cap_pairs), 0xFFFFFF00u);
  if (!dev_realloc(&b->d_pairs, pair_cap, err) || !dev_realloc(&b->d_pair_res, pair_cap, err)) return false;
  b->cap_pairs = pair_cap;
  return true;
}

DeviceBatch* Engine::create_batch(const char* blob, const uint64_t* offsets, uint64_t n, const anl_search_params& p,
                                  bool copy_blob, bool sync, std::string* err, int* status) {
  *status = ANL_ERR_CUDA;
  if (!uploaded()) {
    *err = "model has not been built";
    *status = ANL_ERR_NOT_BUILT;
    return nullptr;
  }
  if (n > 0x7FFFFFF0ull) {
    *err = "batch too large; split it";
    *status = ANL_ERR_INVALID;
    return nullptr;
  }
  if (cudaSetDevice(device_) != cudaSuccess) {  // (a dispatcher thread starts on device 0)
    *err = "cudaSetDevice failed";
    return nullptr;
  }
  BatchParams bp;
  uint32_t needed_j = 0;
  {
    static std::mutex setup_m;  // process-wide: the engines of one model share its host-side tables
    std::lock_guard<std::mutex> lk(setup_m);  // (first use of a distance / confusable list builds and uploads tables)
    if (!make_batch_params(p, &bp, &needed_j, err) || !ensure_msets(needed_j, err)) {
      *status = ANL_ERR_UNSUPPORTED;
      return nullptr;
    }
    if (!ensure_confusable_table(err)) return nullptr;
  }
  if (cudaSetDevice(device_) != cudaSuccess) {
    *err = "cudaSetDevice failed";
    return nullptr;
  }
  PhaseTimer pt;
  DeviceBatch* b = nullptr;
  {
    std::lock_guard<std::mutex> lk(cache_m_);
    if (!cache_.empty()) {
      b = cache_.back();
      cache_.pop_back();
    }
  }
  if (!b) b = new DeviceBatch();
  auto fail = [&]() -> DeviceBatch* {
    destroy_batch(b);
    return nullptr;
  };
  if (!b->stream && cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) {
    *err = "cudaStreamCreate failed";
    return fail();
  }
  if (!b->ev_done && cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming) != cudaSuccess) {
    *err = "cudaEventCreate failed";
    return fail();
  }
  if (!b->uploaded && cudaEventCreateWithFlags(&b->uploaded, cudaEventDisableTiming) != cudaSuccess) {
    *err = "cudaEventCreate failed";
    return fail();
  }
  if (!b->aux && (cudaStreamCreateWithFlags(&b->aux, cudaStreamNonBlocking) != cudaSuccess ||
                  cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                  cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming) != cudaSuccess)) {
    *err = "cudaStreamCreate failed";
    return fail();
  }
  b->n = (uint32_t)n;
  b->params = p;
  b->reruns = 0;
  b->settled = false;
  b->results = 0;
  b->offsets.resize(n + 1);
  const uint64_t base = offsets[0];
  for (uint64_t i = 0; i <= n; ++i) b->offsets[i] = offsets[i] - base;
  if (copy_blob) {
    b->owned_blob.assign(blob + base, blob + offsets[n]);
    b->blob = b->owned_blob.data();
  } else {
    b->blob = blob + base;
  }
  // row stride from the longest query in bytes (a symbol consumes at least one byte)
  uint64_t maxbytes = 0;
  for (uint64_t i = 0; i < n; ++i) maxbytes = std::max<uint64_t>(maxbytes, b->offsets[i + 1] - b->offsets[i]);
  const uint32_t stride = (uint32_t)((std::min<uint64_t>(maxbytes, 254) + 2 + 15) & ~15ull);
  bp.query_stride = stride;
  // packed result pool: sized for the common case, grown (and the score kernel re-run) on overflow
  uint64_t per_query = bp.finish_mode == FINISH_GATHER ? 64 : (bp.max_matches ? std::min<uint32_t>(bp.max_matches, 24) : 32);
  if (const char* e = getenv("ANL_POOL_PER_QUERY")) per_query = (uint64_t)std::max(1, atoi(e));
  const uint32_t pool_cap = (uint32_t)std::min<uint64_t>(0xFFFFFF00ull, std::max<uint64_t>(1024, n * per_query));
  if (!ensure_capacity(b, (uint32_t)n, stride, bp.hit_cap, pool_cap, score_scratch_bytes(bp, sm_count_, (uint32_t)n), err)) return fail();
  bp.pool_cap = b->cap_pool;
  {
    // split probe path (Bloom stage -> global queue of staged nodes -> exact stage): capacity per query by the
    // anagram distance; an overflow is detected after the run and answered by the fused kernel
    static int split_on = -1;
    if (split_on < 0) {
      const char* e = getenv("ANL_SPLIT");
      split_on = e ? (atoi(e) != 0) : 1;
    }
    b->split = split_on && !bp.stop_at_exact && n > 0;
    if (b->split) {
      const uint32_t kcap = threshold_cap(p.max_anagram_distance);
      uint64_t per_query = kcap <= 3 ? 160 : (kcap == 4 ? 320 : 640);
      if (const char* e = getenv("ANL_QUEUE_PER_QUERY")) per_query = (uint64_t)std::max(1, atoi(e));
      // (+ one reservation chunk per resident warp of the Bloom stage)
      const uint64_t want = std::min<uint64_t>(0x7FFFFFF0ull, n * per_query + 4096 + 148ull * 64 * 128);
      std::string e2;
      bool okq = true;
      if (want > b->cap_queue || !b->d_queue) {
        okq = dev_realloc(&b->d_queue, (size_t)want, &e2);
        b->cap_queue = okq ? (size_t)want : 0;
      }
      if (okq && (n > b->cap_qctx || !b->d_qctx)) {
        okq = dev_realloc(&b->d_qctx, (size_t)n, &e2);
        b->cap_qctx = okq ? (size_t)n : 0;
      }
      if (!okq) b->split = false;  // not enough memory for the queue: the fused kernel needs none
    }
  }
  {
    // pair-list score stage: room for the pairs that survive the length check and the OSA filter; an overflow is
    // detected after the run and answered by running the stage again with the exact size
    static int pairs_on = -1;
    if (pairs_on < 0) {
      const char* e = getenv("ANL_PAIRS");
      pairs_on = e ? (atoi(e) != 0) : 1;
    }
    b->use_pairs = pairs_on && n > 0;
    if (b->use_pairs) {
      const uint32_t kcap = std::max(threshold_cap(p.max_anagram_distance), threshold_cap(p.max_edit_distance));
      uint64_t per_query = kcap <= 2 ? 96 : (kcap == 3 ? 160 : (kcap == 4 ? 512 : 1024));
      if (const char* e = getenv("ANL_PAIRS_PER_QUERY")) per_query = (uint64_t)std::max(1, atoi(e));
      std::string e2;
      if (!grow_pairs(b, (size_t)std::min<uint64_t>(0xFFFFFF00ull, n * per_query + 4096), &e2)) b->use_pairs = false;
    }
  }
  b->sharded = hm_->index.n_shards > 1;
  b->merged = false;
  b->final_mode = hm_->confusables.empty() ? FINISH_FULL : (hm_->confusables_before_pruning ? FINISH_GATHER : FINISH_CROP);
  if (hm_->any_variants) b->final_mode = FINISH_GATHER;
  b->bp = bp;
  pt.lap("create: buffers");

  // query normalisation (src/anahash.rs:50-80) into fixed-stride rows (len, flags, symbols): on the device
  // from the raw text (encode_kernel), or on the host when the alphabet does not fit the device tables
  b->host_flags.assign(n, 0);
  const uint64_t blob_bytes = b->offsets[n];
  b->dev_encode = h_ix_.device_encode && blob_bytes < 0xFFFFFFF0ull;
  uint8_t* rows = b->h_rows;
  if (!b->dev_encode) {
    const Alphabet& ab = hm_->alphabet;
    const char* text = b->blob;
    parallel_ranges(n, 2048, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) {
        uint8_t* row = rows + (size_t)i * stride;
        const char* s = text + b->offsets[i];
        const size_t len = (size_t)(b->offsets[i + 1] - b->offsets[i]);
        size_t c = ab.encode_into(s, len, row + 2, stride - 2);
        uint8_t flags = 0;
        if (len > 0) {
          // first char lowercase?  (src/lib.rs:1374)
          unsigned char c0 = (unsigned char)s[0];
          uint32_t cp = c0;
          if (c0 >= 0x80) {
            unsigned l = (c0 & 0xE0) == 0xC0 ? 2 : ((c0 & 0xF0) == 0xE0 ? 3 : ((c0 & 0xF8) == 0xF0 ? 4 : 1));
            if (l > len) l = (unsigned)len;
            cp = c0 & (0xFFu >> (l + 1));
            for (unsigned k = 1; k < l; ++k) cp = (cp << 6) | ((unsigned char)s[k] & 0x3F);
          }
          if (anl_unicode::is_lowercase(cp)) flags |= Q_FIRST_LOWER;
        }
        if (c > (size_t)ANL_MAX_SYMBOLS) {
          // longer than the device rows hold.  If even after max_anagram_distance deletions the
          // query is longer than the longest indexed entry, the result is empty (exact); otherwise
          // the query is outside the supported range.
          const uint32_t ka = host_threshold(p.max_anagram_distance, c);
          b->host_flags[i] = (c > (size_t)hm_->index.max_charcount + ka) ? 1 : 2;
          c = 0;
        }
        row[0] = (uint8_t)c;
        row[1] = flags;
      }
    });
    // (host_flags == 2: outside the supported range -- that query alone gets an empty list and flag bit 1)
  }
  pt.lap("create: encode");
  // raw query bytes for the device-side confusable stage (only when a confusable post-pass follows)
  b->has_qblob = false;
  const bool conf_stage = h_ix_.conf_prefilter && (bp.finish_mode == FINISH_CROP || bp.finish_mode == FINISH_GATHER);
  if ((conf_stage || b->dev_encode) && n > 0 && blob_bytes < 0xFFFFFFF0ull) {
    std::string e2;
    bool okb = true;
    if (blob_bytes > b->cap_qblob || !b->d_qblob) {
      // headroom: chunk sizes vary a little and cudaFree synchronises the whole device
      const size_t want = (size_t)blob_bytes + (size_t)blob_bytes / 4 + 4096;
      okb = dev_realloc(&b->d_qblob, want, &e2) && pinned_realloc(&b->h_qblob, want, &e2);
      b->cap_qblob = okb ? want : 0;
    }
    if (okb && (n + 1 > b->cap_qboff || !b->d_qboff)) {
      okb = dev_realloc(&b->d_qboff, (size_t)n + 1, &e2) && pinned_realloc(&b->h_qboff, (size_t)n + 1, &e2);
      b->cap_qboff = okb ? (size_t)n + 1 : 0;
    }
    if (!okb) {
      *err = e2;
      return fail();
    }
    parallel_ranges(n + 1, 1u << 16, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) b->h_qboff[i] = (uint32_t)b->offsets[i];
    });
    parallel_ranges(blob_bytes, 1u << 20, [&](unsigned, uint64_t lo, uint64_t hi) {
      memcpy(b->h_qblob + lo, b->blob + lo, (size_t)(hi - lo));  // pinned staging: the caller's memory is pageable
    });
    if (cudaMemcpyAsync(b->d_qblob, b->h_qblob, (size_t)blob_bytes, cudaMemcpyHostToDevice, b->stream) != cudaSuccess ||
        cudaMemcpyAsync(b->d_qboff, b->h_qboff, ((size_t)n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream) != cudaSuccess) {
      *err = "H2D copy of the query text failed";
      return fail();
    }
    b->has_qblob = true;
  }
  b->dev_conf = conf_stage && b->has_qblob && b->d_conf_work != nullptr && !b->sharded && !hm_->any_variants;
  if (b->dev_encode && !b->has_qblob && n > 0) b->dev_encode = false;  // (cannot happen: the text was staged above)
  pt.lap("create: query text");
  if (b->dev_encode) {
    if (launch_encode(d_ix_, bp, reinterpret_cast<const uint8_t*>(b->d_qblob), b->d_qboff, (uint32_t)n, b->d_rows, b->d_enc_status,
                      b->stream) != cudaSuccess) {
      *err = "encode kernel launch failed";
      return fail();
    }
  } else if (n > 0) {
    // rows encoded on the host; the per-query encode status follows them (the export stage turns it into flags)
    memcpy(b->h_enc_status, b->host_flags.data(), (size_t)n);
    if (cudaMemcpyAsync(b->d_rows, rows, (size_t)n * stride, cudaMemcpyHostToDevice, b->stream) != cudaSuccess ||
        cudaMemcpyAsync(b->d_enc_status, b->h_enc_status, (size_t)n, cudaMemcpyHostToDevice, b->stream) != cudaSuccess) {
      *err = "H2D copy failed";
      return fail();
    }
  }
  if (cudaEventRecord(b->uploaded, b->stream) != cudaSuccess || (sync && cudaStreamSynchronize(b->stream) != cudaSuccess)) {
    *err = "H2D sync failed";
    return fail();
  }
  pt.lap("create: H2D");
  *status = ANL_OK;
  return b;
}

static LaunchBuffers launch_buffers(const DeviceBatch* b) {
  LaunchBuffers lb;
  lb.queries = b->d_rows;
  lb.qlist = nullptr;
  lb.qblob = b->has_qblob ? reinterpret_cast<const uint8_t*>(b->d_qblob) : nullptr;
  lb.qboff = b->has_qblob ? b->d_qboff : nullptr;
  lb.rec_query = b->has_qblob ? b->d_rec_query : nullptr;  // (allocated iff the model has a confusable table)
  lb.conf_work = b->dev_conf ? b->d_conf_work : nullptr;
  if (b->split) {
    lb.queue = b->d_queue;
    lb.queue_cap = (uint32_t)std::min<size_t>(b->cap_queue, 0x7FFFFFF0u);
    lb.qctx = b->d_qctx;
  }
  if (b->use_pairs) {
    lb.qbase = b->d_qbase;
    lb.pairs = b->d_pairs;
    lb.pair_res = b->d_pair_res;
    lb.pair_cap = (uint32_t)b->cap_pairs;
    lb.pair_hist = b->d_pair_tab;
    lb.pair_first = b->d_pair_tab + PAIR_TABLE;
    lb.pair_cursor = b->d_pair_tab + 2 * PAIR_TABLE;
  }
  lb.n = b->n;
  lb.hits = b->d_hits;
  lb.hit_count = b->d_hit_count;
  lb.qflags = b->d_qflags;
  lb.out = b->d_out;
  lb.out_gid = b->sharded ? b->d_gid : nullptr;
  lb.out_head = b->d_head;
  lb.scratch = b->d_scratch;
  lb.work = b->d_work;
  lb.counters = b->d_counters;
  lb.aux_stream = b->aux;
  lb.ev_fork = b->ev_fork;
  lb.ev_join = b->ev_join;
  lb.scratch_bytes = b->cap_scratch;
  return lb;
}

bool Engine::launch_export_chain(DeviceBatch* b, cudaStream_t st, std::string* err) {
  // export stage + the download of what the host needs to decide anything about this pass: pool cursor, staged-node
  // queue length, and the summary (records, queries to re-run, queries to finish on the host)
  if (b->bp.finish_mode != FINISH_SHARD) {  // (a shard's survivors are ranked after the exchange: nothing final yet)
    ExportBuffers eb;
    eb.n = b->n;
    eb.head = b->d_head;
    eb.qflags = b->d_qflags;
    eb.enc_status = b->d_enc_status;
    eb.pool = b->d_out;
    eb.tile_sum = b->d_tile_sum;
    eb.loff = b->d_loff;
    eb.oflags = b->d_oflags;
    eb.out = b->d_final;
    eb.out_cap = b->cap_pool;
    eb.summary = b->d_summary;
    CU_TRY(launch_export(eb, st));
  }
  return true;
}

static bool download_summary(DeviceBatch* b, cudaStream_t st, std::string* err) {
  CU_TRY(cudaMemcpyAsync(b->h_work, b->d_work, WORK_SLOTS * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(b->h_summary, b->d_summary, sizeof(ExportSummary), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaEventRecord(b->ev_done, st));
  return true;
}

// prefilter + score: over the shape-sorted pair list, or query by query (ANL_PAIRS=0, re-runs of overflowed queries)
bool Engine::launch_score_stage(DeviceBatch* b, const LaunchBuffers& lb, cudaStream_t st, cudaEvent_t ev_filter, std::string* err) {
  if (b->use_pairs && lb.pairs) {
    CU_TRY(launch_score_pairs(d_ix_, h_ix_, b->bp, lb, sm_count_, st, ev_filter));
    return true;
  }
  CU_TRY(launch_prefilter(d_ix_, b->bp, lb, sm_count_, st));
  if (ev_filter) CU_TRY(cudaEventRecord(ev_filter, st));
  CU_TRY(launch_score(d_ix_, h_ix_, b->bp, lb, sm_count_, st));
  return true;
}

bool Engine::relaunch_from_score(DeviceBatch* b, std::string* err) {
  cudaStream_t st = b->stream;
  CU_TRY(cudaStreamWaitEvent(st, b->ev_done, 0));
  LaunchBuffers lb = launch_buffers(b);
  lb.counters = nullptr;
  if (!launch_score_stage(b, lb, st, nullptr, err)) return false;
  CU_TRY(launch_confusables(d_ix_, b->bp, lb, sm_count_, st));
  CU_TRY(launch_finish(b->bp, lb, sm_count_, st));
  return launch_export_chain(b, st, err) && download_summary(b, st, err);
}

bool Engine::run_batch(DeviceBatch* b, cudaStream_t stream, std::string* err) {
  CU_TRY(cudaSetDevice(device_));
  if (!stream)
    stream = b->stream;
  else
    CU_TRY(cudaStreamWaitEvent(stream, b->uploaded, 0));  // foreign stream: order after the H2D copy
  LaunchBuffers lb = launch_buffers(b);
  if (b->runs_recorded >= 1024) b->runs_recorded = 0;  // keep the newest window
  while (b->events.size() < (size_t)(b->runs_recorded + 1) * EV_PER_RUN) {
    cudaEvent_t ev = nullptr;
    CU_TRY(cudaEventCreate(&ev));
    b->events.push_back(ev);
  }
  cudaEvent_t* ev = b->events.data() + (size_t)b->runs_recorded * EV_PER_RUN;
  CU_TRY(cudaMemsetAsync(b->d_counters, 0, sizeof(Counters), stream));
  CU_TRY(cudaEventRecord(ev[0], stream));
  lb.ev_bloom_done = ev[1];
  CU_TRY(launch_probe(d_ix_, h_ix_, b->bp, lb, sm_count_, stream));
  CU_TRY(cudaEventRecord(ev[2], stream));
  if (!launch_score_stage(b, lb, stream, ev[3], err)) return false;
  CU_TRY(cudaEventRecord(ev[4], stream));
  CU_TRY(launch_confusables(d_ix_, b->bp, lb, sm_count_, stream));
  CU_TRY(cudaEventRecord(ev[5], stream));
  CU_TRY(launch_finish(b->bp, lb, sm_count_, stream));
  CU_TRY(cudaEventRecord(ev[6], stream));
  if (!launch_export_chain(b, stream, err)) return false;
  CU_TRY(cudaEventRecord(ev[7], stream));
  if (!download_summary(b, stream, err)) return false;
  b->last_done = ev[7];
  ++b->runs_recorded;
  b->ran = true;
  b->settled = false;
  return true;
}

bool Engine::timings(DeviceBatch* b, float* stage_ms, std::string* err) {
  if (!b->ran) {
    *err = "batch has not been run";
    return false;
  }
  // averages over the runs since the previous call (CUDA events on the launching stream)
  CU_TRY(cudaEventSynchronize(b->last_done));
  double acc[EV_PER_RUN - 1] = {0};
  for (uint32_t r = 0; r < b->runs_recorded; ++r)
    for (int k = 0; k + 1 < EV_PER_RUN; ++k) {
      float ms = 0;
      CU_TRY(cudaEventElapsedTime(&ms, b->events[r * EV_PER_RUN + k], b->events[r * EV_PER_RUN + k + 1]));
      acc[k] += ms;
    }
  const uint32_t nr = std::max(1u, b->runs_recorded);
  for (int k = 0; k + 1 < EV_PER_RUN; ++k) stage_ms[k] = (float)(acc[k] / nr);
  b->runs_recorded = 0;
  return true;
}

bool Engine::counters(DeviceBatch* b, anl_counters* out, std::string* err) {
  if (!b->ran) {
    *err = "batch has not been run";
    return false;
  }
  Counters c;
  CU_TRY(cudaEventSynchronize(b->last_done));
  CU_TRY(cudaMemcpy(&c, b->d_counters, sizeof c, cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof *out);
  out->queries = b->n;
  out->deletion_keys = c.deletion_keys;
  out->probes = c.probes;
  out->filter_pass = c.filter_pass;
  out->probe_steps = c.table_steps;
  out->postings = c.postings;
  out->anagram_hits = c.anagram_hits;
  out->instance_pairs = c.instance_pairs;
  out->dl_pairs = c.dl_pairs;
  out->dl_cells = c.dl_cells;
  out->survivors = c.survivors;
  out->results = c.results;
  out->dp_pairs = c.dp_pairs;
  out->dp_cells = c.dp_cells;
  out->reruns = b->reruns;
  return true;
}

// ---- host post-pass ---------------------------------------------------------------------------------------
static inline double variant_score(const anl_variant& v, float fw) {  // src/types.rs:335-341
  if (fw == 0.0f) return v.dist_score;
  return (v.dist_score + ((double)fw * v.freq_score)) / (1.0 + (double)fw);
}
template <class It>
static void rank_variants(It first, It last, float fw) {  // src/lib.rs:1667 + src/types.rs:344-365
  std::stable_sort(first, last, [fw](const anl_variant& a, const anl_variant& b) {
    if (fw > 0.0f) return variant_score(a, fw) > variant_score(b, fw);
    if (a.dist_score != b.dist_score) return a.dist_score > b.dist_score;
    return a.freq_score > b.freq_score;
  });
}
// src/lib.rs:1536-1589; returns the number of results kept
static size_t crop_variants(const anl_variant* r, size_t n, size_t max_matches, float fw) {
  if (max_matches == 0 || n <= max_matches) return n;
  const double last = variant_score(r[max_matches - 1], fw), cropped = variant_score(r[max_matches], fw);
  if (cropped < last) return max_matches;
  size_t early = 0, late = 0;
  for (size_t i = 0; i < n; ++i) {
    if (r[i].dist_score == cropped && early == 0) early = i;
    if (r[i].dist_score < cropped) {
      late = i;
      break;
    }
  }
  if (early > 0) return early + 1;
  if (late > 0) return late + 1;
  return n;
}
// src/lib.rs:1598-1622
static size_t cutoff_variants(const anl_variant* r, size_t n, double cutoff_threshold, float fw) {
  if (!(cutoff_threshold >= 1.0) || n == 0) return n;
  const double best = variant_score(r[0], fw);
  for (size_t i = 1; i < n; ++i)
    if (variant_score(r[i], fw) <= best / cutoff_threshold) return i;
  return n;
}

// Host finish of a query of a model with variant lists (src/lib.rs:1504-1622 with expand_variants :1677-1727):
// the device delivered every candidate that passed the score threshold, in gather order, with absolute frequencies.
void Engine::finish_query_variants(const DeviceBatch& b, uint64_t qi, const OutRec* recs, uint32_t count, double max_freq,
                                   std::vector<anl_variant>* out) const {
  const size_t start = out->size();
  if (count == 0) return;
  const char* in = b.blob + b.offsets[qi];
  const size_t inlen = (size_t)(b.offsets[qi + 1] - b.offsets[qi]);
  const float fw = b.params.freq_weight;
  const bool early = !hm_->confusables.empty() && hm_->confusables_before_pruning;
  const bool late = !hm_->confusables.empty() && !hm_->confusables_before_pruning;
  bool expand = false;
  for (uint32_t i = 0; i < count; ++i) expand = expand || hm_->decoder[recs[i].vocab_id & ~OUT_SKIP_CONFUSABLES].has_variants;
  for (uint32_t i = 0; i < count; ++i) {
    const uint64_t id = recs[i].vocab_id & ~OUT_SKIP_CONFUSABLES;
    const bool skip = (recs[i].vocab_id & OUT_SKIP_CONFUSABLES) != 0;
    double dist = recs[i].dist_score;
    if (early && !skip) dist *= hm_->compute_confusable_weight(in, inlen, id);
    const double f = (double)recs[i].freq;
    const VocabEntry& item = hm_->decoder[id];
    if (expand) {
      for (const auto& vr : item.variant_of) {
        const double tf = (double)hm_->decoder[vr.first].frequency;  // the smaller of the two absolute frequencies
        out->push_back(anl_variant{vr.first, dist * vr.second, tf < f ? tf : f, id});
      }
      if (item.vocabtype & VT_TRANSPARENT) continue;
    }
    out->push_back(anl_variant{id, dist, f, ANL_NO_VIA});
  }
  size_t n = out->size() - start;
  anl_variant* v = out->data() + start;
  if (expand)
    for (size_t i = 0; i < n; ++i) max_freq = v[i].freq_score > max_freq ? v[i].freq_score : max_freq;
  if (max_freq > 0.0)
    for (size_t i = 0; i < n; ++i) v[i].freq_score = v[i].freq_score / max_freq;
  rank_variants(v, v + n, fw);
  if (expand) {  // Vec::dedup_by_key(|x| x.vocab_id): consecutive duplicates, the first one stays
    size_t w = 0;
    for (size_t i = 0; i < n; ++i)
      if (w == 0 || v[w - 1].vocab_id != v[i].vocab_id) v[w++] = v[i];
    n = w;
  }
  n = crop_variants(v, n, (size_t)b.params.max_matches, fw);
  if (late) {  // (every record, expanded ones against their target's text: the triage flags do not survive the expansion)
    for (size_t i = 0; i < n; ++i) v[i].dist_score *= hm_->compute_confusable_weight(in, inlen, v[i].vocab_id);
    rank_variants(v, v + n, fw);
  }
  n = cutoff_variants(v, n, b.params.cutoff_threshold, fw);
  out->resize(start + n);
}

void Engine::finish_query(const DeviceBatch& b, uint64_t qi, const OutRec* recs, uint32_t count, double max_freq,
                          std::vector<anl_variant>* out) const {
  if (hm_->any_variants) return finish_query_variants(b, qi, recs, count, max_freq, out);
  const size_t start = out->size();
  for (uint32_t i = 0; i < count; ++i) {
    // frequency normalisation (src/lib.rs:1521-1525): the same IEEE division the device ranked with
    const double f = (double)recs[i].freq;
    out->push_back(anl_variant{recs[i].vocab_id & ~OUT_SKIP_CONFUSABLES, recs[i].dist_score, max_freq > 0.0 ? f / max_freq : f,
                               ANL_NO_VIA});
  }
  if (b.bp.finish_mode == FINISH_FULL || count == 0) return;
  const char* in = b.blob + b.offsets[qi];
  const size_t inlen = (size_t)(b.offsets[qi + 1] - b.offsets[qi]);
  const float fw = b.params.freq_weight;
  anl_variant* v = out->data() + start;
  bool changed = false;
  for (uint32_t i = 0; i < count; ++i) {  // rescore_confusables, src/lib.rs:1656-1663
    if (recs[i].vocab_id & OUT_SKIP_CONFUSABLES) continue;  // the score kernel proved that no pattern can match
    const double w = hm_->compute_confusable_weight(in, inlen, v[i].vocab_id);
    if (w != 1.0) {
      v[i].dist_score *= w;
      changed = true;
    }
  }
  size_t n = count;
  // a stable sort of an already sorted list is the identity: only re-rank when a score changed (with the
  // device confusable stage, settled records of this query may already carry their weight)
  if (changed || b.dev_conf || b.bp.finish_mode == FINISH_GATHER) rank_variants(v, v + n, fw);
  if (b.bp.finish_mode == FINISH_GATHER) n = crop_variants(v, n, (size_t)b.params.max_matches, fw);
  n = cutoff_variants(v, n, b.params.cutoff_threshold, fw);
  out->resize(start + n);
}

// ---- settling a pass --------------------------------------------------------------------------------------------
// Hit-list overflows (a query with more than hit_cap candidate instances): those queries are run again with an exact
// capacity into buffers of their own, the results are patched into the batch's pool and the export stage runs again.
// Rare; the batch's stream simply carries the extra work.
bool Engine::rerun_overflowed(DeviceBatch* b, std::string* err, int* status) {
  *status = ANL_ERR_CUDA;
  const uint32_t n = b->n;
  cudaStream_t st = b->stream;
  CU_TRY(cudaMemcpyAsync(b->h_flags, b->d_qflags, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(b->h_hitcnt, b->d_hit_count, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  std::vector<uint32_t> which;
  uint32_t cap = b->bp.hit_cap;
  for (uint32_t i = 0; i < n; ++i)
    if ((b->h_flags[i] & (QF_HIT_OVERFLOW | QF_UNSUPPORTED)) == QF_HIT_OVERFLOW) {
      which.push_back(i);
      cap = std::max(cap, b->h_hitcnt[i]);
    }
  const uint32_t m = (uint32_t)which.size();
  if (m == 0) return true;
  if (profile_enabled()) fprintf(stderr, "[anl profile] %u queries overflowed hit_cap=%u\n", m, b->bp.hit_cap);
  cap = (cap + 31) & ~31u;
  BatchParams bp = b->bp;
  bp.hit_cap = cap;
  const uint64_t pool64 = (uint64_t)m * cap;  // results <= survivors <= hits: cannot overflow
  if (pool64 > 0xFFFFFF00ull) {
    *err = "hit overflow rerun too large";
    return false;
  }
  bp.pool_cap = (uint32_t)pool64;
  // grow-only buffers kept with the batch: steady state allocates (and frees) nothing, which matters because
  // cudaFree synchronises the device and would stall the other in-flight chunks
  const size_t scratch = score_scratch_bytes(bp, sm_count_, m);
  if (m > b->rr_cap_m) {
    if (!dev_realloc(&b->rr_qlist, m, err) || !dev_realloc(&b->rr_hit_count, m, err) || !dev_realloc(&b->rr_qflags, m, err) ||
        !dev_realloc(&b->rr_head, m, err))
      return false;
    b->rr_cap_m = m;
  }
  if ((size_t)m * cap > b->rr_cap_hits) {
    if (!dev_realloc(&b->rr_hits, (size_t)m * cap, err)) return false;
    b->rr_cap_hits = (size_t)m * cap;
  }
  const bool triage = b->has_qblob && h_ix_.conf_prefilter;
  if (bp.pool_cap > b->rr_cap_pool) {
    if (!dev_realloc(&b->rr_out, bp.pool_cap, err)) return false;
    if (b->dev_conf && !dev_realloc(&b->rr_conf_work, bp.pool_cap, err)) return false;
    if (triage && !dev_realloc(&b->rr_rec_query, bp.pool_cap, err)) return false;
    b->rr_cap_pool = bp.pool_cap;
  }
  if (b->dev_conf && !b->rr_conf_work && !dev_realloc(&b->rr_conf_work, b->rr_cap_pool, err)) return false;
  if (triage && !b->rr_rec_query && !dev_realloc(&b->rr_rec_query, b->rr_cap_pool, err)) return false;
  if (scratch > b->rr_cap_scratch) {
    if (!dev_realloc(&b->rr_scratch, scratch, err)) return false;
    b->rr_cap_scratch = scratch;
  }
  if (!b->rr_work && !dev_realloc(&b->rr_work, WORK_SLOTS, err)) return false;
  CU_TRY(cudaMemcpyAsync(b->rr_qlist, which.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  LaunchBuffers lb;
  lb.queries = b->d_rows;
  lb.qlist = b->rr_qlist;
  lb.qblob = b->has_qblob ? b->d_qblob : nullptr;
  lb.qboff = b->has_qblob ? b->d_qboff : nullptr;
  lb.rec_query = triage ? b->rr_rec_query : nullptr;
  lb.conf_work = b->dev_conf ? b->rr_conf_work : nullptr;
  lb.n = m;
  lb.hits = b->rr_hits;
  lb.hit_count = b->rr_hit_count;
  lb.qflags = b->rr_qflags;
  lb.out = b->rr_out;
  lb.out_head = b->rr_head;
  lb.scratch = b->rr_scratch;
  lb.scratch_bytes = b->rr_cap_scratch;
  lb.work = b->rr_work;
  lb.counters = nullptr;
  CU_TRY(cudaMemsetAsync(b->rr_work, 0, WORK_SLOTS * sizeof(unsigned int), st));
  CU_TRY(launch_probe(d_ix_, h_ix_, bp, lb, sm_count_, st));  // (no staged-node queue: the fused kernel)
  CU_TRY(launch_prefilter(d_ix_, bp, lb, sm_count_, st));         // (and the per-query score kernels: a handful of queries)
  CU_TRY(launch_score(d_ix_, h_ix_, bp, lb, sm_count_, st));
  CU_TRY(launch_confusables(d_ix_, bp, lb, sm_count_, st));
  CU_TRY(launch_finish(bp, lb, sm_count_, st));
  unsigned int rr_work[WORK_SLOTS];
  std::vector<uint32_t> fl(m);
  CU_TRY(cudaMemcpyAsync(rr_work, b->rr_work, sizeof rr_work, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(fl.data(), b->rr_qflags, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  for (uint32_t i = 0; i < m; ++i)
    if (fl[i] & (QF_HIT_OVERFLOW | QF_OUT_OVERFLOW | QF_UNSUPPORTED)) {
      *err = "internal error: overflow persisted after rerun";
      return false;
    }
  const uint32_t rr_total = rr_work[2], base = b->h_work[2];
  if ((uint64_t)base + rr_total > 0xFFFFFF00ull) {
    *err = "result pool too large; split the batch";
    *status = ANL_ERR_UNSUPPORTED;
    return false;
  }
  if (base + rr_total > b->cap_pool) {
    if (!grow_pool(b, base + rr_total + 1024, true, err)) return false;
    b->bp.pool_cap = b->cap_pool;
  }
  CU_TRY(launch_patch(m, b->rr_qlist, b->rr_head, b->rr_qflags, b->rr_out, b->d_head, b->d_qflags, b->d_out, base, b->cap_pool,
                      b->d_work + 2, rr_total, st));
  b->reruns += m;
  if (!launch_export_chain(b, st, err) || !download_summary(b, st, err)) return false;
  CU_TRY(cudaEventSynchronize(b->ev_done));
  if (b->h_summary->n_rerun != 0) {
    *err = "internal error: hit overflow persisted after the patch";
    return false;
  }
  *status = ANL_OK;
  return true;
}

// Pool and staged-node queue overflows of a pass: run the affected stages again with larger buffers.
static const char* kPoolPersisted = "internal error: result pool overflow persisted";
bool Engine::settle(DeviceBatch* b, std::string* err, int* status) {
  *status = ANL_ERR_CUDA;
  if (!b->ran) {
    *err = "batch has not been run";
    *status = ANL_ERR_INVALID;
    return false;
  }
  if (b->settled) {
    *status = ANL_OK;
    return true;
  }
  CU_TRY(cudaSetDevice(device_));
  const uint32_t n = b->n;
  cudaStream_t st = b->stream;
  for (int attempt = 0;; ++attempt) {
    CU_TRY(cudaEventSynchronize(b->ev_done));
    if (b->split && n && b->h_work[4] > std::min<size_t>(b->cap_queue, 0x7FFFFFF0u)) {
      // the staged-node queue was too small for this batch: run it again with the fused probe kernel
      if (profile_enabled()) fprintf(stderr, "[anl profile] staged-node queue overflow (%u > %zu): fused rerun\n", b->h_work[4], b->cap_queue);
      b->split = false;
      CU_TRY(cudaStreamWaitEvent(st, b->ev_done, 0));
      LaunchBuffers lbq = launch_buffers(b);
      lbq.counters = nullptr;
      CU_TRY(launch_probe(d_ix_, h_ix_, b->bp, lbq, sm_count_, st));
      if (!relaunch_from_score(b, err)) return false;
      b->reruns += 1;
      --attempt;
      continue;
    }
    if (b->use_pairs && n && b->h_work[WORK_PAIR_TOTAL] > b->cap_pairs) {
      // more pairs survived the filter than the pair list holds: the counter is the exact requirement
      if (profile_enabled()) fprintf(stderr, "[anl profile] pair list overflow (%u > %zu): score stage again\n", b->h_work[WORK_PAIR_TOTAL], b->cap_pairs);
      if (!grow_pairs(b, (size_t)b->h_work[WORK_PAIR_TOTAL] + 4096, err) || !relaunch_from_score(b, err)) return false;
      b->reruns += 1;
      --attempt;
      continue;
    }
    const unsigned int used = n ? b->h_work[2] : 0;
    if (used <= b->bp.pool_cap || b->merged) break;
    if (attempt >= 2) {
      *err = kPoolPersisted;
      return false;
    }
    // the cursor counted every query's results, so `used` is the exact requirement
    if (!grow_pool(b, (uint32_t)std::min<uint64_t>(0xFFFFFF00ull, (uint64_t)used + 1024), false, err)) return false;
    b->bp.pool_cap = b->cap_pool;
    if (!relaunch_from_score(b, err)) return false;
    b->reruns += 1;
  }
  if (b->bp.finish_mode != FINISH_SHARD && b->h_summary->n_rerun && !rerun_overflowed(b, err, status)) return false;
  b->settled = true;
  *status = ANL_OK;
  return true;
}

// Does the host have to touch this batch's records?  Yes when a confusable / variant-list post-pass follows that the
// device did not run, or when the device flagged queries it could not finish.
bool Engine::needs_host_finish(const DeviceBatch* b) const {
  if (b->bp.finish_mode == FINISH_FULL) return false;
  if (hm_->any_variants || !b->dev_conf) return true;
  return b->h_summary->n_host_finish != 0;
}

bool Engine::issue_download(DeviceBatch* b, ResultSet* out, uint64_t qbase, uint64_t vbase, std::string* err) {
  CU_TRY(cudaSetDevice(device_));
  const uint32_t n = b->n, total = b->h_summary->total;
  cudaStream_t st = b->stream;
  if (n) {
    CU_TRY(launch_offsets(n, b->d_loff, vbase, b->d_off64, st));
    CU_TRY(cudaMemcpyAsync(out->offsets.data() + qbase, b->d_off64, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(out->flags.data() + qbase, b->d_oflags, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  if (total)
    CU_TRY(cudaMemcpyAsync(out->variants.data() + vbase, b->d_final, (size_t)total * sizeof(anl_variant), cudaMemcpyDeviceToHost, st));
  b->results = total;
  return true;
}

bool Engine::wait_download(DeviceBatch* b, std::string* err) {
  CU_TRY(cudaStreamSynchronize(b->stream));
  return true;
}

// Slow path of a batch: the 16-byte records come to the host, the queries the device could not finish go through
// finish_query (confusable rescoring, variant expansion, re-rank, crop, cut-off: src/lib.rs:1504-1622), everything
// is assembled into the batch's part of `out`.  The caller guarantees that no download into `out` is in flight
// (its variants array may move).
bool Engine::finish_on_host(DeviceBatch* b, ResultSet* out, uint64_t qbase, uint64_t vbase, uint64_t* written, std::string* err,
                            int* status) {
  *status = ANL_ERR_CUDA;
  CU_TRY(cudaSetDevice(device_));
  const uint32_t n = b->n;
  cudaStream_t st = b->stream;
  PhaseTimer pt;
  const unsigned int used = n ? std::min<unsigned int>(b->h_work[2], b->cap_pool) : 0;
  if (n) {
    CU_TRY(cudaMemcpyAsync(b->h_head, b->d_head, (size_t)n * sizeof(OutHead), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(b->h_flags, b->d_qflags, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(b->h_enc_status, b->d_enc_status, (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  if (used) CU_TRY(cudaMemcpyAsync(b->h_out, b->d_out, (size_t)used * sizeof(OutRec), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  pt.lap("host finish: D2H");
  const bool post_pass = b->bp.finish_mode != FINISH_FULL;
  auto locate = [&](uint32_t i, const OutRec** recs, uint32_t* count, double* maxf, bool* host) {
    const OutHead& h = b->h_head[i];
    const bool none = (b->h_flags[i] & (QF_EMPTY | QF_HIT_OVERFLOW | QF_UNSUPPORTED | QF_OUT_OVERFLOW)) != 0;
    *recs = b->h_out + h.offset;
    *count = none ? 0 : (h.count & ~HEAD_HOST_FINISH);
    *maxf = h.max_freq;
    *host = post_pass && (hm_->any_variants || !b->dev_conf || (h.count & HEAD_HOST_FINISH));
  };
  uint64_t* offs = out->offsets.data() + qbase;
  uint32_t* flags = out->flags.data() + qbase;
  // pass 1: queries the host must finish go through finish_query into per-thread side buffers; the
  // counts of all other queries are final as they come from the device
  const unsigned maxt = host_threads();
  std::vector<std::vector<anl_variant>> part(post_pass ? maxt : 0);
  std::vector<uint32_t> counts(n, 0), side_pos(post_pass ? n : 0, 0);
  uint64_t host_queries = 0;
  std::vector<uint64_t> hq(maxt, 0);
  parallel_ranges(n, 512, [&](unsigned t, uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) {
      const OutRec* r;
      uint32_t c;
      double mf;
      bool host;
      locate((uint32_t)i, &r, &c, &mf, &host);
      uint32_t of = 0;
      if ((b->h_flags[i] & QF_EMPTY) && b->h_enc_status[i] == ENC_OK) of |= ANL_QUERY_EMPTY;
      if ((b->h_flags[i] & QF_UNSUPPORTED) || b->h_enc_status[i] == ENC_TOO_LONG_UNSUPPORTED) of |= ANL_QUERY_UNSUPPORTED;
      flags[i] = of;
      if (!host) {
        counts[i] = c;
        continue;
      }
      std::vector<anl_variant>& buf = part[t];
      const size_t before = buf.size();
      finish_query(*b, i, r, c, mf, &buf);
      side_pos[i] = (uint32_t)before;
      counts[i] = (uint32_t)(buf.size() - before);
      ++hq[t];
    }
  });
  for (uint64_t v : hq) host_queries += v;
  uint64_t tot = vbase;
  for (uint32_t i = 0; i < n; ++i) {
    offs[i] = tot;
    tot += counts[i];
  }
  if (tot > out->variants.capacity()) out->variants.reserve(std::max<size_t>(tot, out->variants.capacity() + out->variants.capacity() / 2));
  // pass 2: convert / copy straight into place (same thread ranges as pass 1)
  anl_variant* vout = out->variants.data();
  parallel_ranges(n, 512, [&](unsigned t, uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) {
      const OutRec* r;
      uint32_t c;
      double mf;
      bool host;
      locate((uint32_t)i, &r, &c, &mf, &host);
      anl_variant* dst = vout + offs[i];
      if (host) {
        if (counts[i]) memcpy(dst, part[t].data() + side_pos[i], (size_t)counts[i] * sizeof(anl_variant));
        continue;
      }
      for (uint32_t k = 0; k < c; ++k) {
        // frequency normalisation (src/lib.rs:1521-1525): the same IEEE division the device ranked with
        const double f = (double)r[k].freq;
        dst[k] = anl_variant{r[k].vocab_id & ~OUT_SKIP_CONFUSABLES, r[k].dist_score, mf > 0.0 ? f / mf : f, ANL_NO_VIA};
      }
    }
  });
  b->results = tot - vbase;
  *written = tot - vbase;
  pt.lap("host finish: post-pass+assemble");
  if (profile_enabled()) {
    uint64_t v[4];
    confusable_stats(v);
    fprintf(stderr, "[anl profile] %llu of %u queries finished on the host (device confusables: %s)\n",
            (unsigned long long)host_queries, n, b->dev_conf ? "on" : "off");
    fprintf(stderr, "[anl profile] confusable checks %llu, prefilter pass %llu (ascii), single-edit fast %llu, full script %llu\n",
            (unsigned long long)v[0], (unsigned long long)v[1], (unsigned long long)v[2], (unsigned long long)v[3]);
  }
  *status = ANL_OK;
  return true;
}

// One batch -> a fresh result set (device-batch API, sharded merge).
bool Engine::fetch_batch(DeviceBatch* b, ResultSet* out, std::string* err, int* status) {
  if (b->sharded && !b->merged) {
    *err = "this model holds a lexicon shard: use the shard export / merge calls";
    *status = ANL_ERR_INVALID;
    return false;
  }
  if (!settle(b, err, status)) return false;
  *status = ANL_ERR_CUDA;
  const uint32_t n = b->n;
  out->offsets.resize((size_t)n + 1);
  out->flags.resize(std::max<size_t>(n, 1));
  out->flags.resize(n);
  uint64_t total = 0;
  if (needs_host_finish(b)) {
    out->variants.reserve(std::max<size_t>(b->h_summary->total, 16));
    if (!finish_on_host(b, out, 0, 0, &total, err, status)) return false;
  } else {
    total = b->h_summary->total;
    out->variants.reserve(std::max<size_t>(total, 16));
    if (!issue_download(b, out, 0, 0, err) || !wait_download(b, err)) return false;
  }
  out->offsets[n] = total;
  out->variants.resize(total);
  *status = ANL_OK;
  return true;
}

// ---- lexicon-sharded mode ---------------------------------------------------------------------------------------
bool Engine::shard_export_size(DeviceBatch* b, uint64_t* n_records, uint32_t* max_per_query, std::string* err, int* status) {
  *status = ANL_ERR_CUDA;
  if (!b->ran || !b->sharded) {
    *err = "not a sharded batch that has been run";
    *status = ANL_ERR_INVALID;
    return false;
  }
  if (!settle(b, err, status)) return false;  // (pool / queue overflows; a shard's hit overflow is reported at the merge)
  *status = ANL_ERR_CUDA;
  const uint32_t n = b->n;
  if (n) {
    CU_TRY(cudaMemcpyAsync(b->h_head, b->d_head, (size_t)n * sizeof(OutHead), cudaMemcpyDeviceToHost, b->stream));
    CU_TRY(cudaStreamSynchronize(b->stream));
  }
  uint32_t mx = 0;
  for (uint32_t i = 0; i < n; ++i) mx = std::max(mx, b->h_head[i].count);
  *n_records = n ? b->h_work[2] : 0;
  if (max_per_query) *max_per_query = mx;
  *status = ANL_OK;
  return true;
}

bool Engine::shard_export(DeviceBatch* b, void* d_heads, void* d_records, void* d_gids, void* d_flags, std::string* err) {
  CU_TRY(cudaSetDevice(device_));
  const uint32_t n = b->n;
  const unsigned int total = n ? b->h_work[2] : 0;
  cudaStream_t st = b->stream;
  if (n) {
    CU_TRY(cudaMemcpyAsync(d_heads, b->d_head, (size_t)n * sizeof(OutHead), cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_flags, b->d_qflags, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
  if (total) {
    CU_TRY(cudaMemcpyAsync(d_records, b->d_out, (size_t)total * sizeof(OutRec), cudaMemcpyDeviceToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_gids, b->d_gid, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
  CU_TRY(cudaStreamSynchronize(st));
  return true;
}

bool Engine::shard_merge(DeviceBatch* b, uint32_t n_shards, const void* d_heads_all, const void* d_records_all,
                         const void* d_gids_all, const void* d_flags_all, uint64_t record_stride, uint32_t max_survivors,
                         ResultSet* out, std::string* err, int* status) {
  return shard_merge_strided(b, n_shards, d_heads_all, d_records_all, d_gids_all, d_flags_all, b->n, record_stride, max_survivors, out,
                             err, status);
}

// heads / flags of shard r at r * query_stride, records / gather ids at r * record_stride
bool Engine::shard_merge_strided(DeviceBatch* b, uint32_t n_shards, const void* d_heads_all, const void* d_records_all,
                                 const void* d_gids_all, const void* d_flags_all, uint64_t query_stride, uint64_t record_stride,
                                 uint32_t max_survivors, ResultSet* out, std::string* err, int* status) {
  *status = ANL_ERR_CUDA;
  if (!b->sharded) {
    *err = "not a sharded batch";
    *status = ANL_ERR_INVALID;
    return false;
  }
  CU_TRY(cudaSetDevice(device_));
  const uint32_t n = b->n;
  cudaStream_t st = b->stream;
  // a hit-list overflow on any shard cannot be repaired after the exchange: fail on every rank alike.  Checked on the
  // device (8 bytes come back): downloading the n_shards x n flags cost more than the merge itself at 8 shards.
  {
    unsigned int* res = b->d_work + 20;  // (two free work slots; the launchers of the merge / export stage do not touch them)
    CU_TRY(launch_shard_flagcheck(reinterpret_cast<const uint32_t*>(d_flags_all), n, n_shards, query_stride, res, sm_count_, st));
    CU_TRY(cudaMemcpyAsync(b->h_work + 20, res, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (b->h_work[20]) {
      *err = "query " + std::to_string(b->h_work[21]) + " exceeds the per-query candidate capacity on a shard; raise ANL_HIT_CAP";
      *status = ANL_ERR_UNSUPPORTED;
      return false;
    }
  }
  const uint64_t pool64 = std::max<uint64_t>(1024, record_stride * n_shards);
  if (pool64 > 0xFFFFFF00ull) {
    *err = "sharded result pool too large; split the batch";
    *status = ANL_ERR_UNSUPPORTED;
    return false;
  }
  if (!grow_pool(b, (uint32_t)pool64, false, err)) return false;
  const uint32_t cap = std::max<uint32_t>(32, (max_survivors + 31) & ~31u);
  const size_t scratch = merge_scratch_bytes(sm_count_, n, cap);
  if (scratch > b->cap_scratch || !b->d_scratch) {
    if (!dev_realloc(reinterpret_cast<uint8_t**>(&b->d_scratch), scratch, err)) return false;
    b->cap_scratch = scratch;
  }
  BatchParams bp = b->bp;
  bp.finish_mode = b->final_mode;
  bp.pool_cap = b->cap_pool;
  CU_TRY(launch_merge(bp, n, n_shards, reinterpret_cast<const OutHead*>(d_heads_all), (uint32_t)query_stride,
                      reinterpret_cast<const OutRec*>(d_records_all), reinterpret_cast<const uint32_t*>(d_gids_all), (uint32_t)record_stride,
                      reinterpret_cast<const uint32_t*>(d_flags_all), b->d_qflags, b->d_out, b->d_head, b->d_scratch, cap,
                      b->d_work, sm_count_, st));
  // the merged lists are final: the host post-pass / export stage follow the final mode
  b->merged = true;
  b->bp.finish_mode = b->final_mode;
  b->bp.pool_cap = b->cap_pool;
  b->settled = false;
  bool ok = launch_export_chain(b, st, err) && download_summary(b, st, err);
  if (ok) {
    if (out) {
      ok = fetch_batch(b, out, err, status);
    } else {
      ok = settle(b, err, status);  // device-only merge: ranked lists (and their exported form) stay in the batch's buffers
    }
  }
  b->bp.finish_mode = FINISH_SHARD;
  b->merged = false;
  b->settled = false;
  if (ok) *status = ANL_OK;
  return ok;
}

// ---- the batch call: chunks pipelined over one or more devices -------------------------------------------------
namespace {
struct CallShared {
  ResultSet* out = nullptr;
  uint64_t n_total = 0;
  std::mutex m;
  std::condition_variable cv;
  uint64_t next_chunk = 0;  // the chunk whose turn it is to take its place in `out`
  uint64_t vbase = 0;       // variants placed so far
  bool failed = false;
  std::string err;
  int status = ANL_OK;
  std::vector<cudaStream_t> streams;  // every stream that has copied into `out`: synchronised before `out` moves
  // ANL_TIMELINE=1: per chunk, the device times of its stages (CUDA events against `ref`) and the host's own steps
  bool timeline = false;
  bool stagger = false;   // ANL_STAGGER=1: staggered launches (below); off by default -- it gains 2-3 % on cfg 2 and costs 14 % on cfg 5
  bool stagger_always = false;
  int stagger_event = 1;  // which stage of the predecessor must be over (1 = Bloom stage, 2 = exact stage, ...)
  cudaEvent_t ref = nullptr;
  std::chrono::steady_clock::time_point t0;
};
struct InFlight {
  uint64_t chunk, q0;
  DeviceBatch* b;
  double t_launched = 0, t_placed = 0;  // host clock, ms since the call began (timeline only)
};
double host_ms(const CallShared& S) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - S.t0).count();
}

void fail_call(CallShared& S, const std::string& e, int status) {  // (S.m held)
  if (!S.failed) {
    S.failed = true;
    S.err = e;
    S.status = status ? status : ANL_ERR_CUDA;
  }
}

// room for `need` variants in S.out; `hint` = what the whole call is expected to need (S.m held)
bool reserve_variants(CallShared& S, uint64_t need, uint64_t hint) {
  if (need <= S.out->variants.capacity()) return true;
  for (cudaStream_t st : S.streams)
    if (cudaStreamSynchronize(st) != cudaSuccess) return false;
  S.out->variants.resize(S.vbase);  // (what reserve has to preserve)
  S.out->variants.reserve(std::max<uint64_t>(need, hint));
  return true;
}

// One device's share of the call: chunks d, d + D, d + 2D, ... with up to DEPTH of them in flight, each on its own
// stream, so the kernel tails of one chunk overlap the next chunk's kernels and the downloads overlap both.
// Every chunk index takes its turn exactly once -- also after a failure or an exception -- so no other device's
// thread waits for a turn that never comes.
void device_loop_body(Engine* e, unsigned d, unsigned D, CallShared& S, const char* blob, const uint64_t* offsets, uint64_t n,
                      const anl_search_params& p, uint64_t CHUNK, size_t DEPTH, uint64_t nchunks, uint64_t* next_turn,
                      std::deque<InFlight>& running, std::deque<InFlight>& copying) {
  auto finish_copy = [&]() {
    InFlight f = copying.front();
    copying.pop_front();
    std::string e2;
    if (!e->wait_download(f.b, &e2)) {
      std::lock_guard<std::mutex> lk(S.m);
      fail_call(S, e2, ANL_ERR_CUDA);
    }
    if (S.timeline && f.b->runs_recorded > 0) {
      float t[EV_PER_RUN];
      for (int k = 0; k < EV_PER_RUN; ++k) {  // (events of another device cannot be compared with `ref`: relative to the chunk's own start)
        cudaEvent_t from = d == 0 ? S.ref : f.b->events[(size_t)(f.b->runs_recorded - 1) * EV_PER_RUN];
        if (cudaEventElapsedTime(&t[k], from, f.b->events[(size_t)(f.b->runs_recorded - 1) * EV_PER_RUN + k]) != cudaSuccess) t[k] = -1;
      }
      fprintf(stderr,
              "[anl timeline] dev %u chunk %3llu n %6u | device%s: start %7.2f bloom %7.2f exact %7.2f pairs %7.2f score %7.2f conf %7.2f "
              "finish %7.2f export %7.2f | host: launched %7.2f placed %7.2f downloaded %7.2f\n",
              d, (unsigned long long)f.chunk, f.b->n, d == 0 ? "" : " (since its start)", t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7],
              f.t_launched, f.t_placed, host_ms(S));
    }
    e->free_batch(f.b);
  };
  // settle the chunk, then -- when it is its turn -- give it its place in `out` and start the download
  auto place = [&](InFlight f) {
    std::string e2;
    int s2 = ANL_OK;
    bool ok = f.b != nullptr;
    if (ok) {
      bool skip;
      {
        std::lock_guard<std::mutex> lk(S.m);
        skip = S.failed;
      }
      ok = !skip && e->settle(f.b, &e2, &s2);
    }
    std::unique_lock<std::mutex> lk(S.m);
    S.cv.wait(lk, [&]() { return S.next_chunk == f.chunk; });
    struct Advance {  // the turn passes on however this scope is left
      CallShared& S;
      std::unique_lock<std::mutex>& lk;
      uint64_t* next_turn;
      unsigned D;
      ~Advance() {
        if (!lk.owns_lock()) lk.lock();
        ++S.next_chunk;
        *next_turn += D;
        lk.unlock();
        S.cv.notify_all();
      }
    } advance{S, lk, next_turn, D};
    if (ok && !S.failed) {
      const uint32_t total = e->export_total(f.b);
      // size the variants array from the first chunk's density (+ slack): later chunks then rarely move it
      const uint64_t hint = f.b->n ? (uint64_t)((double)std::max<uint32_t>(total, 1) / f.b->n * (double)S.n_total * 1.15) + 65536 : 0;
      if (e->needs_host_finish(f.b)) {
        for (cudaStream_t st : S.streams) cudaStreamSynchronize(st);  // its variants may move
        uint64_t written = 0;
        ok = reserve_variants(S, S.vbase + total, hint) && e->finish_on_host(f.b, S.out, f.q0, S.vbase, &written, &e2, &s2);
        if (ok) S.vbase += written;
        S.out->variants.resize(S.vbase);
      } else {
        ok = reserve_variants(S, S.vbase + total, hint) && e->issue_download(f.b, S.out, f.q0, S.vbase, &e2);
        if (ok) {
          if (std::find(S.streams.begin(), S.streams.end(), f.b->stream) == S.streams.end()) S.streams.push_back(f.b->stream);
          S.vbase += total;
        }
      }
      if (!ok && e2.empty()) e2 = "out of memory for the result arrays";
    }
    if (!ok && f.b && !e2.empty()) fail_call(S, e2, s2);
  };
  auto place_front = [&]() {
    InFlight f = running.front();
    running.pop_front();
    place(f);
    f.t_placed = S.timeline ? host_ms(S) : 0;
    if (f.b) copying.push_back(f);  // (owned by `copying` from here on, whatever happens in place())
    while (copying.size() > 1) finish_copy();
  };
  size_t launched = 0;
  for (uint64_t c = d; c < nchunks; c += D, ++launched) {
    const uint64_t lo = c * CHUNK, m = std::min(CHUNK, n - lo);
    bool skip;
    {
      std::lock_guard<std::mutex> lk(S.m);
      skip = S.failed;
    }
    DeviceBatch* b = nullptr;
    if (!skip) {
      std::string e2;
      int s2 = ANL_OK;
      // Staggered starts (ANL_STAGGER=1, an experiment kept as a knob): chunks launched back to back share the GPU evenly
      // and therefore all END together -- the device then drains before the next wave can start (ANL_TIMELINE: two waves
      // of four chunks).  With the knob a chunk is launched when its predecessor has left the Bloom stage, so the chunks
      // in flight stay apart.  Measured (tools/e2e_matrix.py): cfg 2 +2-3 %, eng3 -2 %, cfg 4 +-0, cfg 5 -14 % -- the
      // latency-bound probe kernels of an HBM-resident index WANT several chunks in the same stage at once (more loads
      // in flight); so the default stays back-to-back.
      // Only while the pipeline fills (ANL_STAGGER_ALWAYS=1: before every launch): afterwards a launch follows a
      // completion, which keeps the spacing by itself, and waiting for the predecessor's Bloom stage would hold launches
      // back where that stage is half of a chunk's work (cfg 4, cfg 5).
      if (S.stagger && (S.stagger_always || launched < DEPTH) && !running.empty() && running.back().b &&
          running.back().b->runs_recorded > 0) {
        DeviceBatch* prev = running.back().b;
        cudaEventSynchronize(prev->events[(size_t)(prev->runs_recorded - 1) * EV_PER_RUN + S.stagger_event]);
      }
      b = e->create_batch(blob, offsets + lo, m, p, false, false, &e2, &s2);
      if (b && !e->run_batch(b, nullptr, &e2)) {
        s2 = ANL_ERR_CUDA;
        cudaStreamSynchronize(b->stream);
        e->free_batch(b);
        b = nullptr;
      }
      if (!b) {
        std::lock_guard<std::mutex> lk(S.m);
        fail_call(S, e2, s2);
      }
    }
    running.push_back(InFlight{c, lo, b, S.timeline ? host_ms(S) : 0, 0});
    while (running.size() >= DEPTH) place_front();
  }
  while (!running.empty()) place_front();
  while (!copying.empty()) finish_copy();
}

void device_loop(Engine* e, unsigned d, unsigned D, CallShared& S, const char* blob, const uint64_t* offsets, uint64_t n,
                 const anl_search_params& p, uint64_t CHUNK, size_t DEPTH) {
  const uint64_t nchunks = n == 0 ? 1 : (n + CHUNK - 1) / CHUNK;
  uint64_t next_turn = d;  // the next chunk of this device that has not taken its turn yet
  std::deque<InFlight> running, copying;
  try {
    device_loop_body(e, d, D, S, blob, offsets, n, p, CHUNK, DEPTH, nchunks, &next_turn, running, copying);
  } catch (const std::exception& ex) {
    std::lock_guard<std::mutex> lk(S.m);
    fail_call(S, std::string("internal error: ") + ex.what(), ANL_ERR_INVALID);
  } catch (...) {
    std::lock_guard<std::mutex> lk(S.m);
    fail_call(S, "internal error: unknown exception", ANL_ERR_INVALID);
  }
  // after an exception: pass on the turns this device still owes, release what is in flight
  for (uint64_t c = next_turn; c < nchunks; c += D) {
    std::unique_lock<std::mutex> lk(S.m);
    S.cv.wait(lk, [&]() { return S.next_chunk == c; });
    ++S.next_chunk;
    lk.unlock();
    S.cv.notify_all();
  }
  for (std::deque<InFlight>* q : {&running, &copying})
    for (InFlight& f : *q)
      if (f.b) {
        cudaStreamSynchronize(f.b->stream);
        e->free_batch(f.b);
      }
}
}  // namespace

bool find_variants_batch_multi(const std::vector<Engine*>& engines, const char* blob, const uint64_t* offsets, uint64_t n,
                               const anl_search_params& p, ResultSet* out, std::string* err, int* status) {
  const unsigned D = (unsigned)engines.size();
  if (D == 0) {
    *err = "model has not been built";
    *status = ANL_ERR_NOT_BUILT;
    return false;
  }
  // Chunks of 65536 queries: large enough for the persistent grids to fill the device, small enough that several are
  // in flight per device and the first results come back early.
  // (measured on cfg 2, 1 M queries, profiles/r02e_sweep.txt: 65536 -> 32.5, 131072 -> 35.4, 262144 -> 34.5 M q/s:
  // larger chunks have fewer kernel tails, smaller ones a shorter ramp; calls that span few chunks stay at 65536)
  uint64_t CHUNK = n >= ((uint64_t)D << 19) ? (1u << 17) : (1u << 16);
  if (const char* e = getenv("ANL_CHUNK")) CHUNK = (uint64_t)std::max(1024, atoi(e));
  size_t DEPTH = 6;  // (free_batch keeps 8 batches cached: no device allocation after the first chunks; measured 3: 36.9, 4: 36.9, 5: 37.6, 6: 37.9 M q/s)
  if (const char* e = getenv("ANL_INFLIGHT")) DEPTH = (size_t)std::min(6, std::max(1, atoi(e)));
  CallShared S;
  S.out = out;
  S.n_total = n;
  S.t0 = std::chrono::steady_clock::now();
  if (const char* e = getenv("ANL_TIMELINE")) S.timeline = atoi(e) != 0;
  if (const char* e = getenv("ANL_STAGGER_ALWAYS")) S.stagger_always = atoi(e) != 0;
  if (const char* e = getenv("ANL_STAGGER")) {
    S.stagger = atoi(e) != 0;
    if (atoi(e) >= 1 && atoi(e) < EV_PER_RUN) S.stagger_event = atoi(e);
  }
  if (S.timeline) {
    cudaSetDevice(engines[0]->device());
    if (cudaEventCreate(&S.ref) != cudaSuccess || cudaEventRecord(S.ref, 0) != cudaSuccess) S.timeline = false;
  }
  out->offsets.resize((size_t)n + 1);
  out->flags.resize(std::max<uint64_t>(n, 1));
  out->flags.resize(n);
  out->variants.clear();
  const uint64_t nchunks = n == 0 ? 1 : (n + CHUNK - 1) / CHUNK;
  const unsigned used = (unsigned)std::min<uint64_t>(D, nchunks);
  if (used <= 1) {
    device_loop(engines[0], 0, 1, S, blob, offsets, n, p, CHUNK, DEPTH);
  } else {
    // one dispatcher thread per device (their host phases run inline: the shared pool serves one job at a time)
    std::vector<std::thread> th;
    for (unsigned d = 1; d < used; ++d)
      th.emplace_back([&, d]() {
        serial_ranges_flag() = true;
        device_loop(engines[d], d, used, S, blob, offsets, n, p, CHUNK, DEPTH);
      });
    const bool was = serial_ranges_flag();
    serial_ranges_flag() = true;
    device_loop(engines[0], 0, used, S, blob, offsets, n, p, CHUNK, DEPTH);
    serial_ranges_flag() = was;
    for (auto& t : th) t.join();
  }
  if (S.timeline) {
    fprintf(stderr, "[anl timeline] call of %llu queries done at %.2f ms (host clock)\n", (unsigned long long)n, host_ms(S));
    if (S.ref) cudaEventDestroy(S.ref);
  }
  if (S.failed) {
    *err = S.err;
    *status = S.status;
    return false;
  }
  out->offsets[n] = S.vbase;
  out->variants.resize(S.vbase);
  *status = ANL_OK;
  return true;
}

}  // namespace anl
