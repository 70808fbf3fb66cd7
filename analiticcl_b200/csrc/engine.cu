// engine.cu -- device residency of the index and the batched lookup pipeline.
//
// Replaces the reference's per-query call chain find_variants -> find_nearest_anahashes ->
// gather_instances -> score_and_rank (src/lib.rs:972-1027) with: host normalisation of a whole
// batch, one H2D copy, probe kernel, score/rank kernel, one D2H copy, and a thin host post-pass
// (late confusable rescoring + cut-off, src/lib.rs:1592-1622) that only exists when confusables are
// loaded.  There is no CPU fallback: every failure of the CUDA path is returned as an error.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "unicode_tables.h"

namespace anl {

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
bool Engine::dev_alloc(T** p, size_t count, std::string* err) {
  void* q = nullptr;
  CU_TRY(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  *p = reinterpret_cast<T*>(q);
  return true;
}

Engine::~Engine() {
  release_index();
  if (stream_) cudaStreamDestroy(stream_);
}

void Engine::release_index() {
  for (void* p : index_allocs_) cudaFree(p);
  index_allocs_.clear();
  if (d_mset_) cudaFree(d_mset_);
  d_mset_ = nullptr;
  if (d_ix_) cudaFree(d_ix_);
  d_ix_ = nullptr;
}

template <class T>
static bool upload_vec(const std::vector<T>& v, const T** out, std::vector<void*>* allocs, std::string* err) {
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
  if (!upload_vec(hx.ana_key, &ix.ana_key, &index_allocs_, err)) return false;
  if (!upload_vec(hx.ana_inst_off, &ix.ana_inst_off, &index_allocs_, err)) return false;
  if (!upload_vec(hx.inst_rows, &ix.inst_rows, &index_allocs_, err)) return false;
  ix.norm_stride = hx.norm_stride;
  if (!upload_vec(hx.inst_vocab, &ix.inst_vocab, &index_allocs_, err)) return false;
  if (!upload_vec(hx.inst_freq, &ix.inst_freq, &index_allocs_, err)) return false;
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
  memcpy(ix.prime_of, hx.prime_of, sizeof ix.prime_of);
  memcpy(ix.charcount_mask, hx.charcount_mask, sizeof ix.charcount_mask);
  ix.max_charcount = hx.max_charcount;
  ix.max_len = hx.max_len;
  ix.n_anagrams = (uint32_t)hx.ana_key.size();
  ix.n_instances = (uint32_t)hx.inst_vocab.size();
  ix.have_freq = hm_->have_freq ? 1 : 0;
  ix.mset = nullptr;
  h_ix_ = ix;
  if (!dev_alloc(&d_ix_, 1, err)) return false;
  CU_TRY(cudaMemcpy(d_ix_, &h_ix_, sizeof h_ix_, cudaMemcpyHostToDevice));
  hm_->index.mset_built_j = 0;
  return true;
}

bool Engine::ensure_msets(uint32_t J, std::string* err) {
  if (J == 0 || (hm_->index.mset_built_j >= J && d_mset_)) return true;
  if (!hm_->ensure_msets(J, err)) return false;
  const HostIndex& hx = hm_->index;
  CU_TRY(cudaStreamSynchronize(stream_));
  if (d_mset_) cudaFree(d_mset_);
  d_mset_ = nullptr;
  CU_TRY(cudaMalloc(&d_mset_, std::max<size_t>(hx.mset.size(), 1) * sizeof(MsetEntry)));
  CU_TRY(cudaMemcpy(d_mset_, hx.mset.data(), hx.mset.size() * sizeof(MsetEntry), cudaMemcpyHostToDevice));
  h_ix_.mset = reinterpret_cast<const MsetEntry*>(d_mset_);
  memcpy(h_ix_.mset_end, hx.mset_end, sizeof h_ix_.mset_end);
  CU_TRY(cudaMemcpy(d_ix_, &h_ix_, sizeof h_ix_, cudaMemcpyHostToDevice));
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
  uint32_t hit_cap = 1024;
  if (const char* e = getenv("ANL_HIT_CAP")) hit_cap = std::max(1, atoi(e));
  bp->hit_cap = hit_cap;
  uint32_t out_cap;
  if (bp->finish_mode == FINISH_GATHER)
    out_cap = 256;
  else if (bp->max_matches > 0)
    out_cap = bp->max_matches + 1;
  else
    out_cap = 64;
  if (const char* e = getenv("ANL_OUT_CAP")) out_cap = std::max(1, atoi(e));
  bp->out_cap = std::min(out_cap, hit_cap);
  const uint32_t kcap = std::min<uint32_t>(threshold_cap(p.max_anagram_distance), ANL_MAX_K);
  const uint32_t sd = (uint32_t)hm_->index.sd;
  *needed_j = kcap > sd ? kcap - sd : 0;
  return true;
}

// ---- batches -------------------------------------------------------------------------------------------
DeviceBatch* Engine::create_batch(const char* blob, const uint64_t* offsets, uint64_t n, const anl_search_params& p,
                                  std::string* err, int* status) {
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
  DeviceBatch* b = new DeviceBatch();
  auto fail = [&]() -> DeviceBatch* {
    free_batch(b);
    return nullptr;
  };
  b->n = (uint32_t)n;
  b->params = p;
  uint32_t needed_j = 0;
  if (!make_batch_params(p, &b->bp, &needed_j, err)) {
    *status = ANL_ERR_UNSUPPORTED;
    return fail();
  }
  if (!ensure_msets(needed_j, err)) {
    *status = ANL_ERR_UNSUPPORTED;
    return fail();
  }
  if (cudaSetDevice(device_) != cudaSuccess) {
    *err = "cudaSetDevice failed";
    return fail();
  }
  b->offsets.assign(offsets, offsets + n + 1);
  b->blob.assign(blob + offsets[0], blob + offsets[n]);
  const uint64_t base = offsets[0];
  for (auto& o : b->offsets) o -= base;

  // host normalisation (src/anahash.rs:50-80) into fixed-stride rows: len, flags, symbols
  // row stride from the longest query in bytes (a symbol consumes at least one byte)
  uint64_t maxbytes = 0;
  for (uint64_t i = 0; i < n; ++i) maxbytes = std::max<uint64_t>(maxbytes, b->offsets[i + 1] - b->offsets[i]);
  const uint32_t stride = (uint32_t)((std::min<uint64_t>(maxbytes, 254) + 2 + 15) & ~15ull);
  b->host_flags.assign(n, 0);
  uint8_t* rows = nullptr;
  if (cudaMallocHost(reinterpret_cast<void**>(&rows), std::max<size_t>((size_t)n * stride, 16)) != cudaSuccess) {
    *err = "cudaMallocHost failed";
    return fail();
  }
  b->h_rows = rows;
  const Alphabet& ab = hm_->alphabet;
  const unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  auto encode_range = [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) {
      uint8_t* row = rows + (size_t)i * stride;
      const char* s = b->blob.data() + b->offsets[i];
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
  };
  if (n < 4096 || nthreads == 1) {
    encode_range(0, n);
  } else {
    std::vector<std::thread> th;
    const uint64_t per = (n + nthreads - 1) / nthreads;
    for (unsigned t = 0; t < nthreads; ++t) {
      const uint64_t lo = t * per, hi = std::min<uint64_t>(n, lo + per);
      if (lo < hi) th.emplace_back(encode_range, lo, hi);
    }
    for (auto& t : th) t.join();
  }
  for (uint64_t i = 0; i < n; ++i) {
    if (b->host_flags[i] == 2) {
      *err = "query " + std::to_string(i) + " is longer than " + std::to_string(ANL_MAX_SYMBOLS) + " symbols";
      *status = ANL_ERR_UNSUPPORTED;
      return fail();
    }
  }
  b->bp.query_stride = stride;

  bool ok = true;
  std::string e2;
  ok = ok && dev_alloc(&b->d_rows, (size_t)n * stride, &e2);
  ok = ok && dev_alloc(&b->d_hits, (size_t)n * b->bp.hit_cap, &e2);
  ok = ok && dev_alloc(&b->d_hit_count, n, &e2);
  ok = ok && dev_alloc(&b->d_qflags, n, &e2);
  ok = ok && dev_alloc(&b->d_out, (size_t)n * b->bp.out_cap, &e2);
  ok = ok && dev_alloc(&b->d_out_count, n, &e2);
  ok = ok && dev_alloc(reinterpret_cast<uint8_t**>(&b->d_scratch), score_scratch_bytes(b->bp, sm_count_), &e2);
  ok = ok && dev_alloc(&b->d_work, 2, &e2);
  ok = ok && dev_alloc(&b->d_counters, 1, &e2);
  if (!ok) {
    *err = e2;
    return fail();
  }
  if (n > 0 && cudaMemcpyAsync(b->d_rows, rows, (size_t)n * stride, cudaMemcpyHostToDevice, stream_) != cudaSuccess) {
    *err = "H2D copy failed";
    return fail();
  }
  if (cudaStreamSynchronize(stream_) != cudaSuccess) {
    *err = "H2D sync failed";
    return fail();
  }
  *status = ANL_OK;
  return b;
}

void Engine::free_batch(DeviceBatch* b) {
  if (!b) return;
  if (b->h_rows) cudaFreeHost(b->h_rows);
  for (void* p : {(void*)b->d_rows, (void*)b->d_hits, (void*)b->d_hit_count, (void*)b->d_qflags, (void*)b->d_out,
                  (void*)b->d_out_count, b->d_scratch, (void*)b->d_work, (void*)b->d_counters})
    if (p) cudaFree(p);
  for (auto& ev : b->events)
    if (ev) cudaEventDestroy(ev);
  delete b;
}

bool Engine::run_batch(DeviceBatch* b, cudaStream_t stream, std::string* err) {
  if (!stream) stream = stream_;
  CU_TRY(cudaSetDevice(device_));
  LaunchBuffers lb;
  lb.queries = b->d_rows;
  lb.qlist = nullptr;
  lb.n = b->n;
  lb.hits = b->d_hits;
  lb.hit_count = b->d_hit_count;
  lb.qflags = b->d_qflags;
  lb.out = b->d_out;
  lb.out_count = b->d_out_count;
  lb.scratch = b->d_scratch;
  lb.work = b->d_work;
  lb.counters = b->d_counters;
  if (b->runs_recorded >= 1024) b->runs_recorded = 0;  // keep the newest window
  while (b->events.size() < (size_t)(b->runs_recorded + 1) * 3) {
    cudaEvent_t ev = nullptr;
    CU_TRY(cudaEventCreate(&ev));
    b->events.push_back(ev);
  }
  cudaEvent_t* ev = b->events.data() + (size_t)b->runs_recorded * 3;
  CU_TRY(cudaMemsetAsync(b->d_counters, 0, sizeof(Counters), stream));
  CU_TRY(cudaEventRecord(ev[0], stream));
  CU_TRY(launch_probe(d_ix_, h_ix_, b->bp, lb, sm_count_, stream));
  CU_TRY(cudaEventRecord(ev[1], stream));
  CU_TRY(launch_score(d_ix_, h_ix_, b->bp, lb, sm_count_, stream));
  CU_TRY(cudaEventRecord(ev[2], stream));
  b->last_done = ev[2];
  ++b->runs_recorded;
  b->ran = true;
  return true;
}

bool Engine::timings(DeviceBatch* b, float* probe_ms, float* score_ms, std::string* err) {
  if (!b->ran) {
    *err = "batch has not been run";
    return false;
  }
  // averages over the runs since the previous call (CUDA events on the launching stream)
  CU_TRY(cudaEventSynchronize(b->last_done));
  double p = 0, s = 0;
  for (uint32_t r = 0; r < b->runs_recorded; ++r) {
    float a = 0, c = 0;
    CU_TRY(cudaEventElapsedTime(&a, b->events[r * 3 + 0], b->events[r * 3 + 1]));
    CU_TRY(cudaEventElapsedTime(&c, b->events[r * 3 + 1], b->events[r * 3 + 2]));
    p += a;
    s += c;
  }
  const uint32_t nr = std::max(1u, b->runs_recorded);
  *probe_ms = (float)(p / nr);
  *score_ms = (float)(s / nr);
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
  out->probe_steps = c.table_steps;
  out->anagram_hits = c.anagram_hits;
  out->instance_pairs = c.instance_pairs;
  out->dl_pairs = c.dl_pairs;
  out->dl_cells = c.dl_cells;
  out->survivors = c.survivors;
  out->results = c.results;
  out->reruns = b->reruns;
  out->filter_pass = c.filter_pass;
  out->postings = c.postings;
  return true;
}

// ---- host post-pass ---------------------------------------------------------------------------------------
static inline double variant_score(const anl_variant& v, float fw) {  // src/types.rs:335-341
  if (fw == 0.0f) return v.dist_score;
  return (v.dist_score + ((double)fw * v.freq_score)) / (1.0 + (double)fw);
}
static void rank_variants(std::vector<anl_variant>& r, float fw) {  // src/lib.rs:1667 + src/types.rs:344-365
  std::stable_sort(r.begin(), r.end(), [fw](const anl_variant& a, const anl_variant& b) {
    if (fw > 0.0f) return variant_score(a, fw) > variant_score(b, fw);
    if (a.dist_score != b.dist_score) return a.dist_score > b.dist_score;
    return a.freq_score > b.freq_score;
  });
}
static void crop_variants(std::vector<anl_variant>& r, size_t max_matches, float fw) {  // src/lib.rs:1536-1589
  if (max_matches == 0 || r.size() <= max_matches) return;
  const double last = variant_score(r[max_matches - 1], fw), cropped = variant_score(r[max_matches], fw);
  if (cropped < last) {
    r.resize(max_matches);
    return;
  }
  size_t early = 0, late = 0;
  for (size_t i = 0; i < r.size(); ++i) {
    if (r[i].dist_score == cropped && early == 0) early = i;
    if (r[i].dist_score < cropped) {
      late = i;
      break;
    }
  }
  if (early > 0)
    r.resize(early + 1);
  else if (late > 0)
    r.resize(late + 1);
}
static void cutoff_variants(std::vector<anl_variant>& r, double cutoff_threshold, float fw) {  // src/lib.rs:1598-1622
  if (!(cutoff_threshold >= 1.0) || r.empty()) return;
  const double best = variant_score(r[0], fw);
  for (size_t i = 1; i < r.size(); ++i)
    if (variant_score(r[i], fw) <= best / cutoff_threshold) {
      r.resize(i);
      return;
    }
}

void Engine::finish_query(const DeviceBatch& b, uint64_t qi, const OutRec* recs, uint32_t count,
                          std::vector<anl_variant>* out) const {
  out->clear();
  out->reserve(count);
  for (uint32_t i = 0; i < count; ++i) out->push_back(anl_variant{recs[i].vocab_id, recs[i].dist_score, recs[i].freq_score, ANL_NO_VIA});
  if (b.bp.finish_mode == FINISH_FULL) return;
  const std::string input(b.blob.data() + b.offsets[qi], (size_t)(b.offsets[qi + 1] - b.offsets[qi]));
  const float fw = b.params.freq_weight;
  for (anl_variant& v : *out) v.dist_score *= hm_->compute_confusable_weight(input, v.vocab_id);  // src/lib.rs:1660-1662
  rank_variants(*out, fw);
  if (b.bp.finish_mode == FINISH_GATHER) crop_variants(*out, (size_t)b.params.max_matches, fw);
  cutoff_variants(*out, b.params.cutoff_threshold, fw);
}

bool Engine::rerun_overflow(DeviceBatch* b, const std::vector<uint32_t>& which, const std::vector<uint32_t>& hit_counts,
                            const std::vector<uint32_t>& out_counts, std::vector<std::vector<OutRec>>* recs,
                            std::string* err, int* status) {
  (void)out_counts;
  *status = ANL_ERR_CUDA;
  const uint32_t m = (uint32_t)which.size();
  uint32_t cap = b->bp.hit_cap;
  for (uint32_t i = 0; i < m; ++i) cap = std::max(cap, hit_counts[which[i]]);
  cap = (cap + 31) & ~31u;
  BatchParams bp = b->bp;
  bp.hit_cap = cap;
  bp.out_cap = cap;  // results <= survivors <= hits: cannot overflow again
  uint32_t* d_qlist = nullptr;
  uint32_t *d_hits = nullptr, *d_hit_count = nullptr, *d_qflags = nullptr, *d_out_count = nullptr;
  OutRec* d_out = nullptr;
  uint8_t* d_scratch = nullptr;
  auto cleanup = [&]() {
    for (void* p : {(void*)d_qlist, (void*)d_hits, (void*)d_hit_count, (void*)d_qflags, (void*)d_out_count, (void*)d_out,
                    (void*)d_scratch})
      if (p) cudaFree(p);
  };
  bool ok = dev_alloc(&d_qlist, m, err) && dev_alloc(&d_hits, (size_t)m * cap, err) && dev_alloc(&d_hit_count, m, err) &&
            dev_alloc(&d_qflags, m, err) && dev_alloc(&d_out_count, m, err) && dev_alloc(&d_out, (size_t)m * cap, err) &&
            dev_alloc(&d_scratch, score_scratch_bytes(bp, sm_count_), err);
  if (!ok) {
    cleanup();
    return false;
  }
  auto body = [&]() -> bool {
    CU_TRY(cudaMemcpyAsync(d_qlist, which.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
    LaunchBuffers lb;
    lb.queries = b->d_rows;
    lb.qlist = d_qlist;
    lb.n = m;
    lb.hits = d_hits;
    lb.hit_count = d_hit_count;
    lb.qflags = d_qflags;
    lb.out = d_out;
    lb.out_count = d_out_count;
    lb.scratch = d_scratch;
    lb.work = b->d_work;
    lb.counters = nullptr;
    CU_TRY(launch_probe(d_ix_, h_ix_, bp, lb, sm_count_, stream_));
    CU_TRY(launch_score(d_ix_, h_ix_, bp, lb, sm_count_, stream_));
    std::vector<uint32_t> oc(m), fl(m);
    CU_TRY(cudaMemcpyAsync(oc.data(), d_out_count, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
    CU_TRY(cudaMemcpyAsync(fl.data(), d_qflags, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
    CU_TRY(cudaStreamSynchronize(stream_));
    recs->resize(m);
    for (uint32_t i = 0; i < m; ++i) {
      if (fl[i] & (QF_HIT_OVERFLOW | QF_OUT_OVERFLOW | QF_UNSUPPORTED)) {
        *err = "internal error: overflow persisted after rerun";
        return false;
      }
      (*recs)[i].resize(oc[i]);
      if (oc[i])
        CU_TRY(cudaMemcpy((*recs)[i].data(), d_out + (size_t)i * cap, oc[i] * sizeof(OutRec), cudaMemcpyDeviceToHost));
    }
    return true;
  };
  ok = body();
  cleanup();
  if (ok) *status = ANL_OK;
  return ok;
}

bool Engine::fetch_batch(DeviceBatch* b, ResultSet* out, std::string* err, int* status) {
  *status = ANL_ERR_CUDA;
  if (!b->ran) {
    *err = "batch has not been run";
    *status = ANL_ERR_INVALID;
    return false;
  }
  CU_TRY(cudaSetDevice(device_));
  const uint32_t n = b->n;
  const uint32_t ocap = b->bp.out_cap;
  std::vector<uint32_t> out_count(n), qflags(n), hit_count(n);
  OutRec* h_out = nullptr;
  CU_TRY(cudaMallocHost(reinterpret_cast<void**>(&h_out), std::max<size_t>((size_t)n * ocap, 1) * sizeof(OutRec)));
  auto body = [&]() -> bool {
    CU_TRY(cudaEventSynchronize(b->last_done));
    if (n) {
      CU_TRY(cudaMemcpyAsync(out_count.data(), b->d_out_count, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
      CU_TRY(cudaMemcpyAsync(qflags.data(), b->d_qflags, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
      CU_TRY(cudaMemcpyAsync(hit_count.data(), b->d_hit_count, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
      CU_TRY(cudaMemcpyAsync(h_out, b->d_out, (size_t)n * ocap * sizeof(OutRec), cudaMemcpyDeviceToHost, stream_));
    }
    CU_TRY(cudaStreamSynchronize(stream_));
    return true;
  };
  if (!body()) {
    cudaFreeHost(h_out);
    return false;
  }
  // queries whose fixed-capacity buffers overflowed are run again with exact capacities
  std::vector<uint32_t> which;
  for (uint32_t i = 0; i < n; ++i) {
    if (qflags[i] & QF_UNSUPPORTED) {
      *err = "query " + std::to_string(i) + ": max_anagram_distance after thresholding exceeds " +
             std::to_string(ANL_MAX_K) + " (or its deletion neighbourhood is too large) -- unsupported by the GPU path";
      *status = ANL_ERR_UNSUPPORTED;
      cudaFreeHost(h_out);
      return false;
    }
    if (qflags[i] & (QF_HIT_OVERFLOW | QF_OUT_OVERFLOW)) which.push_back(i);
  }
  std::vector<std::vector<OutRec>> rerun_recs;
  if (!which.empty()) {
    // a hit overflow hides the true result count; size by hits, which bounds everything
    std::vector<uint32_t> hc = hit_count;
    for (uint32_t i : which) hc[i] = std::max(hc[i], std::max(out_count[i], b->bp.hit_cap));
    if (!rerun_overflow(b, which, hc, out_count, &rerun_recs, err, status)) {
      cudaFreeHost(h_out);
      return false;
    }
    b->reruns += which.size();
  }
  // assemble (parallel over queries when the confusable post-pass makes it worthwhile)
  std::vector<int32_t> rerun_index(n, -1);
  for (size_t k = 0; k < which.size(); ++k) rerun_index[which[k]] = (int32_t)k;
  out->offsets.assign((size_t)n + 1, 0);
  out->flags.assign(n, 0);
  std::vector<std::vector<anl_variant>> per(n);
  auto work = [&](uint32_t lo, uint32_t hi) {
    for (uint32_t i = lo; i < hi; ++i) {
      if (qflags[i] & QF_EMPTY) {
        if (b->host_flags[i] == 0) out->flags[i] |= 1;
        continue;
      }
      if (rerun_index[i] >= 0) {
        const auto& r = rerun_recs[rerun_index[i]];
        finish_query(*b, i, r.data(), (uint32_t)r.size(), &per[i]);
      } else {
        finish_query(*b, i, h_out + (size_t)i * ocap, std::min(out_count[i], ocap), &per[i]);
      }
    }
  };
  const unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < 2048 || nthreads == 1) {
    work(0, n);
  } else {
    std::vector<std::thread> th;
    const uint32_t per_t = (n + nthreads - 1) / nthreads;
    for (unsigned t = 0; t < nthreads; ++t) {
      const uint32_t lo = t * per_t, hi = std::min<uint32_t>(n, lo + per_t);
      if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto& t : th) t.join();
  }
  cudaFreeHost(h_out);
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    out->offsets[i] = total;
    total += per[i].size();
  }
  out->offsets[n] = total;
  out->variants.resize(total);
  for (uint32_t i = 0; i < n; ++i)
    if (!per[i].empty()) memcpy(out->variants.data() + out->offsets[i], per[i].data(), per[i].size() * sizeof(anl_variant));
  b->results = total;
  *status = ANL_OK;
  return true;
}

bool Engine::find_variants_batch(const char* blob, const uint64_t* offsets, uint64_t n, const anl_search_params& p,
                                 ResultSet* out, std::string* err, int* status) {
  out->offsets.assign(1, 0);
  out->variants.clear();
  out->flags.clear();
  const uint64_t CHUNK = 1u << 20;
  for (uint64_t lo = 0; lo < n || (n == 0 && lo == 0); lo += CHUNK) {
    const uint64_t m = std::min(CHUNK, n - lo);
    DeviceBatch* b = create_batch(blob, offsets + lo, m, p, err, status);
    if (!b) return false;
    ResultSet part;
    bool ok = run_batch(b, nullptr, err);
    if (!ok) *status = ANL_ERR_CUDA;
    ok = ok && fetch_batch(b, &part, err, status);
    free_batch(b);
    if (!ok) return false;
    const uint64_t base = out->variants.size();
    out->variants.insert(out->variants.end(), part.variants.begin(), part.variants.end());
    for (uint64_t i = 1; i <= m; ++i) out->offsets.push_back(base + part.offsets[i]);
    out->flags.insert(out->flags.end(), part.flags.begin(), part.flags.end());
    if (n == 0) break;
  }
  *status = ANL_OK;
  return true;
}

}  // namespace anl
