// capi.cpp -- the extern "C" boundary declared in include/analiticcl_b200.h.
#include <algorithm>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/analiticcl_b200.h"
#include "editscript_fixed.h"
#include "engine.h"
#include "host_model.h"
#include "hostpool.h"
#include "search.h"

using namespace anl;

struct anl_model {
  HostModel host;
  // one Engine per device that holds a replica of the index (anl_model_build: one; anl_model_build_multi: several)
  std::vector<std::unique_ptr<Engine>> engines;
  Engine& engine;  // engines[0]: the device-batch API and the sharded mode work on the first device
  anl_model(const Weights& w, int debug) : host(w, debug), engines(make_first(&host)), engine(*engines[0]) {}
  std::vector<Engine*> replicas() const {
    std::vector<Engine*> v;
    for (const auto& e : engines)
      if (e->uploaded()) v.push_back(e.get());
    return v;
  }

 private:
  static std::vector<std::unique_ptr<Engine>> make_first(HostModel* h) {
    std::vector<std::unique_ptr<Engine>> v;
    v.emplace_back(new Engine(h));
    return v;
  }
};
struct anl_result_set {
  ResultSet rs;
};
struct anl_match_set {
  uint64_t logical_lookups = 0, distinct_lookups = 0;
  PodBuffer<anl_match> matches;
  // all variant lists back to back; anl_match.variants points into it.  Shared: a consolidated match set refers to the
  // lists of the set it was made from instead of copying them, and keeps them alive when that set is freed.
  std::shared_ptr<PodBuffer<anl_variant>> variants = std::make_shared<PodBuffer<anl_variant>>();
  std::shared_ptr<Segmentation> seg;  // the producer's segmentation (anl_find_all_matches), reused by the consolidation
  // tags assigned by context rules (Match.tag / Match.seqnr, src/search.rs:57-60): CSR over the matches; empty = none
  std::vector<uint64_t> tag_first;
  std::vector<uint16_t> tag;
  std::vector<uint8_t> seqnr;
};
struct anl_device_batch {
  DeviceBatch* b;
};

static thread_local std::string g_last_error;
static anl_status fail(anl_status code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// No exception crosses the C boundary: every entry point that can allocate or call into the engine is a function
// try block ending here (std::bad_alloc from the result arrays is the realistic case).
static anl_status on_exception() noexcept {
  try {
    try {
      throw;
    } catch (const std::bad_alloc&) {
      return fail(ANL_ERR_INVALID, "out of memory");
    } catch (const std::exception& e) {
      return fail(ANL_ERR_INVALID, std::string("internal error: ") + e.what());
    } catch (...) {
      return fail(ANL_ERR_INVALID, "internal error: unknown exception");
    }
  } catch (...) {  // even the message could not be stored
    return ANL_ERR_INVALID;
  }
}

// The positions i of [lo, hi) with pred(i), ascending, into *out -- on all cores (count per range, prefix, write): the
// selections of anl_find_all_matches run over millions of segments per window.
template <class Pred>
static void parallel_select(uint64_t lo, uint64_t hi, Pred pred, std::vector<uint64_t>* out) {
  const unsigned nt_max = host_threads();
  std::vector<uint64_t> count(nt_max + 1, 0);
  std::vector<std::pair<uint64_t, uint64_t>> range(nt_max, {0, 0});
  const unsigned used = parallel_ranges(hi - lo, 1u << 15, [&](unsigned t, uint64_t a, uint64_t b) {
    range[t] = {lo + a, lo + b};
    uint64_t c = 0;
    for (uint64_t i = lo + a; i < lo + b; ++i) c += pred(i) ? 1 : 0;
    count[t + 1] = c;
  });
  for (unsigned t = 0; t < used; ++t) count[t + 1] += count[t];
  out->resize(count[used]);
  uint64_t* dst = out->data();
  parallel_ranges(used, 1, [&](unsigned, uint64_t ta, uint64_t tb) {
    for (uint64_t t = ta; t < tb; ++t) {
      uint64_t w = count[t];
      for (uint64_t i = range[t].first; i < range[t].second; ++i)
        if (pred(i)) dst[w++] = i;
    }
  });
}

extern "C" {

const char* anl_last_error(void) { return g_last_error.c_str(); }
uint64_t anl_kernel_launches(void) { return kernel_launches(); }
const char* anl_version(void) { return "analiticcl_b200 0.1 (reference: analiticcl 0.4.9; sm_100a)"; }

void anl_weights_default(anl_weights* w) {
  Weights d;
  w->ld = d.ld;
  w->lcs = d.lcs;
  w->prefix = d.prefix;
  w->suffix = d.suffix;
  w->case_ = d.case_;
}
void anl_search_params_default(anl_search_params* p) {  // src/types.rs:170-192
  memset(p, 0, sizeof *p);
  p->max_anagram_distance = anl_distance_threshold{ANL_THRESHOLD_ABSOLUTE, 0.f, 3};
  p->max_edit_distance = anl_distance_threshold{ANL_THRESHOLD_ABSOLUTE, 0.f, 3};
  p->max_matches = 20;
  p->score_threshold = 0.25;
  p->cutoff_threshold = 2.0;
  p->stop_criterion = ANL_STOP_EXHAUSTIVE;
  p->max_ngram = 3;
  p->lm_order = 3;
  p->max_seq = 250;
  p->single_thread = 0;
  p->context_weight = 0.0f;
  p->variantmodel_weight = 3.0f;
  p->lm_weight = 1.0f;
  p->contextrules_weight = 1.0f;
  p->freq_weight = 0.0f;
  p->consolidate_matches = 1;
  p->unicodeoffsets = 0;
}
void anl_vocab_params_default(anl_vocab_params* p) {  // src/vocab.rs:121-131
  p->text_column = 0;
  p->freq_column = 1;
  p->freq_handling = ANL_FREQ_MAX;
  p->vocab_type = ANL_VOCAB_INDEXED;
  p->index = 0;
}

static Weights to_weights(const anl_weights* w) {
  Weights r;
  if (w) {
    r.ld = w->ld;
    r.lcs = w->lcs;
    r.prefix = w->prefix;
    r.suffix = w->suffix;
    r.case_ = w->case_;
  }
  return r;
}
static VocabParams to_vocab_params(const anl_vocab_params* p) {
  VocabParams r;
  if (p) {
    r.text_column = p->text_column;
    r.freq_column = p->freq_column;
    r.freq_handling = p->freq_handling;
    r.vocab_type = p->vocab_type;
    r.index = p->index;
  }
  return r;
}

anl_status anl_model_new(const char* alphabet_file, const anl_weights* weights, int32_t debug, anl_model** out) try {
  if (!alphabet_file || !out) return fail(ANL_ERR_INVALID, "null argument");
  anl_model* m = new (std::nothrow) anl_model(to_weights(weights), debug);
  if (!m) return fail(ANL_ERR_INVALID, "out of memory");
  std::string err;
  if (!m->host.read_alphabet_file(alphabet_file, &err)) {
    delete m;
    return fail(ANL_ERR_IO, err);
  }
  m->host.init_vocab();
  *out = m;
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_model_new_from_tsv(const char* alphabet_tsv, size_t len, const anl_weights* weights, int32_t debug,
                                  anl_model** out) try {
  if (!alphabet_tsv || !out) return fail(ANL_ERR_INVALID, "null argument");
  anl_model* m = new (std::nothrow) anl_model(to_weights(weights), debug);
  if (!m) return fail(ANL_ERR_INVALID, "out of memory");
  m->host.read_alphabet_text(std::string(alphabet_tsv, len));
  m->host.init_vocab();
  *out = m;
  return ANL_OK;
} catch (...) {
  return on_exception();
}
void anl_model_free(anl_model* m) { delete m; }

anl_status anl_model_read_vocabulary(anl_model* m, const char* filename, const anl_vocab_params* params) try {
  if (!m || !filename) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->host.read_vocabulary(filename, to_vocab_params(params), &err)) return fail(ANL_ERR_IO, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_model_add_to_vocabulary(anl_model* m, const char* text, size_t len, int32_t has_frequency, uint32_t frequency,
                                       const anl_vocab_params* params, uint64_t* vocab_id) try {
  if (!m || !text) return fail(ANL_ERR_INVALID, "null argument");
  uint64_t id = m->host.add_to_vocabulary(text, len, has_frequency != 0, frequency, to_vocab_params(params));
  if (vocab_id) *vocab_id = id;
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_model_add_variant(anl_model* m, uint64_t ref_id, const char* text, size_t len, double score, int32_t has_frequency,
                                 uint32_t frequency, const anl_vocab_params* params, int32_t* added) try {
  if (!m || !text) return fail(ANL_ERR_INVALID, "null argument");
  if (ref_id >= m->host.decoder.size()) return fail(ANL_ERR_INVALID, "add_variant: unknown reference id");
  const bool ok = m->host.add_variant(ref_id, text, len, score, has_frequency != 0, frequency, to_vocab_params(params));
  if (added) *added = ok ? 1 : 0;
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_model_read_variants(anl_model* m, const char* filename, const anl_vocab_params* params, int32_t transparent) try {
  if (!m || !filename) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->host.read_variants(filename, to_vocab_params(params), transparent != 0, &err)) return fail(ANL_ERR_IO, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_model_read_confusablelist(anl_model* m, const char* filename) try {
  if (!m || !filename) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->host.read_confusablelist(filename, &err)) return fail(ANL_ERR_IO, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_model_add_to_confusables(anl_model* m, const char* editscript, double weight) try {
  if (!m || !editscript) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->host.add_to_confusables(editscript, weight, &err)) return fail(ANL_ERR_INVALID, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
void anl_model_set_confusables_before_pruning(anl_model* m) {
  if (m) m->host.confusables_before_pruning = true;
}

// device side of build / load_index: one replica of the host index per listed device
static anl_status upload_replicas(anl_model* m, const int32_t* devices, uint32_t n_devices) {
  std::string err;
  while (m->engines.size() > std::max<uint32_t>(n_devices, 1)) m->engines.pop_back();
  while (m->engines.size() < n_devices) m->engines.emplace_back(new Engine(&m->host));
  for (uint32_t i = 0; i < std::max<uint32_t>(n_devices, 1); ++i)
    if (!m->engines[i]->upload(devices ? devices[i] : -1, &err)) return fail(ANL_ERR_CUDA, err);
  return ANL_OK;
}
// Which build?  ANL_GPU_BUILD=1 / 0 forces the device / host build; by default lexicons of a million entries and more
// are built on the device (seconds instead of tens of seconds), smaller ones on the host (the fixed cost of the device
// pipeline -- allocations, a dozen launches, the download -- is what a small host build takes in total).
static bool build_any(anl_model* m, int sd, uint32_t shard, uint32_t n_shards, int32_t device, int32_t where, std::string* err) {
  int use_gpu = where;  // -1 = choose
  if (use_gpu < 0) {
    if (const char* e = getenv("ANL_GPU_BUILD")) use_gpu = atoi(e) ? 1 : 0;
  }
  if (use_gpu < 0) use_gpu = m->host.decoder.size() >= 1000000 ? 1 : 0;
  return use_gpu ? gpu_build_index(&m->host, sd, shard, n_shards, device, err) : m->host.build_index(sd, shard, n_shards, err);
}
anl_status anl_model_build_sharded(anl_model* m, int32_t device, uint32_t shard, uint32_t n_shards) try {
  if (!m) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  int sd = 1;
  if (const char* e = getenv("ANL_SD")) sd = atoi(e) ? 1 : 0;
  if (!build_any(m, sd, shard, n_shards, device, -1, &err)) return fail(ANL_ERR_UNSUPPORTED, err);
  return upload_replicas(m, &device, 1);
} catch (...) {
  return on_exception();
}
anl_status anl_model_build_on(anl_model* m, int32_t device, int32_t build_on_device) try {
  if (!m) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  int sd = 1;
  if (const char* e = getenv("ANL_SD")) sd = atoi(e) ? 1 : 0;
  if (!build_any(m, sd, 0, 1, device, build_on_device ? 1 : 0, &err)) return fail(ANL_ERR_UNSUPPORTED, err);
  return upload_replicas(m, &device, 1);
} catch (...) {
  return on_exception();
}
anl_status anl_model_build_multi(anl_model* m, const int32_t* devices, uint32_t n_devices) try {
  if (!m || (!devices && n_devices > 0)) return fail(ANL_ERR_INVALID, "null argument");
  if (n_devices == 0 || n_devices > 64) return fail(ANL_ERR_INVALID, "anl_model_build_multi: between 1 and 64 devices");
  for (uint32_t i = 0; i < n_devices; ++i)
    for (uint32_t j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return fail(ANL_ERR_INVALID, "anl_model_build_multi: a device is listed twice");
  std::string err;
  int sd = 1;
  if (const char* e = getenv("ANL_SD")) sd = atoi(e) ? 1 : 0;
  if (!build_any(m, sd, 0, 1, devices[0], -1, &err)) return fail(ANL_ERR_UNSUPPORTED, err);
  return upload_replicas(m, devices, n_devices);
} catch (...) {
  return on_exception();
}
uint32_t anl_model_device_count(const anl_model* m) { return m ? (uint32_t)m->replicas().size() : 0; }

static anl_status consolidate_impl(const HostModel* hm, const anl_match_set* in, const char* text, size_t len,
                                   const anl_search_params* params, anl_match_set** out);
// learn_variants (src/lib.rs:1062-1139).  The lookups are the model's own batched GPU lookups: strict mode is ONE
// anl_find_variants_batch over all inputs (the reference's par_iter over find_variants, :1083-1088), else
// find_all_matches + the sequence stage per input; the found (input, variant) pairs are then stored in order.
anl_status anl_model_learn_variants(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n_inputs,
                                    const anl_search_params* params, int32_t strict, int32_t auto_build, uint64_t* count) try {
  if (!m || !offsets || !params || (!blob && n_inputs > 0)) return fail(ANL_ERR_INVALID, "null argument");
  if (!m->host.built || !m->engine.uploaded())
    return fail(ANL_ERR_NOT_BUILT, "Model has not been built yet! Call build() before learn_variants()");
  if (m->host.index.n_shards > 1) return fail(ANL_ERR_UNSUPPORTED, "learn_variants is not available on a lexicon shard");
  std::vector<HostModel::LearnedVariant> items;
  std::string err;
  int status = ANL_OK;
  if (strict) {
    ResultSet rs;
    if (!find_variants_batch_multi(m->replicas(), blob ? blob : "", offsets, n_inputs, *params, &rs, &err, &status))
      return fail(status ? status : ANL_ERR_CUDA, err);
    for (uint64_t i = 0; i < n_inputs; ++i)
      for (uint64_t j = rs.offsets[i]; j < rs.offsets[i + 1]; ++j)
        items.push_back({std::string(blob + offsets[i], offsets[i + 1] - offsets[i]), rs.variants[j].vocab_id, rs.variants[j].dist_score});
  } else {
    const bool sequence = params->max_ngram > 1 || m->host.have_lm() || !m->host.context_rules.empty();  // :1912
    for (uint64_t i = 0; i < n_inputs; ++i) {
      const char* text = blob + offsets[i];
      const size_t len = offsets[i + 1] - offsets[i];
      anl_match_set* all = nullptr;
      anl_status st = anl_find_all_matches(m, text, len, params, &all);
      if (st != ANL_OK) return st;
      std::unique_ptr<anl_match_set> owner(all), best;
      const anl_match_set* use = all;
      if (sequence) {
        anl_match_set* b = nullptr;
        st = consolidate_impl(&m->host, all, text, len, params, &b);
        if (st != ANL_OK) return st;
        best.reset(b);
        use = b;
      }
      std::vector<uint64_t> cp2byte;
      if (params->unicodeoffsets) {  // match offsets are code points then: back to bytes for the text slice
        const std::vector<uint64_t> map = byte_to_codepoint_map(std::string(text, len));
        cp2byte.assign(map.back() + 1, 0);
        for (size_t b = len + 1; b-- > 0;) cp2byte[map[b]] = b;
      }
      for (size_t k = 0; k < use->matches.size(); ++k) {
        const anl_match& mm = use->matches[k];
        if (!mm.variants || mm.selected < 0 || (uint64_t)mm.selected >= mm.n_variants) continue;
        const uint64_t b = params->unicodeoffsets ? cp2byte[mm.begin] : mm.begin, e = params->unicodeoffsets ? cp2byte[mm.end] : mm.end;
        items.push_back({std::string(text + b, e - b), mm.variants[mm.selected].vocab_id, mm.variants[mm.selected].dist_score});
      }
    }
  }
  const uint64_t added = m->host.learn_apply(items);
  if (count) *count = added;
  if (auto_build) {  // (re)build on the devices that hold the index now
    std::vector<int32_t> devices;
    for (Engine* e : m->replicas()) devices.push_back(e->device());
    int sd = 1;
    if (const char* e = getenv("ANL_SD")) sd = atoi(e) ? 1 : 0;
    if (!build_any(m, sd, 0, 1, devices[0], -1, &err)) return fail(ANL_ERR_UNSUPPORTED, err);
    return upload_replicas(m, devices.data(), (uint32_t)devices.size());
  }
  return ANL_OK;
} catch (...) {
  return on_exception();
}
// Test hooks: the bookkeeping half of learn_variants on explicit (input, result id, score) triples, and an entry's links.
anl_status anl_debug_learn_apply(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n, const uint64_t* vocab_ids,
                                 const double* scores, uint64_t* count) try {
  if (!m || !offsets || (n && (!blob || !vocab_ids || !scores))) return fail(ANL_ERR_INVALID, "null argument");
  std::vector<HostModel::LearnedVariant> items;
  for (uint64_t i = 0; i < n; ++i) items.push_back({std::string(blob + offsets[i], offsets[i + 1] - offsets[i]), vocab_ids[i], scores[i]});
  const uint64_t added = m->host.learn_apply(items);
  if (count) *count = added;
  return ANL_OK;
} catch (...) {
  return on_exception();
}
int64_t anl_debug_vocab_links(const anl_model* m, uint64_t id, int32_t kind, uint64_t* ids, double* scores, size_t cap) {
  if (!m || id >= m->host.decoder.size()) return -1;
  const VocabEntry& v = m->host.decoder[id];
  if (kind == 0) {
    for (size_t i = 0; i < v.variant_of.size() && i < cap; ++i) {
      if (ids) ids[i] = v.variant_of[i].first;
      if (scores) scores[i] = v.variant_of[i].second;
    }
    return (int64_t)v.variant_of.size();
  }
  for (size_t i = 0; i < v.reference_for.size() && i < cap; ++i)
    if (ids) ids[i] = v.reference_for[i];
  return (int64_t)v.reference_for.size();
}
anl_status anl_model_save_index(const anl_model* m, const char* filename) try {
  if (!m || !filename) return fail(ANL_ERR_INVALID, "null argument");
  if (!m->host.built) return fail(ANL_ERR_NOT_BUILT, "Model has not been built yet! Call build() before save_index()");
  std::string err;
  if (!m->host.save_index(filename, &err)) return fail(ANL_ERR_IO, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_model_load_index(anl_model* m, const char* filename, int32_t device) try {
  if (!m || !filename) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->host.load_index(filename, &err)) return fail(ANL_ERR_IO, err);
  return upload_replicas(m, &device, 1);
} catch (...) {
  return on_exception();
}
void anl_model_shard(const anl_model* m, uint32_t* shard, uint32_t* n_shards) {
  if (shard) *shard = m && m->host.built ? m->host.index.shard : 0;
  if (n_shards) *n_shards = m && m->host.built ? m->host.index.n_shards : 1;
}
anl_status anl_model_build(anl_model* m, int32_t device) try { return anl_model_build_sharded(m, device, 0, 1); } catch (...) { return on_exception(); }

int32_t anl_model_has(const anl_model* m, const char* text, size_t len) { return m && m->host.has(text, len) ? 1 : 0; }
int64_t anl_model_vocab_id(const anl_model* m, const char* text, size_t len) { return m ? m->host.vocab_id(text, len) : -1; }
uint64_t anl_model_vocab_size(const anl_model* m) { return m ? m->host.decoder.size() : 0; }
anl_status anl_model_get_vocab(const anl_model* m, uint64_t vocab_id, anl_vocab_info* out) try {
  if (!m || !out) return fail(ANL_ERR_INVALID, "null argument");
  if (vocab_id >= m->host.decoder.size()) return fail(ANL_ERR_INVALID, "vocabulary id out of range");
  const VocabEntry& e = m->host.decoder[vocab_id];
  out->text = e.text.c_str();
  out->text_len = (uint32_t)e.text.size();
  out->frequency = e.frequency;
  out->lexindex = e.lexindex;
  out->vocabtype = e.vocabtype;
  out->tokencount = e.tokencount;
  out->norm_len = (uint32_t)e.syms.size();
  return ANL_OK;
} catch (...) {
  return on_exception();
}
uint32_t anl_model_lexicon_count(const anl_model* m) { return m ? (uint32_t)m->host.lexicons.size() : 0; }
const char* anl_model_lexicon_name(const anl_model* m, uint32_t index) {
  if (!m || index >= m->host.lexicons.size()) return nullptr;
  return m->host.lexicons[index].c_str();
}
uint32_t anl_model_alphabet_size(const anl_model* m) { return m ? m->host.alphabet_size() : 0; }
uint64_t anl_model_index_size(const anl_model* m) { return m && m->host.built ? m->host.index.ana_key.size() : 0; }
uint64_t anl_model_instance_count(const anl_model* m) { return m && m->host.built ? m->host.index.inst_vocab.size() : 0; }
uint64_t anl_model_anagram_count_of_length(const anl_model* m, uint32_t charcount) {
  if (!m || !m->host.built) return 0;
  uint64_t n = 0;
  for (uint16_t cc : m->host.index.ana_charcount) n += (cc == charcount);
  return n;
}
uint32_t anl_model_max_key_bits(const anl_model* m) { return m && m->host.built ? m->host.index.max_key_bits : 0; }

int64_t anl_normalize(const anl_model* m, const char* text, size_t len, uint8_t* out, size_t cap) {
  if (!m || !text) return -1;
  std::vector<uint8_t> syms;
  m->host.alphabet.encode(text, len, &syms);
  // the reference's NormString encodes unknown symbols as alphabet.len() + 1 (src/anahash.rs:76)
  const uint8_t unk = (uint8_t)m->host.alphabet.unk_symbol();
  for (size_t i = 0; i < syms.size() && i < cap; ++i) out[i] = syms[i] == unk ? (uint8_t)(unk + 1) : syms[i];
  return (int64_t)syms.size();
}
int64_t anl_anahash(const anl_model* m, const char* text, size_t len, uint64_t* limbs, size_t cap) {
  if (!m || !text) return -1;
  std::vector<uint64_t> v = m->host.anahash_limbs(text, len);
  for (size_t i = 0; i < v.size() && i < cap; ++i) limbs[i] = v[i];
  return (int64_t)v.size();
}

int64_t anl_shortest_edit_script(const char* src, size_t src_len, const char* dst, size_t dst_len, char* out, size_t cap) {
  std::string text;
  for (const EditInstruction& e : shortest_edit_script(std::string(src, src_len), std::string(dst, dst_len))) {
    text += e.op == 0 ? "=[" : (e.op > 0 ? "+[" : "-[");
    text += e.text;
    text += "]";
  }
  if (out && cap) {
    const size_t n = std::min(text.size(), cap - 1);
    memcpy(out, text.data(), n);
    out[n] = 0;
  }
  return (int64_t)text.size();
}
int64_t anl_shortest_edit_script_fixed(const char* src, size_t src_len, const char* dst, size_t dst_len, char* out, size_t cap) {
  // host builds of the fixed-capacity implementation (csrc/editscript_fixed.h): over bytes for pure-ASCII pairs
  // (exactly what the confusable kernel runs per thread), over Unicode scalar values otherwise (what the host
  // post-pass runs for the pairs the kernel declines)
  bool ascii = true;
  for (size_t i = 0; i < src_len; ++i) ascii = ascii && (unsigned char)src[i] < 0x80;
  for (size_t i = 0; i < dst_len; ++i) ascii = ascii && (unsigned char)dst[i] < 0x80;
  std::string text;
  if (ascii) {
    if (src_len > (size_t)esf::MAXLEN || dst_len > (size_t)esf::MAXLEN) return -1;
    esf::View v[esf::MAXSEG];
    const uint8_t* a = reinterpret_cast<const uint8_t*>(src);
    const uint8_t* b = reinterpret_cast<const uint8_t*>(dst);
    const int nv = esf::shortest_edit_script(a, (int)src_len, b, (int)dst_len, v);
    if (nv < 0) return -1;
    for (int i = 0; i < nv; ++i) {
      text += v[i].op == 0 ? "=[" : (v[i].op > 0 ? "+[" : "-[");
      text.append(reinterpret_cast<const char*>((v[i].op > 0 ? b : a) + v[i].pos), v[i].len);
      text += "]";
    }
  } else {
    EditView v[64];
    size_t nv = 0;
    if (!edit_views_fixed(src, src_len, dst, dst_len, v, &nv)) return -1;
    for (size_t i = 0; i < nv; ++i) {
      text += v[i].op == 0 ? "=[" : (v[i].op > 0 ? "+[" : "-[");
      text.append(v[i].p, v[i].n);
      text += "]";
    }
  }
  if (out && cap) {
    const size_t n = std::min(text.size(), cap - 1);
    memcpy(out, text.data(), n);
    out[n] = 0;
  }
  return (int64_t)text.size();
}
int32_t anl_confusable_found_in(const char* pattern, const char* src, size_t src_len, const char* dst, size_t dst_len) {
  Confusable c;
  if (!pattern || !parse_confusable(pattern, 1.0, &c)) return -1;
  return confusable_found_in(c, shortest_edit_script(std::string(src, src_len), std::string(dst, dst_len))) ? 1 : 0;
}

// ---- lookup ------------------------------------------------------------------------------------------
anl_status anl_find_variants_batch(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n_queries,
                                   const anl_search_params* params, anl_result_set** out) try {
  if (!m || !offsets || !params || !out || (!blob && n_queries > 0)) return fail(ANL_ERR_INVALID, "null argument");
  if (!m->host.built || !m->engine.uploaded())
    return fail(ANL_ERR_NOT_BUILT, "Model has not been built yet! Call build() before find_variants()");
  std::unique_ptr<anl_result_set> rs(new anl_result_set());
  std::string err;
  int status = ANL_OK;
  if (!find_variants_batch_multi(m->replicas(), blob ? blob : "", offsets, n_queries, *params, &rs->rs, &err, &status))
    return fail(status ? status : ANL_ERR_CUDA, err);
  *out = rs.release();
  return ANL_OK;
} catch (...) {
  return on_exception();
}
uint64_t anl_result_set_len(const anl_result_set* rs) { return rs ? rs->rs.offsets.size() - 1 : 0; }
const anl_variant* anl_result_set_get(const anl_result_set* rs, uint64_t i, uint64_t* count) {
  if (!rs || i + 1 >= rs->rs.offsets.size()) {
    if (count) *count = 0;
    return nullptr;
  }
  if (count) *count = rs->rs.offsets[i + 1] - rs->rs.offsets[i];
  return rs->rs.variants.data() + rs->rs.offsets[i];
}
const uint64_t* anl_result_set_offsets(const anl_result_set* rs) { return rs ? rs->rs.offsets.data() : nullptr; }
const anl_variant* anl_result_set_variants(const anl_result_set* rs) { return rs ? rs->rs.variants.data() : nullptr; }
uint32_t anl_result_set_flags(const anl_result_set* rs, uint64_t i) {
  return rs && i < rs->rs.flags.size() ? rs->rs.flags[i] : 0;
}
void anl_result_set_free(anl_result_set* rs) { delete rs; }

// find_all_matches: src/lib.rs:1790-1957 without the FST stage (see the header).
// The text is processed in windows of whole hard-delimited batches (bounded temporary memory); each
// window costs two pipelined GPU batches: all unigrams, then the higher-order segments that the
// unigram results do not make redundant (src/search.rs:317-336).
anl_status anl_find_all_matches(anl_model* m, const char* text, size_t len, const anl_search_params* params,
                                anl_match_set** out) try {
  if (!m || !params || !out || (!text && len > 0)) return fail(ANL_ERR_INVALID, "null argument");
  if (!m->host.built || !m->engine.uploaded())
    return fail(ANL_ERR_NOT_BUILT, "Model has not been built yet! Call build() before find_all_matches()");
  // The batch producer (src/lib.rs:1790-1957) as host phases that each run on all cores, around at most
  // two GPU batches per window of text: all unigrams, then every higher-order segment that the unigram
  // results do not make redundant (redundant_match only reads unigram results, src/search.rs:317-336).
  PhaseTimer pt;
  const std::string t(text ? text : "", len);
  std::unique_ptr<anl_match_set> ms_owner(new anl_match_set());  // (released to *out on success only)
  anl_match_set* ms = ms_owner.get();
  std::string err;
  ms->seg = std::make_shared<Segmentation>();
  // the batch producer: on the device for running text, the host loop for short strings (search.h segment_on_device)
  if (!segment_any(m->replicas()[0]->device(), t, params->max_ngram, ms->seg.get(), &err)) return fail(ANL_ERR_CUDA, err);
  SegmentedText& st = ms->seg->st;
  pt.lap("search: segmentation");
  std::vector<uint64_t> cpmap;
  if (params->unicodeoffsets) cpmap = byte_to_codepoint_map(t);  // src/lib.rs:1949-1956
  const size_t nseg = st.segs.size(), nbatch = st.batch_first.size() - 1;
  // per segment: was it looked up, in which pass, where its variants are in that pass's result set
  // (uninitialised, recycled buffers: each window fills its own part in parallel)
  PodBuffer<uint8_t> looked;
  PodBuffer<uint32_t> cnt;
  PodBuffer<uint64_t> off;
  looked.resize(nseg + 1);
  cnt.resize(nseg + 1);
  off.resize(nseg + 1);
  ms->matches.resize(nseg);
  ms->variants->reserve(1);  // (never a null base: a looked-up segment keeps a non-null variants pointer)
  int status = ANL_OK;
  size_t WINDOW = 1u << 20;  // unigram segments per window
  if (const char* e = getenv("ANL_SEARCH_WINDOW")) WINDOW = (size_t)std::max(1, atoi(e));
  std::vector<uint64_t> pick;
  std::string blob;
  std::vector<uint64_t> offs;
  ResultSet rs[2];
  // Running text repeats itself: a window's segments are de-duplicated first (the lookup is a pure function
  // of the segment's text) and only the distinct strings go to the GPU; every occurrence then points at its
  // representative's variant list.  Exact: hash + byte comparison, the first occurrence is the representative.
  PodBuffer<uint64_t> hashes;
  PodBuffer<uint32_t> rep;      // per picked segment: position (in pick) of its first occurrence
  PodBuffer<uint32_t> uniq_of;  // per picked segment: rank of its representative among the distinct strings
  std::vector<uint64_t> uniq;   // positions (in pick) of the distinct strings, ascending
  uint64_t distinct_lookups = 0, logical_lookups = 0;
  auto seg_text = [&](uint64_t k, size_t* n) {
    *n = st.segs[k].end - st.segs[k].begin;
    return t.data() + st.segs[k].begin;
  };
  auto dedupe = [&]() {
    const size_t np = pick.size();
    hashes.resize(np);
    rep.resize(np);
    uniq_of.resize(np);
    parallel_ranges(np, 8192, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) {
        size_t n;
        const char* p = seg_text(pick[i], &n);
        uint64_t h = 0x9E3779B97F4A7C15ULL ^ (n * 0xD6E8FEB86659FD93ULL);
        while (n >= 8) {
          uint64_t w;
          memcpy(&w, p, 8);
          h = (h ^ w) * 0xFF51AFD7ED558CCDULL;
          h ^= h >> 32;
          p += 8;
          n -= 8;
        }
        uint64_t w = 0;
        memcpy(&w, p, n);
        h = (h ^ w) * 0xC4CEB9FE1A85EC53ULL;
        h ^= h >> 29;
        hashes[i] = h;
      }
    });
    // hash-partitioned insertion: thread p owns the strings whose hash falls into its partition, so the
    // tables need no locks and "first occurrence" is the smallest position in every partition
    const unsigned parts = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(host_threads(), np / 4096));
    parallel_ranges(parts, 1, [&](unsigned, uint64_t plo, uint64_t phi) {
      for (uint64_t part = plo; part < phi; ++part) {
        size_t cap = 1024;
        while (cap < 2 * (np / parts + 1) + 1024) cap <<= 1;
        static thread_local std::vector<uint32_t> table;  // position + 1, 0 = empty
        table.assign(cap, 0);
        for (uint64_t i = 0; i < np; ++i) {
          const uint64_t h = hashes[i];
          if ((h >> 40) % parts != part) continue;
          size_t slot = (size_t)h & (cap - 1);
          size_t n;
          const char* p = seg_text(pick[i], &n);
          for (;;) {
            const uint32_t e = table[slot];
            if (e == 0) {
              table[slot] = (uint32_t)i + 1;
              rep[i] = (uint32_t)i;
              break;
            }
            const uint64_t j = e - 1;
            if (hashes[j] == h) {
              size_t m;
              const char* q = seg_text(pick[j], &m);
              if (m == n && memcmp(p, q, n) == 0) {
                rep[i] = (uint32_t)j;
                break;
              }
            }
            slot = (slot + 1) & (cap - 1);
          }
        }
      }
    });
    parallel_select(0, np, [&](uint64_t i) { return rep[i] == i; }, &uniq);
    parallel_ranges(uniq.size(), 1u << 15, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t u = lo; u < hi; ++u) uniq_of[uniq[u]] = (uint32_t)u;
    });
    parallel_ranges(np, 1u << 16, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i)
        if (rep[i] != i) uniq_of[i] = uniq_of[rep[i]];
    });
  };
  auto lookup = [&](int pass) -> bool {
    const size_t np = pick.size();
    rs[pass].offsets.assign(1, 0);
    rs[pass].variants.clear();
    if (np == 0) return true;
    if (np > 0xFFFFFFF0ull) {
      err = "window too large";
      status = ANL_ERR_UNSUPPORTED;
      return false;
    }
    dedupe();
    const size_t nu = uniq.size();
    logical_lookups += np;
    distinct_lookups += nu;
    offs.resize(nu + 1);
    offs[0] = 0;
    for (size_t i = 0; i < nu; ++i) {
      const SegmentSpan& sp = st.segs[pick[uniq[i]]];
      offs[i + 1] = offs[i] + (sp.end - sp.begin);
    }
    blob.resize(offs[nu]);
    parallel_ranges(nu, 4096, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) memcpy(&blob[offs[i]], t.data() + st.segs[pick[uniq[i]]].begin, offs[i + 1] - offs[i]);
    });
    if (!find_variants_batch_multi(m->replicas(), blob.data(), offs.data(), nu, *params, &rs[pass], &err, &status)) return false;
    parallel_ranges(np, 4096, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) {
        const uint64_t k = pick[i];
        const uint64_t u = uniq_of[i];
        looked[k] = (uint8_t)(1 + pass);
        off[k] = rs[pass].offsets[u];
        cnt[k] = (uint32_t)(rs[pass].offsets[u + 1] - rs[pass].offsets[u]);
      }
    });
    return true;
  };
  bool ok = true;
  size_t b0 = 0;
  uint64_t vtotal = 0;
  const unsigned nt_max = host_threads();
  while (ok && b0 < nbatch) {
    // a window: whole batches until it holds enough unigrams
    size_t b1 = b0, unigrams = 0;
    while (b1 < nbatch && (unigrams < WINDOW || b1 == b0)) {
      for (uint64_t k = st.batch_first[b1]; k < st.batch_first[b1 + 1] && st.segs[k].n == 1; ++k) ++unigrams;
      ++b1;
    }
    const uint64_t s0 = st.batch_first[b0], s1 = st.batch_first[b1];
    parallel_ranges(s1 - s0, 1u << 16, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t k = s0 + lo; k < s0 + hi; ++k) {
        looked[k] = 0;
        cnt[k] = 0;
        off[k] = 0;
      }
    });
    parallel_select(s0, s1, [&](uint64_t k) { return st.segs[k].n == 1; }, &pick);
    ok = lookup(0);
    pt.lap("search: unigram lookups");
    if (ok && params->max_ngram > 1) {
      // redundant_match (src/search.rs:317-336), batch by batch; per-thread pick lists keep the segment order
      std::vector<std::vector<uint64_t>> part(nt_max);
      const unsigned used = parallel_ranges(b1 - b0, 64, [&](unsigned tid, uint64_t lo, uint64_t hi) {
        std::vector<uint64_t>& mine = part[tid];
        for (uint64_t bb = b0 + lo; bb < b0 + hi; ++bb) {
          const uint64_t f = st.batch_first[bb], l = st.batch_first[bb + 1];
          for (uint64_t k = f; k < l; ++k) {
            const SegmentSpan& sp = st.segs[k];
            if (sp.n == 1) continue;
            bool redundant = true;
            for (uint64_t u = f; u < l && st.segs[u].n == 1; ++u) {
              if (st.segs[u].begin >= sp.begin && st.segs[u].end <= sp.end) {
                if (cnt[u] == 0 || rs[0].variants[off[u]].dist_score < 1.0) {
                  redundant = false;
                  break;
                }
              }
            }
            if (!redundant) mine.push_back(k);
          }
        }
      });
      pick.clear();
      for (unsigned tid = 0; tid < used; ++tid) pick.insert(pick.end(), part[tid].begin(), part[tid].end());
      pt.lap("search: redundancy pruning");
      ok = lookup(1);
      pt.lap("search: n-gram lookups");
    }
    if (!ok) break;
    // this window's matches.  The variant lists of the window's DISTINCT strings go behind those of the earlier windows
    // (two block copies: unigram pass, n-gram pass); every occurrence of a string points at the one list of its
    // representative -- running text repeats itself, and a copy per occurrence was a quarter of the call.
    const uint64_t n0 = rs[0].offsets.size() ? rs[0].offsets[rs[0].offsets.size() - 1] : 0;
    const uint64_t n1 = (params->max_ngram > 1 && rs[1].offsets.size()) ? rs[1].offsets[rs[1].offsets.size() - 1] : 0;
    const uint64_t base[2] = {vtotal, vtotal + n0};
    vtotal += n0 + n1;
    ms->variants->resize(vtotal);
    anl_variant* vbase = ms->variants->data();
    for (int pass = 0; pass < 2; ++pass) {
      const uint64_t cnt_pass = pass == 0 ? n0 : n1;
      const anl_variant* src = rs[pass].variants.data();
      anl_variant* dstp = vbase + base[pass];
      parallel_ranges(cnt_pass, 1u << 16, [&](unsigned, uint64_t lo, uint64_t hi) { memcpy(dstp + lo, src + lo, (size_t)(hi - lo) * sizeof(anl_variant)); });
    }
    parallel_ranges(s1 - s0, 4096, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t k = s0 + lo; k < s0 + hi; ++k) {
        const SegmentSpan& sp = st.segs[k];
        anl_match& mm = ms->matches[k];
        mm.begin = params->unicodeoffsets ? cpmap[sp.begin] : sp.begin;
        mm.end = params->unicodeoffsets ? cpmap[sp.end] : sp.end;
        mm.n = sp.n;
        mm.n_variants = cnt[k];
        mm.selected = (looked[k] && cnt[k] > 0) ? 0 : -1;
        // index + 1 for now, turned into a pointer below (the buffer may still move)
        mm.variants = looked[k] ? reinterpret_cast<const anl_variant*>(base[looked[k] - 1] + off[k] + 1) : nullptr;
      }
    });
    pt.lap("search: assemble matches");
    b0 = b1;
  }
  if (!ok) return fail(status ? status : ANL_ERR_CUDA, err);
  const anl_variant* vbase = ms->variants->data();
  parallel_ranges(nseg, 1u << 16, [&](unsigned, uint64_t lo, uint64_t hi) {
    for (uint64_t k = lo; k < hi; ++k) {
      anl_match& mm = ms->matches[k];
      if (mm.variants) mm.variants = vbase + (reinterpret_cast<uintptr_t>(mm.variants) - 1);
    }
  });
  if (profile_enabled())
    fprintf(stderr, "[anl profile] search: %llu segment lookups, %llu distinct strings sent to the GPU\n",
            (unsigned long long)logical_lookups, (unsigned long long)distinct_lookups);
  ms->logical_lookups = logical_lookups;
  ms->distinct_lookups = distinct_lookups;
  *out = ms_owner.release();
  return ANL_OK;
} catch (...) {
  return on_exception();
}
// most_likely_sequence per hard-delimited batch (src/lib.rs:1912-1924, 2088-2495): host post-pass over the
// match set of anl_find_all_matches.  The segments are re-derived from the text (cheap next to the lookups) so
// that the match set itself stays a flat list; batches are independent and run on all cores.
static anl_status consolidate_impl(const HostModel* hm, const anl_match_set* in, const char* text, size_t len,
                                   const anl_search_params* params, anl_match_set** out) {
  if (!in || !params || !out || (!text && len > 0)) return fail(ANL_ERR_INVALID, "null argument");
  PhaseTimer pt;
  // (a private copy of the text only where one is read: re-segmentation, boundary tokens of the language model)
  std::string t;
  // the producer's own segmentation when the match set still carries it (anl_find_all_matches), else re-derived
  std::shared_ptr<Segmentation> seg = in->seg;
  const bool resegment = !seg || seg->text_len != len || seg->max_ngram != params->max_ngram;
  if (resegment || (hm && hm->have_lm())) t.assign(text ? text : "", len);
  if (resegment) {
    seg = std::make_shared<Segmentation>();
    std::string err;
    segment_any(-1, t, params->max_ngram, seg.get(), &err);
  }
  const SegmentedText& st = seg->st;
  pt.lap("consolidate: segmentation");
  if (st.segs.size() != in->matches.size())
    return fail(ANL_ERR_INVALID, "match set does not belong to this text / max_ngram (segment count differs)");
  const PodBuffer<Boundary>& bounds = seg->bounds;
  const PodBuffer<BatchDesc>& descs = seg->batches;
  const size_t nbatch = st.batch_first.size() - 1;
  if (descs.size() != nbatch) return fail(ANL_ERR_INVALID, "match set does not belong to this text (batch count differs)");
  const float fw = params->freq_weight;
  const bool scored = hm && (hm->have_lm() || !hm->context_rules.empty());
  const bool run_fst = params->max_ngram > 1 || scored;  // :1912
  SequenceWeights sw;
  sw.max_seq = params->max_seq;
  sw.lm_weight = params->lm_weight;
  sw.variantmodel_weight = params->variantmodel_weight;
  sw.contextrules_weight = params->contextrules_weight;
  // per batch: the chosen (segment, variant) steps; thread-local lists keep the batch order
  const unsigned nt_max = host_threads();
  std::vector<std::vector<SequenceStep>> part(nt_max);
  std::vector<std::vector<StepTags>> part_tags(nt_max);   // parallel to `part` when context rules are loaded
  std::vector<std::vector<uint64_t>> part_count(nt_max);  // steps per batch
  std::vector<std::pair<uint64_t, uint64_t>> range(nt_max, {0, 0});
  const bool want_tags = hm && !hm->context_rules.empty();
  const unsigned used = parallel_ranges(nbatch, 64, [&](unsigned tid, uint64_t lo, uint64_t hi) {
    range[tid] = {lo, hi};
    std::vector<SequenceStep> steps;
    std::vector<StepTags> tags;
    std::vector<uint32_t> count;
    std::vector<uint64_t> first, vids;
    std::vector<double> score;
    for (uint64_t b = lo; b < hi; ++b) {
      const uint64_t s0 = st.batch_first[b], s1 = st.batch_first[b + 1];
      const size_t before = part[tid].size();
      bool chosen = false;
      if (run_fst) {
        count.clear();
        first.clear();
        score.clear();
        vids.clear();
        for (uint64_t k = s0; k < s1; ++k) {
          const anl_match& mm = in->matches[k];
          const uint32_t c = mm.variants ? (uint32_t)mm.n_variants : 0;
          count.push_back(c);
          first.push_back(score.size());
          for (uint32_t j = 0; j < c; ++j) {  // VariantResult::score, src/types.rs:335-341
            const anl_variant& v = mm.variants[j];
            score.push_back(fw == 0.0f ? v.dist_score : (v.dist_score + ((double)fw * v.freq_score)) / (1.0 + (double)fw));
            if (scored) vids.push_back(v.vocab_id);
          }
        }
        const BatchDesc& d = descs[b];
        BatchVariants bv{count.data(), first.data(), score.data(), scored ? vids.data() : nullptr};
        tags.clear();
        if (scored)
          chosen = most_likely_sequence_full(hm, t, bounds.data() + d.begin_index, d.end_index + 1 - d.begin_index, bounds[d.end_index].begin,
                                             st.segs.data() + s0, s1 - s0, bv, sw, &steps, want_tags ? &tags : nullptr);
        else
          chosen = most_likely_sequence(bounds.data() + d.begin_index, d.end_index + 1 - d.begin_index, bounds[d.end_index].begin,
                                        st.segs.data() + s0, s1 - s0, bv, &steps);
        if (chosen) {
          part[tid].insert(part[tid].end(), steps.begin(), steps.end());
          if (want_tags) {
            tags.resize(steps.size());
            part_tags[tid].insert(part_tags[tid].end(), tags.begin(), tags.end());
          }
        }
      }
      if (!chosen) {  // unigram-only models (:1929-1932) and empty lattices (:2261-2267): every match as it is
        for (uint64_t k = s0; k < s1; ++k) part[tid].push_back(SequenceStep{(uint32_t)(k - s0), in->matches[k].selected});
        if (want_tags) part_tags[tid].resize(part[tid].size());
      }
      part_count[tid].push_back(part[tid].size() - before);
    }
  });
  pt.lap("consolidate: lattices");
  std::unique_ptr<anl_match_set> ms_owner(new anl_match_set());
  anl_match_set* ms = ms_owner.get();
  ms->logical_lookups = in->logical_lookups;
  ms->distinct_lookups = in->distinct_lookups;
  // every thread copies the matches of its own batches behind those of the threads before it
  std::vector<uint64_t> first_match(used + 1, 0);
  for (unsigned tid = 0; tid < used; ++tid) first_match[tid + 1] = first_match[tid] + part[tid].size();
  ms->matches.resize(first_match[used]);
  ms->variants = in->variants;  // (shared, not copied: the chosen matches keep pointing at their lists)
  parallel_ranges(used, 1, [&](unsigned, uint64_t tlo, uint64_t thi) {
    for (uint64_t tid = tlo; tid < thi; ++tid) {
      uint64_t o = first_match[tid];
      size_t pos = 0;
      for (uint64_t b = range[tid].first; b < range[tid].second; ++b) {
        const uint64_t c = part_count[tid][b - range[tid].first];
        for (uint64_t i = 0; i < c; ++i, ++pos) {
          const SequenceStep& s = part[tid][pos];
          anl_match mm = in->matches[st.batch_first[b] + s.seg];
          mm.selected = s.variant < 0 ? -1 : s.variant;
          ms->matches[o++] = mm;
        }
      }
    }
  });
  if (want_tags) {  // Match.tag / Match.seqnr of the chosen sequence (:2476-2492)
    ms->tag_first.assign(1, 0);
    for (unsigned tid = 0; tid < used; ++tid)
      for (const StepTags& stg : part_tags[tid]) {
        ms->tag.insert(ms->tag.end(), stg.tag.begin(), stg.tag.end());
        ms->seqnr.insert(ms->seqnr.end(), stg.seqnr.begin(), stg.seqnr.end());
        ms->tag_first.push_back(ms->tag.size());
      }
    if (ms->tag_first.size() != ms->matches.size() + 1) return fail(ANL_ERR_INVALID, "internal error: tag list out of step");
  }
  pt.lap("consolidate: assemble");
  *out = ms_owner.release();
  return ANL_OK;
}
anl_status anl_match_set_consolidate(const anl_match_set* in, const char* text, size_t len, const anl_search_params* params,
                                     anl_match_set** out) try {
  return consolidate_impl(nullptr, in, text, len, params, out);
} catch (...) {
  return on_exception();
}
anl_status anl_model_consolidate(const anl_model* m, const anl_match_set* in, const char* text, size_t len,
                                 const anl_search_params* params, anl_match_set** out) try {
  if (!m) return fail(ANL_ERR_INVALID, "null argument");
  if (!m->host.built) return fail(ANL_ERR_NOT_BUILT, "Model has not been built yet! Call build() before find_all_matches()");
  return consolidate_impl(&m->host, in, text, len, params, out);
} catch (...) {
  return on_exception();
}
uint64_t anl_match_set_tags(const anl_match_set* ms, uint64_t i, const uint16_t** tags, const uint8_t** seqnr) {
  if (tags) *tags = nullptr;
  if (seqnr) *seqnr = nullptr;
  if (!ms || ms->tag_first.empty() || i + 1 >= ms->tag_first.size()) return 0;
  const uint64_t a = ms->tag_first[i], b = ms->tag_first[i + 1];
  if (tags) *tags = ms->tag.data() + a;
  if (seqnr) *seqnr = ms->seqnr.data() + a;
  return b - a;
}
// ---- language model / context rules of the model ----------------------------------------------------------------------
anl_status anl_model_read_contextrules(anl_model* m, const char* filename) try {
  if (!m || !filename) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->host.read_contextrules(filename, &err)) return fail(ANL_ERR_IO, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_model_add_contextrule(anl_model* m, const char* pattern, float score, const char* const* tags, uint32_t n_tags,
                                     const char* const* tagoffsets, uint32_t n_tagoffsets) try {
  if (!m || !pattern || (n_tags && !tags) || (n_tagoffsets && !tagoffsets)) return fail(ANL_ERR_INVALID, "null argument");
  std::vector<std::string> tg, to;
  for (uint32_t i = 0; i < n_tags; ++i) tg.push_back(tags[i] ? tags[i] : "");
  for (uint32_t i = 0; i < n_tagoffsets; ++i) to.push_back(tagoffsets[i] ? tagoffsets[i] : "");
  std::string err;
  if (!m->host.add_contextrule(pattern, score, tg, to, &err)) return fail(ANL_ERR_IO, err);  // (the reference returns an io::Error)
  return ANL_OK;
} catch (...) {
  return on_exception();
}
int32_t anl_model_have_lm(const anl_model* m) { return m && m->host.built && m->host.have_lm() ? 1 : 0; }
uint64_t anl_model_ngram_count(const anl_model* m) { return m && m->host.built ? m->host.lm_ngrams : 0; }
uint32_t anl_model_contextrule_count(const anl_model* m) { return m ? (uint32_t)m->host.context_rules.size() : 0; }
uint32_t anl_model_tag_count(const anl_model* m) { return m ? (uint32_t)m->host.tags.size() : 0; }
const char* anl_model_tag_name(const anl_model* m, uint32_t i) { return m && i < m->host.tags.size() ? m->host.tags[i].c_str() : nullptr; }
void anl_debug_lm_score_tokens(const anl_model* m, const int64_t* tokens, uint64_t n, float* logprob, double* perplexity) {
  if (m && tokens && n && logprob && perplexity) m->host.lm_score_tokens(tokens, n, logprob, perplexity);
}

// ---- test hooks for the host-side producer (no model, no GPU needed) ------------------------------------------
int64_t anl_debug_find_boundaries(const char* text, size_t len, uint64_t* begin, uint64_t* end, int32_t* strength, size_t cap) {
  const std::string t(text ? text : "", len);
  const std::vector<Boundary>& b = find_boundaries(t);
  for (size_t i = 0; i < b.size() && i < cap; ++i) {
    begin[i] = b[i].begin;
    end[i] = b[i].end;
    strength[i] = b[i].strength;
  }
  return (int64_t)b.size();
}
int64_t anl_debug_segment_text(const char* text, size_t len, uint32_t max_ngram, uint64_t* begin, uint64_t* end, uint32_t* order,
                               uint32_t* batch, size_t cap) {
  return anl_debug_segment_text_device(-1, text, len, max_ngram, begin, end, order, batch, cap, nullptr, nullptr, nullptr, 0, nullptr);
}
int64_t anl_debug_segment_text_device(int32_t device, const char* text, size_t len, uint32_t max_ngram, uint64_t* begin, uint64_t* end,
                                      uint32_t* order, uint32_t* batch, size_t cap, uint64_t* bound_begin, uint64_t* bound_end,
                                      int32_t* bound_strength, size_t bound_cap, uint64_t* n_bounds) try {
  const std::string t(text ? text : "", len);
  Segmentation sg;
  std::string err;
  if (device >= 0) {
    sg.text_len = len;
    sg.max_ngram = max_ngram;
    if (!segment_text_device(device, t, max_ngram, &sg.st, &err, &sg.bounds, &sg.batches)) {
      fail(ANL_ERR_CUDA, err);
      return -1;
    }
  } else {
    segment_any(-1, t, max_ngram, &sg, &err);
  }
  const SegmentedText& st = sg.st;
  if (n_bounds) *n_bounds = sg.bounds.size();
  for (size_t i = 0; i < sg.bounds.size() && i < bound_cap; ++i) {
    if (bound_begin) bound_begin[i] = sg.bounds[i].begin;
    if (bound_end) bound_end[i] = sg.bounds[i].end;
    if (bound_strength) bound_strength[i] = sg.bounds[i].strength;
  }
  const size_t nb = st.batch_first.size() - 1;
  for (size_t b = 0; b < nb; ++b)
    for (uint64_t k = st.batch_first[b]; k < st.batch_first[b + 1] && k < cap; ++k) {
      begin[k] = st.segs[k].begin;
      end[k] = st.segs[k].end;
      order[k] = st.segs[k].n;
      batch[k] = (uint32_t)b;
    }
  return (int64_t)st.segs.size();
} catch (...) {
  on_exception();
  return -1;
}

// A match set as anl_find_all_matches assembles it, from caller-supplied variant lists (one per segment, in the
// producer's order) instead of GPU lookups: lets the CPU tests drive anl_match_set_consolidate.
anl_status anl_debug_match_set_build(const char* text, size_t len, uint32_t max_ngram, int32_t unicodeoffsets,
                                     const uint8_t* looked, const uint64_t* offsets, const anl_variant* variants, uint64_t nseg,
                                     anl_match_set** out) try {
  if (!out || (!text && len > 0) || (nseg && (!looked || !offsets))) return fail(ANL_ERR_INVALID, "null argument");
  const std::string t(text ? text : "", len);
  SegmentedText st;
  segment_text(t, max_ngram, &st);
  if (st.segs.size() != nseg) return fail(ANL_ERR_INVALID, "segment count differs from the producer's");
  std::vector<uint64_t> cpmap;
  if (unicodeoffsets) cpmap = byte_to_codepoint_map(t);
  std::unique_ptr<anl_match_set> ms_owner(new anl_match_set());
  anl_match_set* ms = ms_owner.get();
  ms->matches.resize(nseg);
  ms->variants->reserve((nseg ? offsets[nseg] : 0) + 1);
  ms->variants->resize(nseg ? offsets[nseg] : 0);
  if (nseg && offsets[nseg]) memcpy(ms->variants->data(), variants, (size_t)offsets[nseg] * sizeof(anl_variant));
  for (uint64_t k = 0; k < nseg; ++k) {
    const SegmentSpan& sp = st.segs[k];
    anl_match& mm = ms->matches[k];
    const uint64_t cnt = looked[k] ? offsets[k + 1] - offsets[k] : 0;
    mm.begin = unicodeoffsets ? cpmap[sp.begin] : sp.begin;
    mm.end = unicodeoffsets ? cpmap[sp.end] : sp.end;
    mm.n = sp.n;
    mm.n_variants = cnt;
    mm.selected = (looked[k] && cnt > 0) ? 0 : -1;
    mm.variants = looked[k] ? ms->variants->data() + offsets[k] : nullptr;
  }
  *out = ms_owner.release();
  return ANL_OK;
} catch (...) {
  return on_exception();
}

uint64_t anl_match_set_len(const anl_match_set* ms) { return ms ? ms->matches.size() : 0; }
anl_status anl_match_set_get(const anl_match_set* ms, uint64_t i, anl_match* out) try {
  if (!ms || !out || i >= ms->matches.size()) return fail(ANL_ERR_INVALID, "match index out of range");
  *out = ms->matches[i];
  return ANL_OK;
} catch (...) {
  return on_exception();
}
void anl_match_set_free(anl_match_set* ms) { delete ms; }
void anl_match_set_lookup_counts(const anl_match_set* ms, uint64_t* segment_lookups, uint64_t* distinct_strings) {
  if (segment_lookups) *segment_lookups = ms ? ms->logical_lookups : 0;
  if (distinct_strings) *distinct_strings = ms ? ms->distinct_lookups : 0;
}

// ---- device-resident path ---------------------------------------------------------------------------------
anl_status anl_device_batch_create(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n_queries,
                                   const anl_search_params* params, anl_device_batch** out) try {
  if (!m || !offsets || !params || !out) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  int status = ANL_OK;
  DeviceBatch* b = m->engine.create_batch(blob ? blob : "", offsets, n_queries, *params, true, true, &err, &status);
  if (!b) return fail(status ? status : ANL_ERR_CUDA, err);
  *out = new anl_device_batch{b};
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_device_batch_run(anl_model* m, anl_device_batch* b, void* stream) try {
  if (!m || !b) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->engine.run_batch(b->b, reinterpret_cast<cudaStream_t>(stream), &err)) return fail(ANL_ERR_CUDA, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_device_batch_timings(anl_model* m, anl_device_batch* b, float* probe_ms, float* score_ms, float* rescore_ms) try {
  if (!m || !b || !probe_ms || !score_ms) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  float st[7];
  if (!m->engine.timings(b->b, st, &err)) return fail(ANL_ERR_CUDA, err);
  *probe_ms = st[0] + st[1];
  *score_ms = st[2] + st[3];
  if (rescore_ms) *rescore_ms = st[4] + st[5] + st[6];
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_device_batch_stage_timings(anl_model* m, anl_device_batch* b, float* stage_ms) try {
  if (!m || !b || !stage_ms) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->engine.timings(b->b, stage_ms, &err)) return fail(ANL_ERR_CUDA, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_device_batch_fetch(anl_model* m, anl_device_batch* b, anl_result_set** out) try {
  if (!m || !b || !out) return fail(ANL_ERR_INVALID, "null argument");
  std::unique_ptr<anl_result_set> rs(new anl_result_set());
  std::string err;
  int status = ANL_OK;
  if (!m->engine.fetch_batch(b->b, &rs->rs, &err, &status)) return fail(status ? status : ANL_ERR_CUDA, err);
  *out = rs.release();
  return ANL_OK;
} catch (...) {
  return on_exception();
}
void anl_device_batch_free(anl_model* m, anl_device_batch* b) {
  if (!b) return;
  if (m) m->engine.free_batch(b->b);
  delete b;
}
anl_status anl_device_batch_counters(anl_model* m, anl_device_batch* b, anl_counters* out) try {
  if (!m || !b || !out) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->engine.counters(b->b, out, &err)) return fail(ANL_ERR_CUDA, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
// ---- lexicon-sharded mode --------------------------------------------------------------------------------
anl_status anl_shard_export_size(anl_model* m, anl_device_batch* b, uint64_t* n_records, uint32_t* max_per_query) try {
  if (!m || !b || !n_records) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  int status = ANL_OK;
  if (!m->engine.shard_export_size(b->b, n_records, max_per_query, &err, &status)) return fail(status ? status : ANL_ERR_CUDA, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_shard_export(anl_model* m, anl_device_batch* b, void* d_heads, void* d_records, void* d_gids, void* d_flags) try {
  if (!m || !b || !d_heads || !d_records || !d_gids || !d_flags) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->engine.shard_export(b->b, d_heads, d_records, d_gids, d_flags, &err)) return fail(ANL_ERR_CUDA, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_shard_merge(anl_model* m, anl_device_batch* b, uint32_t n_shards, const void* d_heads_all,
                           const void* d_records_all, const void* d_gids_all, const void* d_flags_all, uint64_t record_stride,
                           uint32_t max_survivors, anl_result_set** out) try {
  if (!m || !b || !d_heads_all || !d_records_all || !d_gids_all || !d_flags_all)
    return fail(ANL_ERR_INVALID, "null argument");
  std::unique_ptr<anl_result_set> rs(out ? new anl_result_set() : nullptr);
  std::string err;
  int status = ANL_OK;
  if (!m->engine.shard_merge(b->b, n_shards, d_heads_all, d_records_all, d_gids_all, d_flags_all, record_stride, max_survivors,
                             rs ? &rs->rs : nullptr, &err, &status))
    return fail(status ? status : ANL_ERR_CUDA, err);
  if (out) *out = rs.release();
  return ANL_OK;
} catch (...) {
  return on_exception();
}

anl_status anl_shard_comm_id(uint8_t id[ANL_SHARD_ID_BYTES]) try {
  if (!id) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!shard_unique_id(id, &err)) return fail(ANL_ERR_CUDA, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_shard_comm_init(anl_model* m, const uint8_t id[ANL_SHARD_ID_BYTES], int32_t rank, int32_t n_ranks) try {
  if (!m || !id) return fail(ANL_ERR_INVALID, "null argument");
  std::string err;
  if (!m->engine.shard_comm_init(id, rank, n_ranks, &err)) return fail(ANL_ERR_CUDA, err);
  return ANL_OK;
} catch (...) {
  return on_exception();
}
void anl_shard_comm_free(anl_model* m) {
  if (m) m->engine.shard_comm_free();
}
anl_status anl_shard_batch_step(anl_model* m, anl_device_batch* b, anl_shard_step_stats* stats, anl_result_set** out) try {
  if (!m || !b) return fail(ANL_ERR_INVALID, "null argument");
  std::unique_ptr<anl_result_set> rs(out ? new anl_result_set() : nullptr);
  std::string err;
  int status = ANL_OK;
  ShardStepStats st;
  if (!m->engine.shard_step(b->b, rs ? &rs->rs : nullptr, &st, &err, &status)) return fail(status ? status : ANL_ERR_CUDA, err);
  if (stats) {
    stats->score_ms = st.score_ms;
    stats->exchange_ms = st.exchange_ms;
    stats->merge_ms = st.merge_ms;
    stats->bytes_received = st.bytes_received;
    stats->records_local = st.records_local;
    stats->records_total = st.records_total;
  }
  if (out) *out = rs.release();
  return ANL_OK;
} catch (...) {
  return on_exception();
}
anl_status anl_shard_find_variants_batch(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n_queries,
                                         const anl_search_params* params, anl_result_set** out) try {
  if (!m || !offsets || !params || !out || (!blob && n_queries > 0)) return fail(ANL_ERR_INVALID, "null argument");
  if (!m->host.built || !m->engine.uploaded())
    return fail(ANL_ERR_NOT_BUILT, "Model has not been built yet! Call build() before find_variants()");
  std::string err;
  int status = ANL_OK;
  DeviceBatch* b = m->engine.create_batch(blob ? blob : "", offsets, n_queries, *params, false, true, &err, &status);
  if (!b) return fail(status ? status : ANL_ERR_CUDA, err);
  std::unique_ptr<anl_result_set> rs(new anl_result_set());
  const bool ok = m->engine.shard_step(b, &rs->rs, nullptr, &err, &status);
  m->engine.free_batch(b);
  if (!ok) return fail(status ? status : ANL_ERR_CUDA, err);
  *out = rs.release();
  return ANL_OK;
} catch (...) {
  return on_exception();
}

void anl_debug_index_digest(const anl_model* m, uint64_t* out, size_t cap) {
  if (m && out) m->host.index_digest(out, cap);
}

anl_status anl_model_index_stats(const anl_model* m, anl_index_stats* out) try {
  if (!m || !out) return fail(ANL_ERR_INVALID, "null argument");
  if (!m->host.built) return fail(ANL_ERR_NOT_BUILT, "model has not been built");
  const HostIndex& ix = m->host.index;
  memset(out, 0, sizeof *out);
  out->table_slots = ix.table.size();
  out->slot_bytes = sizeof(Slot);
  out->table_bytes = ix.table.size() * sizeof(Slot);
  out->table_keys = ix.table_keys;
  out->bloom_bytes = ix.bloom.size() * sizeof(uint64_t);
  out->postings = ix.post_ana.size();
  out->anagrams = ix.ana_key.size();
  out->instances = ix.inst_vocab.size();
  out->instance_bytes = ix.inst_rows.size();
  out->norm_stride = ix.norm_stride;
  out->mset_entries = ix.mset.size();
  out->mset_bytes = ix.mset.size() * sizeof(MsetEntry);
  out->max_key_bits = ix.max_key_bits;
  out->max_charcount = ix.max_charcount;
  out->active_classes = (uint32_t)ix.active_classes.size();
  out->sd = (uint32_t)ix.sd;
  return ANL_OK;
} catch (...) {
  return on_exception();
}

}  // extern "C"
