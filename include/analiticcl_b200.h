/* analiticcl_b200.h -- C ABI of the B200-native variant-lookup path.
 *
 * Drop-in boundary for the hot path of proycon/analiticcl v0.4.9: `VariantModel::find_variants`
 * (src/lib.rs:972) and its batching callers (`find_variants_par`, bindings/python/src/lib.rs:720;
 * `find_all_matches`, src/lib.rs:1790).  The reference has no FFI seam of its own; the seam is the
 * public methods of `VariantModel`, so every entry point below names the reference method it
 * replaces.  INTEGRATION.md shows the Rust `-sys` binding and the Python ctypes binding.
 *
 * Conventions: plain C types only, opaque handles, inputs borrowed for the duration of the call,
 * outputs owned by the library until the matching *_free.  Every function that can fail returns
 * an `anl_status` (0 = ok); `anl_last_error()` gives the message of the calling thread's last
 * failure.  There is NO CPU fallback: if CUDA is unavailable, anl_model_build() fails.
 */
#ifndef ANALITICCL_B200_H
#define ANALITICCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t anl_status;
enum {
  ANL_OK = 0,
  ANL_ERR_INVALID = 1,     /* bad argument; also: out of memory / internal C++ exception (never propagated to the caller) */
  ANL_ERR_IO = 2,          /* file could not be read / parsed (reference: io::Error) */
  ANL_ERR_NOT_BUILT = 3,   /* lookup before build() (reference: stderr + empty result, src/lib.rs:973) */
  ANL_ERR_CUDA = 4,        /* CUDA runtime / kernel failure, or no device */
  ANL_ERR_UNSUPPORTED = 5, /* outside the documented limits of the GPU path */
  ANL_ERR_EMPTY_INPUT = 6  /* empty query (reference: assert!(input_length > 0), src/lib.rs:1420) */
};

typedef struct anl_model anl_model;           /* VariantModel, src/lib.rs:50-100 */
typedef struct anl_result_set anl_result_set; /* Vec<Vec<VariantResult>> */
typedef struct anl_match_set anl_match_set;   /* Vec<Match>, src/search.rs:42-68 */

/* Weights, src/types.rs:39-73 (defaults .5 / .125 x4) */
typedef struct anl_weights {
  double ld, lcs, prefix, suffix, case_;
} anl_weights;

/* DistanceThreshold, src/types.rs:75-83 */
enum { ANL_THRESHOLD_RATIO = 0, ANL_THRESHOLD_RATIO_WITH_LIMIT = 1, ANL_THRESHOLD_ABSOLUTE = 2 };
typedef struct anl_distance_threshold {
  int32_t kind;
  float ratio;    /* Ratio, RatioWithLimit */
  uint32_t value; /* Absolute value, or the limit of RatioWithLimit (u8 range) */
} anl_distance_threshold;

/* StopCriterion, src/types.rs:307-313 */
enum { ANL_STOP_EXHAUSTIVE = 0, ANL_STOP_AT_EXACT_MATCH = 1 };

/* SearchParameters, src/types.rs:110-168, field for field (the sequence fields -- max_seq and the three
 * weights -- are read by anl_model_consolidate; lm_order, context_weight and single_thread are carried unread, as in the reference). */
typedef struct anl_search_params {
  anl_distance_threshold max_anagram_distance;
  anl_distance_threshold max_edit_distance;
  uint64_t max_matches;
  double score_threshold;
  double cutoff_threshold;
  int32_t stop_criterion;
  uint32_t max_ngram;
  uint32_t lm_order;
  uint64_t max_seq;
  int32_t single_thread;
  float context_weight;
  float variantmodel_weight;
  float lm_weight;
  float contextrules_weight;
  float freq_weight;
  int32_t consolidate_matches;
  int32_t unicodeoffsets;
} anl_search_params;

/* VocabType bit flags (src/vocab.rs:31-49) and FrequencyHandling (src/vocab.rs:100-106) */
enum { ANL_VOCAB_NONE = 0, ANL_VOCAB_INDEXED = 1, ANL_VOCAB_LM = 2, ANL_VOCAB_TRANSPARENT = 4 };
enum { ANL_FREQ_SUM = 0, ANL_FREQ_MAX = 1, ANL_FREQ_MIN = 2, ANL_FREQ_REPLACE = 3 };

/* VocabParams, src/vocab.rs:108-131 */
typedef struct anl_vocab_params {
  uint32_t text_column;
  int32_t freq_column; /* -1 = None */
  int32_t freq_handling;
  uint32_t vocab_type;
  uint32_t index; /* lexicon index; overwritten by read_vocabulary like the reference does */
} anl_vocab_params;

/* VariantResult, src/types.rs:326-332 */
#define ANL_NO_VIA UINT64_MAX
typedef struct anl_variant {
  uint64_t vocab_id;
  double dist_score;
  double freq_score;
  uint64_t via; /* ANL_NO_VIA = None; else the vocabulary id of the variant this result was reached through (variant lists) */
} anl_variant;

/* VocabValue, src/vocab.rs:7-29 (read-only view) */
typedef struct anl_vocab_info {
  const char* text; /* UTF-8, NUL terminated, owned by the model */
  uint32_t text_len;
  uint32_t frequency;
  uint32_t lexindex;
  uint32_t vocabtype;
  uint32_t tokencount;
  uint32_t norm_len;
} anl_vocab_info;

/* Defaults: Weights::default(), SearchParameters::default(), VocabParams::default() */
void anl_weights_default(anl_weights* w);
void anl_search_params_default(anl_search_params* p);
void anl_vocab_params_default(anl_vocab_params* p);

const char* anl_last_error(void);
const char* anl_version(void);
/* Number of CUDA kernels this library has launched in this process so far (all models). */
uint64_t anl_kernel_launches(void);

/* ---- model construction ------------------------------------------------------------------ */
/* VariantModel::new (src/lib.rs:104): alphabet TSV file + weights. */
anl_status anl_model_new(const char* alphabet_file, const anl_weights* weights, int32_t debug, anl_model** out);
/* VariantModel::new_with_alphabet (src/lib.rs:132): alphabet given as TSV text in memory. */
anl_status anl_model_new_from_tsv(const char* alphabet_tsv, size_t len, const anl_weights* weights, int32_t debug,
                                  anl_model** out);
void anl_model_free(anl_model* m);

/* read_vocabulary (src/lib.rs:519) */
anl_status anl_model_read_vocabulary(anl_model* m, const char* filename, const anl_vocab_params* params);
/* add_variant (src/lib.rs:460-514): adds `text` to the vocabulary (with `params`; set ANL_VOCAB_TRANSPARENT for an
 * error list whose entries should only lead to their reference) and links it to the existing entry `ref_id` with
 * `score`.  *added = 0 when the variant is the reference itself.  Results of later lookups are expanded
 * (expand_variants, src/lib.rs:1677-1727): a matched variant also yields its reference with dist_score * score and
 * via = the variant's id. */
anl_status anl_model_add_variant(anl_model* m, uint64_t ref_id, const char* text, size_t len, double score, int32_t has_frequency,
                                 uint32_t frequency, const anl_vocab_params* params, int32_t* added);
/* read_variants (src/lib.rs:766-897): TSV weighted variant list, `reference (variant score)*` or
 * `reference freq (variant score freq)*` (auto-detected). */
anl_status anl_model_read_variants(anl_model* m, const char* filename, const anl_vocab_params* params, int32_t transparent);
/* add_to_vocabulary (src/lib.rs:900); has_frequency=0 means None */
anl_status anl_model_add_to_vocabulary(anl_model* m, const char* text, size_t len, int32_t has_frequency,
                                       uint32_t frequency, const anl_vocab_params* params, uint64_t* vocab_id);
/* learn_variants (src/lib.rs:1062-1139): looks the inputs up (strict = 1: every input as a whole, find_variants -- one
 * batched GPU call over all inputs; strict = 0: every input as running text, the selected variants of find_all_matches)
 * and stores what was found in the model instead of returning it: an input gains a frequency count (a new input becomes
 * a TRANSPARENT entry) and is linked as a variant of every result, weighted by its distance score.  *count = links
 * added.  auto_build = 1 rebuilds the index and uploads it to the model's devices again (the reference's `build()`).
 * Inputs: UTF-8 blob + offsets[n_inputs + 1], as for anl_find_variants_batch. */
anl_status anl_model_learn_variants(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n_inputs,
                                    const anl_search_params* params, int32_t strict, int32_t auto_build, uint64_t* count);
/* Test hooks: the bookkeeping half of learn_variants on explicit (input text, result vocabulary id, dist_score)
 * triples (host only), and the variant links of an entry (kind 0: VariantOf targets + scores, 1: ReferenceFor ids;
 * returns their number, -1 for an unknown id). */
anl_status anl_debug_learn_apply(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n, const uint64_t* vocab_ids,
                                 const double* scores, uint64_t* count);
int64_t anl_debug_vocab_links(const anl_model* m, uint64_t id, int32_t kind, uint64_t* ids, double* scores, size_t cap);
/* read_confusablelist (src/lib.rs:414), add_to_confusables (src/lib.rs:444) */
anl_status anl_model_read_confusablelist(anl_model* m, const char* filename);
anl_status anl_model_add_to_confusables(anl_model* m, const char* editscript, double weight);
/* set_confusables_before_pruning (src/lib.rs:157) */
void anl_model_set_confusables_before_pruning(anl_model* m);

/* build (src/lib.rs:192): anagram index on the host, then upload to `device` (-1 = current). */
anl_status anl_model_build(anl_model* m, int32_t device);
/* build with the index construction itself forced onto the device (build_on_device = 1: anagram values, sorts,
 * postings, Bloom filter and table as kernels, csrc/gpu_build.cu) or onto the host cores (0).  anl_model_build chooses:
 * the device for lexicons of a million entries and more (ANL_GPU_BUILD=0/1 overrides).  Both builds produce the same
 * index arrays; only the slot order inside the open-addressing table differs (both are valid probe layouts). */
anl_status anl_model_build_on(anl_model* m, int32_t device, int32_t build_on_device);
/* build with one replica of the index on each of `n_devices` CUDA devices of this process.  Lookups
 * (anl_find_variants_batch, anl_find_all_matches) then spread every call over all of them: the batch is cut into
 * chunks that go round-robin to the devices, each device is driven by its own host thread, and the results come back
 * in query order as from one device.  This is the single-process counterpart of the reference's rayon loop over the
 * queries (src/bin/analiticcl.rs:418-482, process_par): one VariantModel, all GPUs of the box.  The device-batch and
 * lexicon-sharded entry points below work on the first listed device. */
anl_status anl_model_build_multi(anl_model* m, const int32_t* devices, uint32_t n_devices);
/* number of devices that hold a replica of the built index (0 before build) */
uint32_t anl_model_device_count(const anl_model* m);
/* Persistence of the built index (the reference has no on-disk index -- its build takes seconds,
 * the 10 M-entry lexicon of BASELINE config 5 takes 24 s here).  save_index writes the host copy of the index of a
 * built model; load_index replaces anl_model_build for a model that holds the SAME alphabet and vocabulary in the
 * same order (checked by a fingerprint; also checked: library data layout, array consistency): it reads the
 * arrays and uploads them to `device`.  ANL_ERR_IO on a missing, foreign, mismatching or corrupt file. */
anl_status anl_model_save_index(const anl_model* m, const char* filename);
anl_status anl_model_load_index(anl_model* m, const char* filename, int32_t device);
/* Which lexicon shard the built / loaded index holds (0 of 1 = the whole lexicon; lexicon-sharded mode below). */
void anl_model_shard(const anl_model* m, uint32_t* shard, uint32_t* n_shards);

/* ---- introspection ------------------------------------------------------------------------- */
int32_t anl_model_has(const anl_model* m, const char* text, size_t len);          /* has(), src/lib.rs:331 */
int64_t anl_model_vocab_id(const anl_model* m, const char* text, size_t len);     /* encoder lookup, -1 if absent */
uint64_t anl_model_vocab_size(const anl_model* m);                                 /* decoder.len() */
anl_status anl_model_get_vocab(const anl_model* m, uint64_t vocab_id, anl_vocab_info* out); /* get_vocab(), :341 */
uint32_t anl_model_lexicon_count(const anl_model* m);
const char* anl_model_lexicon_name(const anl_model* m, uint32_t index);
uint32_t anl_model_alphabet_size(const anl_model* m); /* alphabet_size(), src/lib.rs:163 (incl. UNK) */
uint64_t anl_model_index_size(const anl_model* m);    /* number of anagrams */
uint64_t anl_model_instance_count(const anl_model* m);
uint64_t anl_model_anagram_count_of_length(const anl_model* m, uint32_t charcount); /* sortedindex[cc].len() */
uint32_t anl_model_max_key_bits(const anl_model* m);
/* anahash / normalize_to_alphabet (src/anahash.rs:16,50).  anl_anahash writes the value as little
 * endian 64-bit limbs; returns the number of limbs needed (may exceed cap). */
int64_t anl_normalize(const anl_model* m, const char* text, size_t len, uint8_t* out, size_t cap);
int64_t anl_anahash(const anl_model* m, const char* text, size_t len, uint64_t* limbs, size_t cap);

/* sesdiff::shortest_edit_script(src, dst, false, false, false) as the reference calls it for confusable
 * rescoring (src/lib.rs:1736), rendered in sesdiff's text form `=[..]-[..]+[..]`.  Returns the number of
 * bytes needed (excluding the NUL); writes at most cap-1 bytes + NUL. */
int64_t anl_shortest_edit_script(const char* src, size_t src_len, const char* dst, size_t dst_len, char* out, size_t cap);
/* The same script computed by the fixed-capacity implementation that the confusable kernel runs per
 * thread (csrc/editscript_fixed.h), compiled for the host: a test hook.  Returns -1 when the pair is
 * outside that implementation's limits (non-ASCII text, more than 64 characters, internal capacity). */
int64_t anl_shortest_edit_script_fixed(const char* src, size_t src_len, const char* dst, size_t dst_len, char* out, size_t cap);
/* Confusable::found_in (src/confusables.rs:47): 1 if `pattern` occurs in the edit script of src -> dst,
 * 0 if not, -1 if the pattern does not parse. */
int32_t anl_confusable_found_in(const char* pattern, const char* src, size_t src_len, const char* dst, size_t dst_len);

/* ---- lookup ---------------------------------------------------------------------------------- */
/* find_variants over a batch (== find_variants_par, bindings/python/src/lib.rs:720).
 * `blob` holds the UTF-8 queries back to back, query i = blob[offsets[i] .. offsets[i+1]).
 * Results are final: ranked, cropped, confusable-rescored, cut off -- identical to what
 * VariantModel::find_variants returns for each query.  Failures are per query, never per batch (the reference
 * answers every query on its own, src/lib.rs:972-1027): a query the path cannot answer yields an empty list and
 * a flag bit in anl_result_set_flags():
 *   ANL_QUERY_EMPTY        empty query (the reference panics on it, src/lib.rs:1420)
 *   ANL_QUERY_UNSUPPORTED  outside the limits of the GPU path (DESIGN.md "Limits"): longer than 236 symbols while
 *                          an indexed entry could still be within reach, thresholded anagram distance above 6, or
 *                          a deletion neighbourhood beyond 2^31 nodes.  (A query longer than the longest indexed
 *                          entry + max_anagram_distance has an empty result by construction: no flag.) */
enum { ANL_QUERY_EMPTY = 1, ANL_QUERY_UNSUPPORTED = 2 };
anl_status anl_find_variants_batch(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n_queries,
                                   const anl_search_params* params, anl_result_set** out);
uint64_t anl_result_set_len(const anl_result_set* rs);
/* variants of query i: pointer to a contiguous array + count */
const anl_variant* anl_result_set_get(const anl_result_set* rs, uint64_t i, uint64_t* count);
/* whole CSR: offsets[n+1] and the flat variant array */
const uint64_t* anl_result_set_offsets(const anl_result_set* rs);
const anl_variant* anl_result_set_variants(const anl_result_set* rs);
uint32_t anl_result_set_flags(const anl_result_set* rs, uint64_t i);
void anl_result_set_free(anl_result_set* rs);

/* find_all_matches (src/lib.rs:1790).  Host segmentation (boundaries, n-gram spans, redundant-match
 * pruning) feeding the batched GPU lookup.  The call returns the producer's view: every segment of every
 * order with its variant list and `selected` = 0 where it has variants.  With max_ngram == 1 (and no
 * LM/context rules) that is the reference's result; with max_ngram > 1 the reference goes on to pick one
 * segmentation per hard-delimited batch (most_likely_sequence, src/lib.rs:1912-1924, 2088-2495) --
 * anl_match_set_consolidate below does that as a host post-pass over this match set. */
anl_status anl_find_all_matches(anl_model* m, const char* text, size_t len, const anl_search_params* params,
                                anl_match_set** out);
typedef struct anl_match {
  uint64_t begin, end; /* byte offsets, or code points if params.unicodeoffsets */
  uint32_t n;          /* n-gram order of this segment */
  int32_t selected;    /* index of the selected variant, -1 = none */
  uint64_t n_variants;
  const anl_variant* variants; /* NULL when the segment was skipped as redundant */
} anl_match;
uint64_t anl_match_set_len(const anl_match_set* ms);
anl_status anl_match_set_get(const anl_match_set* ms, uint64_t i, anl_match* out);
void anl_match_set_free(anl_match_set* ms);
/* How many segment lookups the call performed (== what the reference's per-segment find_variants loop
 * would issue, src/lib.rs:1883-1899) and how many distinct strings were actually sent to the GPU (the
 * producer de-duplicates identical segments inside a window). */
void anl_match_set_lookup_counts(const anl_match_set* ms, uint64_t* segment_lookups, uint64_t* distinct_strings);

/* most_likely_sequence (src/lib.rs:2088-2495) for a model without language model and context rules: per
 * hard-delimited batch, the lowest-cost path through the lattice of looked-up segments (cost of a variant =
 * tokens covered + 1 - score, f32; unigram without variants = copied from the input at cost 2; fail-safe
 * epsilon at cost 100).  `in` must be the match set anl_find_all_matches returned for the same text and
 * params->max_ngram; `*out` is a new, independent match set holding only the matches on the best path, in
 * text order, with `selected` = the chosen variant (-1 = out of vocabulary).  With max_ngram == 1 the result is
 * a copy of `in` (the reference skips the FST, :1929-1932).  This entry point knows no model, hence no language
 * model and no context rules (anl_model_consolidate below); tie-breaking among equal-cost paths is documented in
 * DESIGN.md. */
anl_status anl_match_set_consolidate(const anl_match_set* in, const char* text, size_t len, const anl_search_params* params,
                                     anl_match_set** out);
/* The whole of most_likely_sequence (src/lib.rs:2088-2495) with the model's language model and context rules: the
 * params->max_seq shortest paths per batch, each scored by lm_score (bigram model over the output tokens and the
 * boundaries between them, :2570-2674) and test_context_rules (:2501-2566), the three normalised terms weighted by
 * params->lm_weight / variantmodel_weight / contextrules_weight (:2381-2425).  Runs also for max_ngram == 1 when the
 * model holds a language model or context rules (:1912).  Matches of the result carry the tags their context rules
 * assign (anl_match_set_tags). */
anl_status anl_model_consolidate(const anl_model* m, const anl_match_set* in, const char* text, size_t len,
                                 const anl_search_params* params, anl_match_set** out);
/* Match.tag / Match.seqnr (src/search.rs:57-60) of match i: returns how many tags it has; *tags = indices into the model's
 * tag names (anl_model_tag_name), *seqnr = position of the match inside each tagged sequence. */
uint64_t anl_match_set_tags(const anl_match_set* ms, uint64_t i, const uint16_t** tags, const uint8_t** seqnr);

/* Language model: entries read with vocab_type = ANL_VOCAB_LM (read_lm of the Python binding = read_vocabulary with that
 * type, bindings/python/src/lib.rs:659-667) are n-grams "w1 w2 .." with their frequency as count; build() collects
 * them (src/lib.rs:246-295). */
int32_t anl_model_have_lm(const anl_model* m);
uint64_t anl_model_ngram_count(const anl_model* m);
/* read_contextrules (src/lib.rs:570-656): TSV `pattern <TAB> score [<TAB> tag;tag.. [<TAB> begin:length;..]]`;
 * add_contextrule (src/lib.rs:658-765): pattern = ';'-separated positions, each a word of the vocabulary, `?` (any),
 * `^` (in no lexicon), `@lexicon`, `!x` / `!(x|y)` (negation) or `x|y` (disjunction).  Call after the lexicons the
 * rules refer to are loaded. */
anl_status anl_model_read_contextrules(anl_model* m, const char* filename);
anl_status anl_model_add_contextrule(anl_model* m, const char* pattern, float score, const char* const* tags, uint32_t n_tags,
                                     const char* const* tagoffsets, uint32_t n_tagoffsets);
uint32_t anl_model_contextrule_count(const anl_model* m);
uint32_t anl_model_tag_count(const anl_model* m);
const char* anl_model_tag_name(const anl_model* m, uint32_t i);
/* Test hook: lm_score_tokens (src/lib.rs:2643-2674) on explicit vocabulary ids (-1 = out of vocabulary). */
void anl_debug_lm_score_tokens(const anl_model* m, const int64_t* tokens, uint64_t n, float* logprob, double* perplexity);

/* Test hooks for the host-side batch producer of find_all_matches (no model or GPU needed): the boundaries
 * (src/search.rs:190-258; strength 1 weak, 2 normal, 3 hard) and the n-gram segments per hard-delimited batch
 * (src/search.rs:262-312, src/lib.rs:1822-1903) exactly as anl_find_all_matches produces them.  Both return
 * the number of items (may exceed cap; only cap items are written). */
int64_t anl_debug_find_boundaries(const char* text, size_t len, uint64_t* begin, uint64_t* end, int32_t* strength, size_t cap);
int64_t anl_debug_segment_text(const char* text, size_t len, uint32_t max_ngram, uint64_t* begin, uint64_t* end, uint32_t* order,
                               uint32_t* batch, size_t cap);
/* The same producer run as kernels on CUDA device `device` (what anl_find_all_matches uses for running text; device < 0
 * = the host loop): segments as above plus, optionally, the boundaries they were cut from (at most bound_cap are
 * written, *n_bounds = how many there are).  Returns the number of segments, -1 on a CUDA error (anl_last_error). */
int64_t anl_debug_segment_text_device(int32_t device, const char* text, size_t len, uint32_t max_ngram, uint64_t* begin, uint64_t* end,
                                      uint32_t* order, uint32_t* batch, size_t cap, uint64_t* bound_begin, uint64_t* bound_end,
                                      int32_t* bound_strength, size_t bound_cap, uint64_t* n_bounds);

/* Test hook: a match set exactly as anl_find_all_matches assembles it, but from caller-supplied variant lists
 * (CSR `offsets[nseg + 1]` into `variants`, `looked[k]` = segment k was looked up; segments in the order of
 * anl_debug_segment_text) instead of GPU lookups -- drives anl_match_set_consolidate on a CPU-only box. */
anl_status anl_debug_match_set_build(const char* text, size_t len, uint32_t max_ngram, int32_t unicodeoffsets,
                                     const uint8_t* looked, const uint64_t* offsets, const anl_variant* variants, uint64_t nseg,
                                     anl_match_set** out);

/* ---- device-resident path (what bench.py times as `value`; plumbing for multi-GPU hosts) ------ */
typedef struct anl_device_batch anl_device_batch; /* encoded queries + result buffers in HBM */
/* Encodes on the host (alphabet normalisation) and uploads; buffers are sized for n_queries. */
anl_status anl_device_batch_create(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n_queries,
                                   const anl_search_params* params, anl_device_batch** out);
/* One pass of the hot path over the resident batch: probe kernels + score/rank kernels (+ confusable and
 * finish kernels when confusables are loaded) + the export kernels that leave the final result arrays (u64 offsets,
 * 32-byte variant records in query order, per-query flags) in HBM, on the model's stream.  `stream` is a cudaStream_t (0 = the model's own stream).  Does not synchronise. */
anl_status anl_device_batch_run(anl_model* m, anl_device_batch* b, void* stream);
/* Device-side cudaEvent timings (ms) of the kernels, averaged over the runs since the previous call
 * (events recorded on the launching stream); synchronises.  rescore_ms (may be NULL) = confusable
 * kernel + finish kernel, which only run when confusables are loaded. */
anl_status anl_device_batch_timings(anl_model* m, anl_device_batch* b, float* probe_ms, float* score_ms, float* rescore_ms);
/* The same window per kernel: stage_ms[7] = Bloom-stage kernel, exact-stage kernel (the whole fused probe
 * kernel when the split path is off), prefilter kernel, score/rank kernel launches, confusable kernel,
 * finish kernel, export kernels (final result arrays in query order).  Resets the window like
 * anl_device_batch_timings. */
anl_status anl_device_batch_stage_timings(anl_model* m, anl_device_batch* b, float* stage_ms);
/* Downloads the results of the last run and finishes them on the host (same output as
 * anl_find_variants_batch). */
anl_status anl_device_batch_fetch(anl_model* m, anl_device_batch* b, anl_result_set** out);
void anl_device_batch_free(anl_model* m, anl_device_batch* b);

/* Work counters of the last run (exact, counted inside the kernels). */
typedef struct anl_counters {
  uint64_t queries;
  uint64_t deletion_keys; /* distinct deletion-neighbourhood keys D (incl. the focus)        */
  uint64_t probes;        /* neighbourhood nodes tested against the Bloom filter (one 8-byte word each) */
  uint64_t filter_pass;   /* nodes that passed the filter and were looked up exactly          */
  uint64_t probe_steps;   /* 16-byte table slots read by the exact lookups (linear probing)   */
  uint64_t postings;      /* postings verified against the anagram keys                       */
  uint64_t anagram_hits;  /* indexed anagrams found (each exactly once per query)             */
  uint64_t instance_pairs;/* (query, instance) pairs handed to the distance kernel           */
  uint64_t dl_pairs;      /* pairs that pass the length pre-check (src/distance.rs:109-130)  */
  uint64_t dl_cells;      /* sum len_q*len_c over dl_pairs = the reference's matrix cell updates */
  uint64_t survivors;     /* pairs within max edit distance                                  */
  uint64_t results;       /* variants returned                                               */
  uint64_t reruns;        /* queries re-run because a fixed-capacity buffer overflowed       */
  uint64_t dp_pairs;      /* pairs the exact DP actually processed (after the bit-parallel OSA prefilter) */
  uint64_t dp_cells;      /* warp-cells of the exact DP: sum over batches of len_q * longest candidate * 32 lanes */
} anl_counters;
anl_status anl_device_batch_counters(anl_model* m, anl_device_batch* b, anl_counters* out);

/* ---- lexicon-sharded mode (mode 2): used only when the index is split over GPUs ------
 * Each rank builds the model with the same vocabulary and anl_model_build_sharded(); anagram keys are
 * partitioned by hash(key) mod n_shards.  Every rank runs the whole query batch against its shard
 * (anl_device_batch_create / _run as usual), exports its per-query survivors, the host framework
 * all-gathers the exports over NCCL (torch.distributed in analiticcl_b200/sharded.py), and
 * anl_shard_merge() ranks the union with the GLOBAL max frequency -- results identical to the
 * unsharded model.  All d_* arguments are device pointers on the model's device. */
anl_status anl_model_build_sharded(anl_model* m, int32_t device, uint32_t shard, uint32_t n_shards);
/* After anl_device_batch_run: number of survivor records this shard exports, and the largest
 * per-query survivor count. */
anl_status anl_shard_export_size(anl_model* m, anl_device_batch* b, uint64_t* n_records, uint32_t* max_per_query);
/* Copies the export into caller buffers: d_heads [n] x 16 B {f64 max_freq, u32 offset, u32 count},
 * d_records [n_records] x 16 B {f64 dist_score, u32 vocab_id, u32 frequency}, d_gids [n_records] x u32
 * (global gather ids = tie-break order), d_flags [n] x u32. */
anl_status anl_shard_export(anl_model* m, anl_device_batch* b, void* d_heads, void* d_records, void* d_gids, void* d_flags);
/* d_*_all hold the exports of all shards back to back: heads/flags with stride n, records/gids with
 * stride record_stride.  max_survivors >= the largest per-query survivor count summed over shards.
 * out == NULL: merge on the device only and leave the ranked lists in the batch's result pool. */
anl_status anl_shard_merge(anl_model* m, anl_device_batch* b, uint32_t n_shards, const void* d_heads_all,
                           const void* d_records_all, const void* d_gids_all, const void* d_flags_all, uint64_t record_stride,
                           uint32_t max_survivors, anl_result_set** out);

/* The same mode with the exchange inside the library: an NCCL communicator over the ranks that hold the shards
 * (one process per GPU; NCCL is bound at run time, in a torch process it is the copy torch loaded).  Rank 0 obtains
 * an id with anl_shard_comm_id and hands it to the other ranks by any host-side means (torch.distributed broadcast,
 * MPI, a file); every rank then calls anl_shard_comm_init with its shard coordinates.  Per batch (all ranks, same
 * queries): score against the shard, one 8-byte all-gather of sizes, ONE grouped collective that moves every shard's
 * survivors with exact sizes over NVLink, merge kernel with the global max frequency, export stage. */
#define ANL_SHARD_ID_BYTES 128
anl_status anl_shard_comm_id(uint8_t id[ANL_SHARD_ID_BYTES]);
anl_status anl_shard_comm_init(anl_model* m, const uint8_t id[ANL_SHARD_ID_BYTES], int32_t rank, int32_t n_ranks);
void anl_shard_comm_free(anl_model* m);
typedef struct anl_shard_step_stats {
  float score_ms, exchange_ms, merge_ms;                         /* CUDA events on the batch's stream */
  uint64_t bytes_received, records_local, records_total;         /* NVLink bytes this rank received; survivor records */
} anl_shard_step_stats;
/* One pass over a resident batch (anl_device_batch_create on a sharded model): score, exchange, merge.  stats and out
 * may be NULL (out == NULL: the merged, exported results stay on the device). */
anl_status anl_shard_batch_step(anl_model* m, anl_device_batch* b, anl_shard_step_stats* stats, anl_result_set** out);
/* anl_find_variants_batch for a sharded model: every rank passes the same queries and receives the full result. */
anl_status anl_shard_find_variants_batch(anl_model* m, const char* blob, const uint64_t* offsets, uint64_t n_queries,
                                         const anl_search_params* params, anl_result_set** out);

/* Test hook: out[0..11] = checksums of the index arrays in file order (anagram keys, instance offsets, charcounts,
 * vocabulary ids, frequencies, global gather ids, instance rows, table, Bloom words, posting anagrams, posting classes,
 * active classes); out[12] = order-independent digest of the occupied slots; out[13] = 1 iff every slot is reachable by
 * linear probing from its home position; out[14] = occupied slots; out[15] = digest of the scalar fields.  The host
 * build and the device build must agree on everything but out[7] (the slot order inside the table). */
void anl_debug_index_digest(const anl_model* m, uint64_t* out, size_t cap);

/* Size of the device-resident index (bytes per component) for roofline accounting. */
typedef struct anl_index_stats {
  uint64_t table_slots, table_bytes, slot_bytes, table_keys;
  uint64_t bloom_bytes, postings;
  uint64_t anagrams, instances;
  uint64_t instance_bytes, norm_stride;
  uint64_t mset_entries, mset_bytes;
  uint32_t max_key_bits, max_charcount, active_classes;
  uint32_t sd; /* symmetric-delete depth of the neighbour table */
} anl_index_stats;
anl_status anl_model_index_stats(const anl_model* m, anl_index_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* ANALITICCL_B200_H */
