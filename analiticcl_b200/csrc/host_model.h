// host_model.h -- host-side model of the variant-lookup path: alphabet, vocabulary, confusables and
// the builder that turns the lexicon into the device-resident index.
//
// Mirrors the public surface of the reference's VariantModel for this path (src/lib.rs:50-100,
// 104-165, 192-245, 331-343, 369-452, 519-568, 900-967) -- same names, argument meaning and error
// behaviour -- but none of its data structures: there is no HashMap<AnaValue,..> and no
// sortedindex here, only flat arrays laid out for the GPU (device_types.h).
#pragma once
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "device_types.h"
#include "hostpool.h"

namespace anl {

struct Weights {
  double ld = 0.5, lcs = 0.125, prefix = 0.125, suffix = 0.125, case_ = 0.125;  // src/types.rs:58-67
  double sum() const { return ld + lcs + prefix + suffix + case_; }             // src/types.rs:69-73
};

enum : uint32_t { VT_NONE = 0, VT_INDEXED = 1, VT_LM = 2, VT_TRANSPARENT = 4 };
enum : int32_t { FH_SUM = 0, FH_MAX = 1, FH_MIN = 2, FH_REPLACE = 3 };

struct VocabParams {  // src/vocab.rs:108-131
  uint32_t text_column = 0;
  int32_t freq_column = 1;  // -1 = None
  int32_t freq_handling = FH_MAX;
  uint32_t vocab_type = VT_INDEXED;
  uint32_t index = 0;
};

struct VocabEntry {  // src/vocab.rs:7-29
  std::string text;
  std::vector<uint8_t> syms;  // normalised form in prime-index space (UNK = alphabet length)
  uint32_t frequency = 1;
  uint32_t lexindex = 0;
  uint8_t tokencount = 1;
  uint8_t vocabtype = VT_NONE;
  bool first_lower = false;  // char::is_lowercase of the first char (src/lib.rs:1367-1374)
  bool ascii = false;        // text is pure ASCII (bytes == Unicode scalar values)
  // variants: Option<Vec<VariantReference>> (src/vocab.rs:22, :52-61)
  std::vector<std::pair<uint64_t, double>> variant_of;  // VariantOf(target id, score), insertion order
  std::vector<uint64_t> reference_for;                  // ReferenceFor(variant id)
  bool has_variants = false;                            // the Option is Some
};

// Greedy alphabet matcher (src/anahash.rs:16-80) with a first-byte dispatch table instead of
// the reference's nested scan over all alphabet lines at every position.
class Alphabet {
 public:
  void load_tsv(const std::string& text);  // src/lib.rs:369-407
  size_t size() const { return lines_.size(); }
  uint32_t unk_symbol() const { return (uint32_t)lines_.size(); }  // prime index of UNK (src/anahash.rs:42)
  // Appends one symbol (prime index) per matched alphabet member / unknown char.
  void encode(const char* s, size_t n, std::vector<uint8_t>* out) const;
  // Encode into a fixed buffer; returns the number of symbols (may exceed cap; extra symbols dropped).
  size_t encode_into(const char* s, size_t n, uint8_t* out, size_t cap) const;
  const std::vector<std::vector<std::string>>& lines() const { return lines_; }
  // Flat tables for the device encode kernel (device_types.h); false if a member does not fit.
  bool export_tables(std::vector<AlphaMember>* members, std::vector<AlphaFirst>* first) const;

 private:
  struct Member {
    uint32_t seqnr;
    std::string bytes;
  };
  void finalize();
  std::vector<std::vector<std::string>> lines_;
  std::vector<Member> by_first_[256];  // members starting with this byte, in (line, member) priority order
};

struct ConfusableInstr {
  int op;  // -1 deletion, 0 identity, +1 insertion
  std::vector<std::string> options;
  std::vector<std::u32string> options32;  // the same options as Unicode scalar values (prefilter)
};
struct Confusable {  // src/confusables.rs:5-11
  std::vector<ConfusableInstr> script;
  double weight = 1.0;
  bool strictbegin = false, strictend = false;
  bool simple = false;  // only insertions / deletions, no anchors: matching depends on edit chunks alone
};

// One context rule (src/search.rs:354-365).  The pattern (PatternMatch, :338-352) of every position is a prefix program:
// an operator followed by its operands (NOT: one expression, OR: n expressions).
struct RuleOp {
  uint8_t kind;     // sequence.cpp: any, no-lexicon, vocabulary id, from-lexicon, not, or
  uint8_t lexicon;  // from-lexicon: the lexicon's index
  uint16_t n;       // or: number of alternatives
  uint64_t vocab_id;
};
struct ContextRule {
  std::vector<RuleOp> code;
  std::vector<uint32_t> start;  // per pattern position: where its expression begins in `code`
  float score = 1.0f;           // > 1 bonus, < 1 penalty
  std::vector<uint16_t> tag;
  std::vector<std::pair<uint8_t, uint8_t>> tagoffset;  // begin, length (in pattern positions)
};
std::string trim_unicode(const std::string& s);  // str::trim (Unicode White_Space)

// Host copy of the built index (used for has(), statistics and to size device buffers).
struct HostIndex {
  std::vector<Key192> ana_key;         // ascending
  std::vector<uint32_t> ana_inst_off;  // n_anagrams + 1
  std::vector<uint16_t> ana_charcount;
  std::vector<uint32_t> inst_vocab;  // gather order
  std::vector<uint32_t> inst_freq;
  std::vector<uint32_t> inst_gid;    // lexicon-sharded index: global gather id per local one (empty = identity)
  uint32_t shard = 0, n_shards = 1;
  RawVec<uint8_t> inst_rows;         // (RawVec: sized without a fill, then first-touched on all cores)
  uint32_t norm_stride = 0;
  RawVec<Slot> table;
  RawVec<uint64_t> bloom;
  RawVec<uint32_t> post_ana;
  RawVec<uint8_t> post_cls;
  std::vector<uint8_t> active_classes;  // symbols that occur in indexed entries, ascending
  std::vector<MsetEntry> mset;
  uint32_t mset_end[ANL_MAX_K + 1] = {0};
  uint32_t mset_built_j = 0;
  uint32_t prime_of[256] = {0};
  uint64_t charcount_mask[4] = {0, 0, 0, 0};
  uint32_t max_charcount = 0, max_len = 0, max_key_bits = 0;
  uint64_t table_keys = 0;
  int sd = 1;
};

class HostModel {
 public:
  HostModel(const Weights& w, int debug) : weights(w), debug(debug) {}
  // -- construction (mirrors VariantModel) -----------------------------------------------------
  void init_vocab();  // src/vocab.rs:150-181
  bool read_alphabet_file(const std::string& filename, std::string* err);
  void read_alphabet_text(const std::string& tsv) { alphabet.load_tsv(tsv); }
  bool read_vocabulary(const std::string& filename, const VocabParams& p, std::string* err);
  uint64_t add_to_vocabulary(const char* text, size_t len, bool has_freq, uint32_t freq, const VocabParams& p);
  // weighted variant lists (src/lib.rs:460-514, 766-897); ref_id must exist
  bool add_variant(uint64_t ref_id, const char* text, size_t len, double score, bool has_freq, uint32_t freq, const VocabParams& p);
  bool add_variant_by_id(uint64_t ref_id, uint64_t variant_id, double score);
  bool read_variants(const std::string& filename, const VocabParams& p, bool transparent, std::string* err);
  // learn_variants (src/lib.rs:1062-1139), second half: stores (input, found variant) pairs; returns how many links were added
  struct LearnedVariant {
    std::string input;
    uint64_t vocab_id;
    double dist_score;
  };
  uint64_t learn_apply(const std::vector<LearnedVariant>& items);
  bool add_to_confusables(const std::string& editscript, double weight, std::string* err);
  bool read_confusablelist(const std::string& filename, std::string* err);
  // src/lib.rs:192-245: anagram values, grouping, ordering -> flat arrays (host side of build())
  // shard / n_shards: keep only the anagrams whose key hashes to this shard (lexicon-sharded mode)
  bool build_index(int sd, uint32_t shard, uint32_t n_shards, std::string* err);
  // builds the insertion-multiset table up to size J (idempotent)
  bool ensure_msets(uint32_t J, std::string* err);
  // Persistence of the built index (the reference has no on-disk index, its nearest format is the
  // TSV of `analiticcl index`, src/bin/analiticcl.rs:1190-1204): the flat arrays of HostIndex behind a header that
  // pins the library's struct layout and a fingerprint of the vocabulary the index was built from.  load_index
  // replaces build_index (same arrays, bit for bit) for a model that holds the same vocabulary in the same order.
  uint64_t vocabulary_fingerprint() const;
  void index_digest(uint64_t* out, size_t cap) const;  // test hook (see anl_debug_index_digest)
  // what build_index refuses about variant lists (also re-checked by load_index)
  bool check_variant_support(uint32_t n_shards, std::string* err) const;
  bool save_index(const std::string& path, std::string* err) const;
  bool load_index(const std::string& path, std::string* err);

  // -- queries against the host copy ---------------------------------------------------------------
  bool has(const char* text, size_t len) const;  // src/lib.rs:331-338
  int64_t vocab_id(const char* text, size_t len) const;
  // anahash as arbitrary-precision little-endian limbs (src/anahash.rs:16-47)
  std::vector<uint64_t> anahash_limbs(const char* text, size_t len) const;
  bool key_of(const uint8_t* syms, size_t n, Key192* out) const;  // false on 192-bit overflow
  uint32_t alphabet_size() const { return (uint32_t)((alphabet.size() + 1) & 0xFF); }  // src/lib.rs:163-165

  // -- host post-pass ---------------------------------------------------------------------------------
  double compute_confusable_weight(const char* input, size_t len, uint64_t candidate) const;  // src/lib.rs:1733-1756

  // -- language model and context rules of the sequence consolidation (sequence.cpp) ------------------
  void build_language_model();  // src/lib.rs:246-295, part of build()
  bool have_lm() const { return lm_ngrams > 0; }
  bool entry_tokens(uint64_t id, uint32_t* out, unsigned* n) const;  // into_ngram, src/lib.rs:2688-2751 (out holds 5)
  void lm_score_tokens(const int64_t* tokens, size_t n, float* logprob, double* perplexity) const;  // src/lib.rs:2643-2674
  bool add_contextrule(const std::string& pattern, float score, const std::vector<std::string>& tags,
                       const std::vector<std::string>& tagoffsets, std::string* err);  // src/lib.rs:658-765
  bool read_contextrules(const std::string& filename, std::string* err);               // src/lib.rs:570-656

  Weights weights;
  int debug;
  Alphabet alphabet;
  std::vector<VocabEntry> decoder;
  std::unordered_map<std::string, uint64_t> encoder;
  std::vector<std::string> lexicons;
  bool have_freq = false;
  bool any_variants = false;  // some entry holds variant references: results are expanded on the host (expand_variants)
  std::vector<Confusable> confusables;
  bool confusables_before_pruning = false;
  bool all_confusables_simple = true;
  bool built = false;
  HostIndex index;
  // n-gram counts keyed by vocabulary ids (unigram: id; bigram: first << 32 | second), src/lib.rs:75-79
  std::unordered_map<uint32_t, uint32_t> lm_unigram;
  std::unordered_map<uint64_t, uint32_t> lm_bigram;
  std::unordered_set<std::string> lm_higher;  // scratch of build_language_model
  uint64_t lm_ngrams = 0;
  std::vector<ContextRule> context_rules;  // src/lib.rs:82
  std::vector<std::string> tags;           // src/lib.rs:85
};

// sesdiff::shortest_edit_script(src, dst, false, false, false) -- see editscript.cpp
struct EditInstruction {
  int op;  // -1 deletion, 0 identity, +1 insertion
  std::string text;
};
struct EditView {  // non-owning form of an edit instruction
  int op;
  const char* p;
  size_t n;
};
bool confusable_found_in_views(const Confusable& c, const EditView* ref, size_t nref);
// allocation-free edit script for strings of at most 64 scalars (csrc/editscript_fixed.h over Unicode scalar
// values); out must hold 64 views; false = outside its limits
bool edit_views_fixed(const char* src, size_t src_len, const char* dst, size_t dst_len, EditView* out, size_t* nout);
std::vector<EditInstruction> shortest_edit_script(const std::string& src, const std::string& dst);
bool parse_confusable(const std::string& editscript, double weight, Confusable* out);
bool confusable_found_in(const Confusable& c, const std::vector<EditInstruction>& script);  // src/confusables.rs:47-128

extern const uint32_t kPrimes[168];  // src/types.rs:20-30

// The same build on the GPU (gpu_build.cu): fills hm->index like HostModel::build_index.
bool gpu_build_index(HostModel* hm, int sd, uint32_t shard, uint32_t n_shards, int device, std::string* err);

}  // namespace anl
