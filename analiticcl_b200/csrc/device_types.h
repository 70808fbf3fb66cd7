// device_types.h -- data layout shared by the host index builder and the CUDA kernels.
//
// Everything the kernels read is laid out here (see DESIGN.md "Data layout in HBM").
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

namespace anl {

// ---- anagram keys -----------------------------------------------------------------------------
// The reference's AnaValue is an unbounded UBig (src/types.rs:33).  On the device a key is a
// fixed-width 192-bit little-endian integer (three 64-bit limbs).  The host builder rejects a
// lexicon whose largest key needs more than 192 bits (nld needs 166, eng 102).  Query-side keys
// that overflow 192 bits cannot equal any indexed key and are skipped (exact).
struct Key192 {
  uint64_t w0, w1, w2;
};

// 64-bit mix of a 192-bit key (host only: partitions the anagrams of the lexicon-sharded mode).
__host__ __device__ __forceinline__ uint64_t hash_key(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t h = a * 0x9E3779B97F4A7C15ULL;
  h ^= (b + 0x7F4A7C159E3779B9ULL) * 0xC2B2AE3D27D4EB4FULL;
  h ^= (c + 0x165667B19E3779F9ULL) * 0xD6E8FEB86659FD93ULL;
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ULL;
  h ^= h >> 29;
  return h;
}

// ---- linear multiset fingerprints ---------------------------------------------------------------------
// The probe side never multiplies primes.  A multiset of symbols X is fingerprinted by
//     mhash(X) = sum over its symbols s of class_rnd(s)   (mod 2^64)
// which is LINEAR: mhash(F - deleted + inserted) = mhash(F) - sum rnd(deleted) + sum rnd(inserted), so every
// node of a query's neighbourhood costs one 64-bit add instead of a 192-bit multiply plus a hash.  The
// fingerprint only routes a probe to postings; each posting is then verified EXACTLY against the
// anagram's own 192-bit prime-product key (X * p_x == key(C), X's key computed lazily for the few nodes
// that reach that stage), so fingerprint collisions cost a wasted compare, never a wrong result.
__host__ __device__ __forceinline__ uint64_t class_rnd(uint32_t s) {  // splitmix64 of the class index
  uint64_t z = (uint64_t)(s + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
// slot / Bloom word of a fingerprint: the low byte is skipped (a difference of two multisets whose symbol
// counts differ by even amounts only is even, so the lowest bits are the least uniform ones)
__host__ __device__ __forceinline__ uint64_t fp_index(uint64_t h, uint64_t mask) { return (h >> 8) & mask; }

// ---- the neighbour table ("symmetric delete, depth sd") -------------------------------------------
// Replaces both `index` (HashMap<AnaValue, AnaIndexNode>, src/index.rs:5) and the linear scan of
// `sortedindex[charcount]` (src/lib.rs:1268-1281).  Keys X of the table are
//     sd = 0:  every indexed anagram C                       (posting: C itself, cls = POST_SELF)
//     sd = 1:  additionally C / p_x for every class x in C   (posting: C, cls = x)
// A slot stores mhash(X) as a fingerprint, not X.
struct __attribute__((aligned(16))) Slot {
  uint64_t fp;        // mhash(X)
  uint32_t post_off;  // first posting (device copy, SLOT_INLINE set: the anagram rank of the slot's only posting)
  uint16_t post_cnt;  // number of postings; 0 = empty slot
  uint16_t pad;       // device copy: SLOT_INLINE | class of the only posting (most keys have one: no posting read)
};
static const uint8_t POST_SELF = 0xFF;
static const uint16_t SLOT_INLINE = 0x8000;
// Everything the exact stage needs about an anagram in one 32-byte sector (device only; built at upload from
// ana_key / ana_inst_off): with the table and the postings HBM-resident every separate array is another random
// DRAM access per verified posting.
struct __attribute__((aligned(32))) AnaRec {
  Key192 key;         // exact prime-product key
  uint32_t inst_off;  // first gather id
  uint32_t inst_cnt;  // instances
};

// Blocked Bloom filter in front of the table: one 64-bit word per key (fp_index), BLOOM_BITS bits inside ONE 32-bit
// half of it (the half is bit 7 of the fingerprint, just below the word index's bits).  A probe is a 4-byte load and a
// 32-bit mask test: with all three bits in a 64-bit word the mask alone cost ~24 instructions per node on the 32-bit
// datapath (three 64-bit one-hot shifts), a tenth of the Bloom stage's instruction stream; here it is three SHL and a LOP3.
// A miss (the overwhelmingly common case) costs one load.  The builders (host_model.cpp, gpu_build.cu) OR the 64-bit form.
static const int BLOOM_BITS = 3;
__host__ __device__ __forceinline__ uint32_t bloom_half(uint64_t h) { return (uint32_t)(h >> 7) & 1u; }
__host__ __device__ __forceinline__ uint32_t bloom_mask32(uint64_t h) {
  // bits taken from the top of the fingerprint (the word index uses bits 8..)
  const uint32_t t = (uint32_t)(h >> 32);
  return (1u << (t >> 27)) | (1u << ((t >> 22) & 31u)) | (1u << ((t >> 17) & 31u));
}
__host__ __device__ __forceinline__ uint64_t bloom_mask(uint64_t h) { return (uint64_t)bloom_mask32(h) << (32u * bloom_half(h)); }
#ifdef __CUDACC__
// the probe: true = all three bits set (the key may be in the table)
__device__ __forceinline__ bool bloom_test(const uint64_t* __restrict__ bloom, uint64_t word_mask, uint64_t h) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(bloom) + (((h >> 8) & word_mask) * 2 + bloom_half(h)));
  const uint32_t m = bloom_mask32(h);
  return (w & m) == m;
}
#endif

// ---- insertion multisets --------------------------------------------------------------------------
// Entry t of the table is one multiset of j >= 1 inserted symbols drawn from the classes that
// occur in the lexicon ("active classes"), ordered by j (all j=1 first, then j=2, ...), so the
// multisets of size <= J are the prefix [0, mset_end[J]).
struct __attribute__((aligned(16))) MsetEntry {
  uint64_t hsum;   // sum of class_rnd over the inserted symbols (what the multiset adds to a fingerprint)
  uint8_t cls[6];  // the symbols (prime indices), ascending, padded with 0xFF
  uint8_t j;       // multiset size
  uint8_t maxcls;  // largest symbol
};

static const int COLEX_N = 32;           // queries up to this many symbols unrank deletion sets by table lookup
static const int ANL_MAX_K = 6;          // largest supported max_anagram_distance after thresholding
static const int ANL_MAX_SYMBOLS = 236;  // longest query / entry (symbols) the device path accepts (row number + 16 fits a byte)

// ---- confusables on the device ---------------------------------------------------------------------
// Confusable rescoring (src/lib.rs:1656-1663,1733-1756) needs sesdiff's edit script of the raw strings.
// Two device stages handle every pair whose text lies in the Basic Multilingual Plane (UTF-8 of at most three
// bytes per character) and has at most 64 characters:
//  (1) triage, in the score kernel: a `-[..]` / `+[..]` instruction can only match characters of the
//      input's / candidate's "middle" (what remains after stripping the common prefix and suffix at character
//      boundaries).  Pairs that fail this necessary condition are settled with weight 1.
//      A pair whose middles share no character has the script =[prefix]-[middle a]+[middle b]=[suffix]; when the
//      patterns are "simple" (insertions / deletions only, no anchors) its weight is computed on the spot.
//  (2) confusable kernel: the full edit script over UTF-16 code units + pattern matching, one pair per thread
//      (editscript_fixed.h), for the pairs that pass (1).
// Options with text outside the BMP are dropped from the table (they cannot match a BMP pair); a pattern with an
// instruction that has no option left is dropped.  Whatever the device cannot settle (text outside the BMP, very long
// strings, internal capacity) is left to the host post-pass.
struct ConfOpt {
  uint64_t lo, hi;     // ASCII characters the option needs: bit c of (lo | hi << 64)
  uint32_t text_off;   // the option's text inside DeviceIndex::conf_text (UTF-16 code units)
  uint32_t text_len;
  uint32_t nonascii;   // the option also needs characters beyond ASCII
  uint32_t pad;
};
struct ConfInstr {
  int8_t op;           // -1 deletion, +1 insertion, 0 identity
  uint8_t n_opts;
  uint16_t first_opt;
};
struct ConfPat {
  double weight;
  uint16_t first_instr, n_instr;
  uint8_t strictbegin, strictend;  // `^` / `$` anchors
  uint8_t pad[2];
};
// OutRec.vocab_id bit 31: this record's confusable weight is settled (already applied, or provably 1)
static const uint32_t OUT_SKIP_CONFUSABLES = 0x80000000u;
// OutHead.count bit 31: the device could not settle every record of this query -- the host finishes it
static const uint32_t HEAD_HOST_FINISH = 0x80000000u;
// one (query, record) pair queued for the confusable kernel
struct ConfWork {
  uint32_t rec;    // record index in the result pool
  uint32_t query;  // query row (index into the batch's raw text offsets)
  uint32_t cost;   // characters in the two middles: the kernel hands pairs of similar cost to the lanes of a warp
  uint32_t pad;
};

// ---- alphabet on the device (query normalisation, src/anahash.rs:50-80) --------------------------------
// Greedy matching in alphabet-file order: members that start with a given byte, in (line, member) order.
struct AlphaMember {
  uint8_t seqnr;      // class index = symbol
  uint8_t len;        // bytes of the member (<= 14)
  uint8_t bytes[14];
};
struct AlphaFirst {
  uint16_t first, count;  // members starting with this byte: alpha_members[first .. first + count)
};
// per-query result of the encode kernel
static const uint8_t ENC_OK = 0, ENC_TOO_LONG_EMPTY = 1, ENC_TOO_LONG_UNSUPPORTED = 2;

// ---- split probe path: staged nodes in a global queue between the Bloom stage and the exact stage --------
// One node X = D + I' of some query's neighbourhood that passed the Bloom filter (40 bytes).
struct __attribute__((aligned(8))) QEntry {
  uint64_t h;       // mhash(X)
  uint64_t dprod;   // product of the primes of the deleted symbols
  uint64_t dd;      // deleted symbols, packed (count in the top byte)
  uint32_t t;       // index of I' in the multiset table (unused when isz == 0)
  uint32_t qi;      // query (position in the launch)
  uint8_t isz, imax;
  uint8_t pad[6];
};
// what the exact stage needs to know about a query
struct __attribute__((aligned(16))) QCtx {
  uint64_t kF[3];   // exact key of the focus
  uint32_t L;       // symbols of the query
  uint32_t ka_ok;   // max anagram distance | (key valid) << 8
};

// ---- per-model constant data ------------------------------------------------------------------------
struct DeviceIndex {
  const Slot* table;
  uint64_t table_mask;  // slots - 1 (power of two)
  const uint64_t* bloom;
  uint64_t bloom_mask;  // words - 1 (power of two)
  const uint32_t* post_ana;  // posting -> anagram rank
  const uint8_t* post_cls;   // posting -> class x (or POST_SELF)
  int32_t sd;                // symmetric-delete depth of the table (0 or 1)
  // anagrams in ascending key order (rank = position): key, first gather id, instance count
  const AnaRec* ana_rec;
  // instances in gather order: (anagram key ascending, vocab id ascending) == the order in which
  // gather_instances (src/lib.rs:1327-1391) visits them, so "gather id ascending" reproduces the
  // reference's stable-sort tie order.
  const uint8_t* inst_rows;  // [instances][norm_stride]: byte0 = length, byte1 = flags, bytes 2.. = symbols
  uint32_t norm_stride;      // multiple of 16
  const uint32_t* inst_vocab;  // vocab id per gather id
  const uint32_t* inst_freq;   // VocabValue.frequency per gather id
  const uint32_t* inst_gid;    // lexicon-sharded index only: global gather id per local gather id (else null)
  // confusable prefilter (null / 0 when the model has no confusables)
  const uint8_t* vocab_text;        // raw UTF-8 text of every vocabulary entry, back to back
  const uint32_t* vocab_text_off;   // vocab id -> offset (n_vocab + 1 entries)
  const ConfPat* conf_pats;
  const ConfInstr* conf_instrs;
  const ConfOpt* conf_opts;
  const uint16_t* conf_text;        // option texts of the patterns (UTF-16 code units), back to back
  uint32_t n_conf_pats;
  int32_t conf_prefilter;           // 1: table valid, the kernel may set OUT_SKIP_CONFUSABLES
  int32_t conf_all_simple;          // 1: every pattern consists of insertions / deletions only and has no anchor
  // alphabet tables for the encode kernel (device_encode = 0: the host normalises the queries)
  const AlphaMember* alpha_members;
  const AlphaFirst* alpha_first;    // [256]
  const uint32_t* lower_ranges;     // [n_lower_ranges][2] inclusive code point ranges of char::is_lowercase
  uint32_t n_lower_ranges;
  uint32_t unk_symbol;
  int32_t device_encode;
  const MsetEntry* mset;
  uint32_t mset_end[ANL_MAX_K + 1];  // mset_end[J] = number of entries with j <= J; [0] = 0
  const uint32_t* binom;             // [256][8] saturating binomials C(n, k)
  // colex unranking tables for the deletion neighbourhood of queries up to COLEX_N symbols: entry t of
  // colex2 / colex3 packs the positions p0 < p1 (< p2) of the t-th 2- / 3-subset in colex order, one per byte
  const uint32_t* colex2;            // C(COLEX_N, 2) entries
  const uint32_t* colex3;            // C(COLEX_N, 3) entries
  uint32_t prime_of[256];            // prime of a symbol (prime index), 0 if unused
  uint64_t charcount_mask[4];        // bit cc set iff some anagram has that charcount (cc < 256)
  uint32_t max_charcount;
  uint32_t max_len;  // longest indexed entry in symbols
  uint32_t n_anagrams;
  uint32_t n_instances;
  int32_t have_freq;
};

// row flags
static const uint8_t ROW_FIRST_LOWER = 1;  // first char of the raw text is_lowercase (src/lib.rs:1367-1374)

// ---- per-batch parameters --------------------------------------------------------------------------
struct Threshold {
  int32_t kind;  // 0 ratio, 1 ratio with limit, 2 absolute
  float ratio;
  uint32_t value;
};

struct BatchParams {
  Threshold max_anagram, max_edit;
  uint32_t max_matches;  // 0 = unlimited
  double score_threshold, cutoff_threshold;
  double w_ld, w_lcs, w_prefix, w_suffix, w_case, w_sum;
  double freq_weight64;  // f32 freq_weight widened to f64 (src/types.rs:339)
  int32_t freq_weight_nonzero, freq_weight_positive;
  int32_t stop_at_exact;
  int32_t finish_mode;    // FINISH_*
  uint32_t hit_cap;       // per-query capacity of the hit list (gather ids)
  uint32_t pool_cap;      // capacity of the packed result pool (records, whole launch)
  uint32_t query_stride;  // bytes per encoded query row (multiple of 16): len, flags, symbols
};
static const int FINISH_FULL = 0;       // rank, crop, cutoff on device (no confusables)
static const int FINISH_CROP = 1;       // rank + crop on device; late confusables + cutoff follow on the host
static const int FINISH_GATHER = 2;     // emit all survivors in gather order (early confusables on the host)
static const int FINISH_SHARD = 3;      // lexicon-sharded mode: emit all survivors unranked; merge_kernel finishes

// query flags
static const uint8_t Q_FIRST_LOWER = 1;

// Result record written by the score/rank kernel into the packed result pool.  The frequency is
// the raw VocabValue.frequency (or 1 when the model has no frequencies); the host divides by the
// query's max_freq (same IEEE division as src/lib.rs:1523, so the bits are identical).
struct __attribute__((aligned(16))) OutRec {
  double dist_score;
  uint32_t vocab_id;
  uint32_t freq;
};
// per-query result header
struct __attribute__((aligned(16))) OutHead {
  double max_freq;    // max over all instances within the edit distance (src/lib.rs:1460)
  uint32_t offset;    // first record in the pool
  uint32_t count;     // number of records
};

// per-query status bits
static const uint32_t QF_EMPTY = 1;         // empty query
static const uint32_t QF_HIT_OVERFLOW = 2;  // more instance hits than hit_cap: rerun with larger cap
static const uint32_t QF_OUT_OVERFLOW = 4;  // the packed result pool was exhausted: rerun the score kernel with a larger pool
static const uint32_t QF_PREFILTERED = 16;   // prefilter_kernel compacted this query's hit list (and counted its pairs)
static const uint32_t QF_UNSUPPORTED = 8;   // thresholded anagram distance > ANL_MAX_K or enumeration too large

// ---- export of the final result arrays (export.cu) ---------------------------------------------------------
static const uint32_t EXPORT_TILE = 1024;  // queries per CTA of the export stage
// what the host needs to know about a finished pass before it touches any per-query array
struct ExportSummary {
  uint32_t total;          // records in the exported arrays
  uint32_t n_rerun;        // queries whose hit list overflowed (they hold no records yet)
  uint32_t n_host_finish;  // queries the device could not finish (HEAD_HOST_FINISH)
  uint32_t pad;
};

struct Counters {
  unsigned long long deletion_keys, probes, filter_pass, table_steps, postings, anagram_hits, instance_pairs, dl_pairs,
      dl_cells, survivors, results, dp_pairs, dp_cells;
};

}  // namespace anl
