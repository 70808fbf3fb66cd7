// host_model.cpp -- alphabet, vocabulary and the index builder (host side of build()).
#include "host_model.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <type_traits>

#include "hostpool.h"
#include "unicode_tables.h"

namespace anl {

const uint32_t kPrimes[168] = {
    2,   3,   5,   7,   11,  13,  17,  19,  23,  29,  31,  37,  41,  43,  47,  53,  59,  61,  67,
    71,  73,  79,  83,  89,  97,  101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163,
    167, 173, 179, 181, 191, 193, 197, 199, 211, 223, 227, 229, 233, 239, 241, 251, 257, 263, 269,
    271, 277, 281, 283, 293, 307, 311, 313, 317, 331, 337, 347, 349, 353, 359, 367, 373, 379, 383,
    389, 397, 401, 409, 419, 421, 431, 433, 439, 443, 449, 457, 461, 463, 467, 479, 487, 491, 499,
    503, 509, 521, 523, 541, 547, 557, 563, 569, 571, 577, 587, 593, 599, 601, 607, 613, 617, 619,
    631, 641, 643, 647, 653, 659, 661, 673, 677, 683, 691, 701, 709, 719, 727, 733, 739, 743, 751,
    757, 761, 769, 773, 787, 797, 809, 811, 821, 823, 827, 829, 839, 853, 857, 859, 863, 877, 881,
    883, 887, 907, 911, 919, 929, 937, 941, 947, 953, 967, 971, 977, 983, 991, 997};

#if defined(__linux__)
}  // namespace anl
#include <sys/mman.h>
#include <unistd.h>
namespace anl {
static void advise_huge_pages(void* p, size_t bytes) {
  const uintptr_t two_mb = (uintptr_t)2 << 20;
  const uintptr_t lo = ((uintptr_t)p + two_mb - 1) & ~(two_mb - 1), hi = ((uintptr_t)p + bytes) & ~(two_mb - 1);
  if (hi > lo) madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
}
#else
static void advise_huge_pages(void*, size_t) {}
#endif
// First touch (zero fill) of a large, freshly sized RawVec on all cores: page faults are the cost of a fresh
// allocation and they scale with threads.
static void prefault(void* p, size_t bytes) {
  const size_t chunk = (size_t)2 << 20;
  parallel_ranges((bytes + chunk - 1) / chunk, 16, [&](unsigned, uint64_t lo, uint64_t hi) {
    const size_t b = (size_t)lo * chunk, e = std::min(bytes, (size_t)hi * chunk);
    if (e > b) memset((char*)p + b, 0, e - b);
  });
}

// Sort under a TOTAL order (no two elements compare equal) on all cores: equal chunks sorted in parallel, then
// merged pairwise in parallel rounds.  With a total order the result is the one std::sort gives.
template <class T, class Less>
static void parallel_sort(std::vector<T>& v, Less less) {
  const size_t n = v.size();
  unsigned parts = 1;
  while (parts * 2 <= host_threads() && n / (parts * 2) >= (1u << 15)) parts *= 2;
  if (parts == 1) {
    std::sort(v.begin(), v.end(), less);
    return;
  }
  std::vector<size_t> cut(parts + 1);
  for (unsigned i = 0; i <= parts; ++i) cut[i] = n / parts * i;
  cut[parts] = n;
  parallel_ranges(parts, 1, [&](unsigned, uint64_t lo, uint64_t hi) {
    for (uint64_t q = lo; q < hi; ++q) std::sort(v.begin() + cut[q], v.begin() + cut[q + 1], less);
  });
  std::vector<T> other(n);
  T* src = v.data();
  T* dst = other.data();
  for (unsigned width = 1; width < parts; width *= 2) {
    parallel_ranges(parts / (2 * width), 1, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t q = lo; q < hi; ++q) {
        const size_t a = cut[q * 2 * width], m = cut[q * 2 * width + width], e = cut[q * 2 * width + 2 * width];
        std::merge(src + a, src + m, src + m, src + e, dst + a, less);
      }
    });
    std::swap(src, dst);
  }
  if (src != v.data()) std::copy(src, src + n, v.data());
}

// ---- UTF-8 ----------------------------------------------------------------------------------------
static inline unsigned u8len(unsigned char c) {
  if (c < 0x80) return 1;
  if ((c & 0xE0) == 0xC0) return 2;
  if ((c & 0xF0) == 0xE0) return 3;
  if ((c & 0xF8) == 0xF0) return 4;
  return 1;
}
static inline uint32_t u8decode(const char* s, size_t avail, unsigned* len) {
  unsigned char c = (unsigned char)s[0];
  unsigned l = u8len(c);
  if (l > avail) l = (unsigned)avail;
  *len = l;
  if (l == 1) return c;
  uint32_t cp = c & (0xFFu >> (l + 1));
  for (unsigned i = 1; i < l; ++i) cp = (cp << 6) | ((unsigned char)s[i] & 0x3F);
  return cp;
}
static bool is_unicode_space(uint32_t c) {
  return c == ' ' || (c >= 9 && c <= 13) || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
         c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
std::string trim_unicode(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b) {
    unsigned l;
    uint32_t cp = u8decode(s.data() + a, b - a, &l);
    if (!is_unicode_space(cp)) break;
    a += l;
  }
  while (b > a) {
    size_t p = b - 1;
    while (p > a && ((unsigned char)s[p] & 0xC0) == 0x80) --p;
    unsigned l;
    uint32_t cp = u8decode(s.data() + p, b - p, &l);
    if (!is_unicode_space(cp)) break;
    b = p;
  }
  return s.substr(a, b - a);
}

// ---- Alphabet -------------------------------------------------------------------------------------
void Alphabet::load_tsv(const std::string& text) {
  size_t pos = 0;
  while (pos < text.size()) {
    size_t nl = text.find('\n', pos);
    size_t end = nl == std::string::npos ? text.size() : nl;
    std::string line = text.substr(pos, end - pos);
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (!line.empty()) {
      std::vector<std::string> fields;
      size_t fp = 0;
      for (;;) {
        size_t tab = line.find('\t', fp);
        std::string f = line.substr(fp, tab == std::string::npos ? std::string::npos : tab - fp);
        if (f == "\\s")
          fields.push_back(" ");
        else if (f == "\\t")
          fields.push_back("\t");
        else if (f == "\\n")
          fields.push_back("\n");
        else {
          std::string t = trim_unicode(f);
          if (!t.empty()) fields.push_back(t);
        }
        if (tab == std::string::npos) break;
        fp = tab + 1;
      }
      lines_.push_back(fields);  // kept even when it has no members, like the reference
    }
    pos = end + 1;
  }
  finalize();
}

void Alphabet::finalize() {
  for (auto& v : by_first_) v.clear();
  for (uint32_t seqnr = 0; seqnr < lines_.size(); ++seqnr)
    for (const std::string& m : lines_[seqnr])
      if (!m.empty()) by_first_[(unsigned char)m[0]].push_back(Member{seqnr, m});
}

bool Alphabet::export_tables(std::vector<AlphaMember>* members, std::vector<AlphaFirst>* first) const {
  members->clear();
  first->assign(256, AlphaFirst{0, 0});
  for (int b = 0; b < 256; ++b) {
    if (by_first_[b].size() > 0xFFFF || members->size() + by_first_[b].size() > 0xFFFF) return false;
    (*first)[b].first = (uint16_t)members->size();
    (*first)[b].count = (uint16_t)by_first_[b].size();
    for (const Member& m : by_first_[b]) {
      if (m.bytes.size() > 14 || m.seqnr > 255) return false;
      AlphaMember am;
      memset(&am, 0, sizeof am);
      am.seqnr = (uint8_t)m.seqnr;
      am.len = (uint8_t)m.bytes.size();
      memcpy(am.bytes, m.bytes.data(), m.bytes.size());
      members->push_back(am);
    }
  }
  return true;
}

size_t Alphabet::encode_into(const char* s, size_t n, uint8_t* out, size_t cap) const {
  size_t count = 0, i = 0;
  const uint32_t unk = unk_symbol();
  while (i < n) {
    const std::vector<Member>& cands = by_first_[(unsigned char)s[i]];
    const Member* hit = nullptr;
    for (const Member& m : cands) {
      if (i + m.bytes.size() <= n && memcmp(s + i, m.bytes.data(), m.bytes.size()) == 0) {
        hit = &m;
        break;
      }
    }
    if (hit) {
      if (count < cap) out[count] = (uint8_t)hit->seqnr;
      i += hit->bytes.size();
    } else {
      if (count < cap) out[count] = (uint8_t)unk;
      i += u8len((unsigned char)s[i]);
    }
    ++count;
  }
  return count;
}

void Alphabet::encode(const char* s, size_t n, std::vector<uint8_t>* out) const {
  uint8_t buf[512];
  size_t c = encode_into(s, n, buf, sizeof buf);
  if (c <= sizeof buf) {
    out->assign(buf, buf + c);
  } else {
    out->resize(c);
    encode_into(s, n, out->data(), c);
  }
}

// ---- HostModel: vocabulary -------------------------------------------------------------------------
void HostModel::init_vocab() {
  const char* names[3] = {"<bos>", "<eos>", "<unk>"};
  for (int i = 0; i < 3; ++i) {
    VocabEntry e;
    e.text = names[i];
    e.frequency = 0;
    e.tokencount = 1;
    e.vocabtype = VT_NONE;
    decoder.push_back(e);
    encoder[names[i]] = (uint64_t)i;
  }
}

bool HostModel::read_alphabet_file(const std::string& filename, std::string* err) {
  std::ifstream f(filename, std::ios::binary);
  if (!f) {
    *err = "cannot open alphabet file " + filename;
    return false;
  }
  std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  alphabet.load_tsv(text);
  return true;
}

uint64_t HostModel::add_to_vocabulary(const char* text, size_t len, bool has_freq, uint32_t freq, const VocabParams& p) {
  const uint32_t frequency = has_freq ? freq : 1;
  std::string key(text, len);
  auto it = encoder.find(key);
  if (it != encoder.end()) {
    VocabEntry& item = decoder[it->second];
    const uint32_t old_frequency = item.frequency;
    const uint8_t old_type = item.vocabtype;
    switch (p.freq_handling) {
      case FH_SUM: item.frequency += frequency; break;
      case FH_MAX: if (frequency > item.frequency) item.frequency = frequency; break;
      case FH_MIN: if (frequency < item.frequency) item.frequency = frequency; break;
      default: item.frequency = frequency; break;
    }
    if (it->second <= 2)
      item.vocabtype = VT_LM;
    else if ((item.vocabtype & VT_TRANSPARENT) && !(p.vocab_type & VT_TRANSPARENT))
      item.vocabtype ^= VT_TRANSPARENT;
    item.lexindex |= 1u << (p.index & 31);
    // the device index holds its own copy of the frequencies (the reference reads decoder[].frequency live in
    // score_and_rank): a changed entry invalidates the built index instead of ranking with stale values
    if (item.frequency != old_frequency || item.vocabtype != old_type) built = false;
    return it->second;
  }
  const uint64_t id = decoder.size();
  VocabEntry e;
  e.text = key;
  alphabet.encode(text, len, &e.syms);
  e.frequency = frequency;
  e.tokencount = (uint8_t)(std::count(key.begin(), key.end(), ' ') + 1);
  e.lexindex = 1u << (p.index & 31);
  e.vocabtype = (uint8_t)p.vocab_type;
  if (len > 0) {
    unsigned l;
    e.first_lower = anl_unicode::is_lowercase(u8decode(text, len, &l));
  }
  e.ascii = true;
  for (size_t i = 0; i < len; ++i) e.ascii = e.ascii && ((unsigned char)text[i] < 0x80);
  encoder.emplace(std::move(key), id);
  decoder.push_back(std::move(e));
  built = false;
  return id;
}

static bool parse_u32(const std::string& s, uint32_t* out) {  // str::parse::<u32>()
  size_t i = 0;
  if (i < s.size() && s[i] == '+') ++i;
  if (i >= s.size()) return false;
  uint64_t v = 0;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (uint64_t)(s[i] - '0');
    if (v > 0xFFFFFFFFull) return false;
  }
  *out = (uint32_t)v;
  return true;
}

bool HostModel::read_vocabulary(const std::string& filename, const VocabParams& params, std::string* err) {
  std::ifstream f(filename, std::ios::binary);
  if (!f) {
    *err = "cannot open vocabulary file " + filename;
    return false;
  }
  VocabParams p = params;
  p.index = (uint32_t)(lexicons.size() & 0xFF);
  std::string line;
  size_t linenr = 0;
  std::vector<std::pair<size_t, size_t>> fields;
  while (std::getline(f, line)) {
    ++linenr;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    fields.clear();
    size_t fp = 0;
    for (;;) {
      size_t tab = line.find('\t', fp);
      size_t end = tab == std::string::npos ? line.size() : tab;
      fields.emplace_back(fp, end - fp);
      if (tab == std::string::npos) break;
      fp = tab + 1;
    }
    if (p.text_column >= fields.size()) {
      *err = filename + ":" + std::to_string(linenr) + ": expected text column not found";
      return false;
    }
    uint32_t frequency = 1;
    if (p.freq_column >= 0) {
      if (p.vocab_type & VT_INDEXED) have_freq = true;
      if ((size_t)p.freq_column < fields.size()) {
        std::string fs = line.substr(fields[p.freq_column].first, fields[p.freq_column].second);
        if (!parse_u32(fs, &frequency)) {
          *err = filename + ":" + std::to_string(linenr) + ": frequency should be a valid integer";
          return false;
        }
      }
    }
    add_to_vocabulary(line.data() + fields[p.text_column].first, fields[p.text_column].second, true, frequency, p);
  }
  lexicons.push_back(filename);
  return true;
}

// src/lib.rs:460-514: the variant is added to the vocabulary and linked with its reference in both directions.
bool HostModel::add_variant(uint64_t ref_id, const char* text, size_t len, double score, bool has_freq, uint32_t freq,
                            const VocabParams& p) {
  return add_variant_by_id(ref_id, add_to_vocabulary(text, len, has_freq, freq, p), score);
}

// src/lib.rs:1106-1130, the bookkeeping half of learn_variants: every (input text, found variant) pair, in order.  An
// input that is already in the vocabulary gains one occurrence per consecutive run of pairs; a new one is added as a
// TRANSPARENT entry (that type alone: it is linked to its reference but -- like in the reference -- not indexed); the
// input then becomes a variant of what was found for it, weighted by the distance score.
uint64_t HostModel::learn_apply(const std::vector<LearnedVariant>& items) {
  uint64_t count = 0;
  VocabParams p;
  p.vocab_type = VT_TRANSPARENT;
  p.freq_handling = FH_MAX;
  const LearnedVariant* prev = nullptr;
  for (const LearnedVariant& it : items) {
    uint64_t id;
    auto e = encoder.find(it.input);
    if (e != encoder.end()) {
      id = e->second;
      if (!prev || prev->input != it.input) {
        decoder[id].frequency += 1;
        built = false;  // (the device index holds its own copy of the frequencies)
      }
    } else {
      id = add_to_vocabulary(it.input.data(), it.input.size(), true, 1, p);
    }
    if (it.vocab_id != id && it.vocab_id < decoder.size() && add_variant_by_id(it.vocab_id, id, it.dist_score)) ++count;
    prev = &it;
  }
  return count;
}

bool HostModel::add_variant_by_id(uint64_t ref_id, uint64_t vid, double score) {  // src/lib.rs:478-514
  if (vid == ref_id) return false;
  VocabEntry& ref = decoder[ref_id];
  ref.has_variants = true;  // (only the first mention of a variant counts)
  if (std::find(ref.reference_for.begin(), ref.reference_for.end(), vid) == ref.reference_for.end()) ref.reference_for.push_back(vid);
  VocabEntry& var = decoder[vid];
  // The reference's duplicate check on this side compares the stored target with the variant's own id (:505-508),
  // so repeating a (reference, variant) pair stores the link again; kept as is, it shows in the expanded results.
  bool blocked = false;
  if (var.has_variants)
    for (const auto& e : var.variant_of) blocked = blocked || e.first == vid;
  var.has_variants = true;
  if (!blocked) var.variant_of.emplace_back(ref_id, score);
  any_variants = true;
  built = false;
  return true;
}

// src/lib.rs:766-897.  TSV: reference (variant score)*, or with frequencies: reference freq (variant score freq)*;
// which of the two is detected from the first line whose column count fits and whose second column is an integer.
bool HostModel::read_variants(const std::string& filename, const VocabParams& p_in, bool transparent, std::string* err) {
  std::ifstream f(filename, std::ios::binary);
  if (!f) {
    *err = "cannot open variant list " + filename;
    return false;
  }
  VocabParams p = p_in;
  p.index = (uint8_t)(lexicons.size() & 0xFF);
  VocabParams pv = p;
  if (transparent) pv.vocab_type |= VT_TRANSPARENT;
  enum { UNDECIDED, WITH_FREQ, WITHOUT_FREQ } layout = UNDECIDED;
  std::string line;
  size_t lineno = 0;
  std::vector<std::pair<size_t, size_t>> col;
  auto field = [&](size_t k) { return std::string(line.data() + col[k].first, col[k].second); };
  auto bad = [&](const char* what) {
    *err = std::string(what) + " (line " + std::to_string(lineno) + " of " + filename + ")";
    return false;
  };
  while (std::getline(f, line)) {
    ++lineno;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    col.clear();
    for (size_t b = 0;;) {
      const size_t t = line.find('\t', b);
      col.emplace_back(b, (t == std::string::npos ? line.size() : t) - b);
      if (t == std::string::npos) break;
      b = t + 1;
    }
    bool have = false;
    uint32_t freq = 0;
    if (layout == UNDECIDED) {
      if (col.size() < 2) return bad("a variant list line needs at least a reference and one variant");
      if ((col.size() - 2) % 3 == 0) {
        if (parse_u32(field(1), &freq)) {
          layout = WITH_FREQ;
          have = true;
        }
      } else {
        layout = WITHOUT_FREQ;
      }
    } else if (layout == WITH_FREQ) {
      if (col.size() < 2 || !parse_u32(field(1), &freq)) return bad("frequency must be an integer");
      have = true;
    }
    const uint64_t ref_id = add_to_vocabulary(line.data() + col[0].first, col[0].second, have, freq, p);
    const size_t first = layout == WITH_FREQ ? 2 : 1, step = layout == WITH_FREQ ? 3 : 2;
    for (size_t k = first; k + step <= col.size(); k += step) {
      const std::string sc = field(k + 1);
      char* end = nullptr;
      const double score = strtod(sc.c_str(), &end);
      if (sc.empty() || *end != '\0') return bad("variant scores must be floating point values");
      uint32_t vf = 0;
      if (layout == WITH_FREQ && !parse_u32(field(k + 2), &vf)) return bad("variant frequency must be an integer");
      add_variant(ref_id, line.data() + col[k].first, col[k].second, score, layout == WITH_FREQ, vf, pv);
    }
  }
  lexicons.push_back(filename);
  return true;
}

bool HostModel::add_to_confusables(const std::string& editscript, double weight, std::string* err) {
  Confusable c;
  if (!parse_confusable(editscript, weight, &c)) {
    *err = "invalid confusable edit script: " + editscript;
    return false;
  }
  confusables.push_back(c);
  all_confusables_simple = all_confusables_simple && c.simple;
  return true;
}

bool HostModel::read_confusablelist(const std::string& filename, std::string* err) {  // src/lib.rs:414-441
  std::ifstream f(filename, std::ios::binary);
  if (!f) {
    *err = "cannot open confusable list " + filename;
    return false;
  }
  std::string line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    size_t tab = line.find('\t');
    double weight = 1.0;
    std::string script = line.substr(0, tab);
    if (tab != std::string::npos) {
      size_t tab2 = line.find('\t', tab + 1);
      std::string ws = line.substr(tab + 1, tab2 == std::string::npos ? std::string::npos : tab2 - tab - 1);
      char* end = nullptr;
      weight = strtod(ws.c_str(), &end);
      if (ws.empty() || *end != '\0') {
        *err = "confusable score should be a float: " + ws;
        return false;
      }
    }
    if (!add_to_confusables(script, weight, err)) return false;
  }
  return true;
}

// ---- keys ----------------------------------------------------------------------------------------------
static inline bool key_mul(Key192& k, uint64_t m) {
  unsigned __int128 t0 = (unsigned __int128)k.w0 * m;
  unsigned __int128 t1 = (unsigned __int128)k.w1 * m + (uint64_t)(t0 >> 64);
  unsigned __int128 t2 = (unsigned __int128)k.w2 * m + (uint64_t)(t1 >> 64);
  k.w0 = (uint64_t)t0;
  k.w1 = (uint64_t)t1;
  k.w2 = (uint64_t)t2;
  return (uint64_t)(t2 >> 64) == 0;
}
static inline bool key_less(const Key192& a, const Key192& b) {
  if (a.w2 != b.w2) return a.w2 < b.w2;
  if (a.w1 != b.w1) return a.w1 < b.w1;
  return a.w0 < b.w0;
}
static inline bool key_eq(const Key192& a, const Key192& b) { return a.w0 == b.w0 && a.w1 == b.w1 && a.w2 == b.w2; }
static inline unsigned key_bits(const Key192& k) {
  if (k.w2) return 128 + 64 - __builtin_clzll(k.w2);
  if (k.w1) return 64 + 64 - __builtin_clzll(k.w1);
  if (k.w0) return 64 - __builtin_clzll(k.w0);
  return 0;
}

bool HostModel::key_of(const uint8_t* syms, size_t n, Key192* out) const {
  Key192 k{1, 0, 0};
  for (size_t i = 0; i < n; ++i) {
    if (syms[i] >= 168) return false;
    if (!key_mul(k, kPrimes[syms[i]])) return false;
  }
  *out = k;
  return true;
}

std::vector<uint64_t> HostModel::anahash_limbs(const char* text, size_t len) const {
  std::vector<uint8_t> syms;
  alphabet.encode(text, len, &syms);
  std::vector<uint64_t> v{1};
  for (uint8_t s : syms) {
    uint64_t m = kPrimes[s < 168 ? s : 167], carry = 0;
    for (auto& limb : v) {
      unsigned __int128 t = (unsigned __int128)limb * m + carry;
      limb = (uint64_t)t;
      carry = (uint64_t)(t >> 64);
    }
    if (carry) v.push_back(carry);
  }
  return v;
}

bool HostModel::check_variant_support(uint32_t n_shards, std::string* err) const {
  if (!any_variants) return true;
  // Results are expanded on the host from the candidates that pass the score threshold.  The reference decides
  // *whether* to expand from every instance within the edit distance (src/lib.rs:1464); the two differ only for
  // a transparent entry that holds no variant reference (it is dropped iff the list is expanded at all).
  if (n_shards > 1) {
    *err = "variant lists are not supported in the lexicon-sharded mode";
    return false;
  }
  for (size_t id = 3; id < decoder.size(); ++id)
    if ((decoder[id].vocabtype & VT_TRANSPARENT) && !decoder[id].has_variants) {
      *err = "a transparent entry without variant references (" + decoder[id].text +
             ") next to variant lists is not supported by the GPU path";
      return false;
    }
  return true;
}

// ---- index build (src/lib.rs:192-245) ----------------------------------------------------------------------
bool HostModel::build_index(int sd, uint32_t shard, uint32_t n_shards, std::string* err) {
  PhaseTimer pt;
  HostIndex ix;
  ix.sd = sd;
  if (n_shards == 0 || shard >= n_shards) {
    *err = "invalid shard";
    return false;
  }
  ix.shard = shard;
  ix.n_shards = n_shards;
  if (!check_variant_support(n_shards, err)) return false;
  if (alphabet.size() + 1 > 168) {
    *err = "alphabet has more classes than there are primes (168)";
    return false;
  }
  for (uint32_t s = 0; s <= alphabet.size(); ++s) ix.prime_of[s] = kPrimes[s];

  struct Item {
    Key192 key;
    uint32_t id;
  };
  std::vector<Item> items;
  bool class_seen[256] = {false};
  {
    // keys on all cores, a range of vocabulary ids per thread; the lists are joined in id order, and the first
    // offending entry (lowest id) is the one reported, as a serial pass would
    struct Part {
      std::vector<Item> items;
      bool class_seen[256] = {false};
      uint32_t max_len = 0, max_key_bits = 0;
      size_t bad_id = SIZE_MAX;
      int bad_kind = 0;  // 1 = too long, 2 = key overflow
    };
    std::vector<Part> part(host_threads());
    const unsigned used = parallel_ranges(decoder.size(), 1u << 14, [&](unsigned tid, uint64_t lo, uint64_t hi) {
      Part& mine = part[tid];
      mine.items.reserve(hi - lo);
      for (size_t id = lo; id < hi; ++id) {
        const VocabEntry& v = decoder[id];
        if (!(v.vocabtype & VT_INDEXED)) continue;
        if (v.syms.empty()) {
          // reference: anahash of "" is 1 and the entry is indexed under it; it can never be returned
          // (deletions never reach the empty value, src/iterators.rs:177) except as an exact match of an
          // empty query, which the reference rejects (src/lib.rs:1420).  Not indexed here.
          continue;
        }
        if (v.syms.size() > (size_t)ANL_MAX_SYMBOLS) {
          mine.bad_id = id;
          mine.bad_kind = 1;
          return;
        }
        Item it;
        if (!key_of(v.syms.data(), v.syms.size(), &it.key)) {
          mine.bad_id = id;
          mine.bad_kind = 2;
          return;
        }
        it.id = (uint32_t)id;
        mine.items.push_back(it);
        for (uint8_t s : v.syms) mine.class_seen[s] = true;
        mine.max_len = std::max<uint32_t>(mine.max_len, (uint32_t)v.syms.size());
        mine.max_key_bits = std::max(mine.max_key_bits, key_bits(it.key));
      }
    });
    size_t total = 0;
    for (unsigned t = 0; t < used; ++t) {  // ranges ascend with t: the first failing part holds the lowest failing id
      if (part[t].bad_kind == 1) {
        *err = "lexicon entry longer than " + std::to_string(ANL_MAX_SYMBOLS) + " symbols: " + decoder[part[t].bad_id].text;
        return false;
      }
      if (part[t].bad_kind == 2) {
        *err = "anagram value of lexicon entry exceeds 192 bits: " + decoder[part[t].bad_id].text;
        return false;
      }
      total += part[t].items.size();
    }
    items.reserve(total);
    for (unsigned t = 0; t < used; ++t) {
      items.insert(items.end(), part[t].items.begin(), part[t].items.end());
      for (int c = 0; c < 256; ++c) class_seen[c] = class_seen[c] || part[t].class_seen[c];
      ix.max_len = std::max(ix.max_len, part[t].max_len);
      ix.max_key_bits = std::max(ix.max_key_bits, part[t].max_key_bits);
    }
  }
  if (items.empty()) {
    *err = "no indexed vocabulary entries";
    return false;
  }
  if (decoder.size() > 0xFFFFFFF0ull) {
    *err = "vocabulary too large";
    return false;
  }
  pt.lap("build: anagram keys");
  // instances in (key ascending, vocab id ascending) order = the reference's gather order
  parallel_sort(items, [](const Item& a, const Item& b) {
    if (!key_eq(a.key, b.key)) return key_less(a.key, b.key);
    return a.id < b.id;
  });
  pt.lap("build: sort instances");
  ix.norm_stride = ((ix.max_len + 2) + 15) & ~15u;
  // lexicon-sharded mode: anagrams are partitioned by hash(key) mod n_shards (instances follow their
  // key); the position in the global (key, vocab id) order stays the tie-break key on every shard
  std::vector<uint32_t> gids;
  if (n_shards > 1) {
    std::vector<Item> mine;
    for (size_t g = 0; g < items.size(); ++g) {
      const Key192& k = items[g].key;
      if (hash_key(k.w0, k.w1, k.w2) % n_shards == shard) {
        mine.push_back(items[g]);
        gids.push_back((uint32_t)g);
      }
    }
    items.swap(mine);
    if (items.empty()) {
      *err = "shard holds no anagrams";
      return false;
    }
    memset(class_seen, 0, sizeof class_seen);
    ix.max_len = 0;
    ix.max_key_bits = 0;
    for (const Item& it : items) {
      const VocabEntry& v = decoder[it.id];
      for (uint8_t s : v.syms) class_seen[s] = true;
      ix.max_len = std::max<uint32_t>(ix.max_len, (uint32_t)v.syms.size());
      ix.max_key_bits = std::max(ix.max_key_bits, key_bits(it.key));
    }
    ix.norm_stride = ((ix.max_len + 2) + 15) & ~15u;
    ix.inst_gid = gids;
  }
  {
    // Instance rows and the anagram arrays on all cores: a range of instances per thread; an anagram starts where
    // the key changes, so a thread first counts the anagrams that start in its range, the ranges' counts give
    // every thread its first anagram rank, and the second pass fills rows and anagram entries in place.
    const size_t ninst = items.size();
    ix.inst_rows.resize(ninst * ix.norm_stride);
    prefault(ix.inst_rows.data(), ninst * ix.norm_stride);
    ix.inst_vocab.resize(ninst);
    ix.inst_freq.resize(ninst);
    auto starts_anagram = [&](size_t g) { return g == 0 || !key_eq(items[g].key, items[g - 1].key); };
    const unsigned nt_max = host_threads();
    std::vector<uint64_t> first_rank(nt_max + 1, 0);
    std::vector<std::pair<uint64_t, uint64_t>> range(nt_max, {0, 0});
    const unsigned used = parallel_ranges(ninst, 1u << 14, [&](unsigned tid, uint64_t lo, uint64_t hi) {
      range[tid] = {lo, hi};
      uint64_t c = 0;
      for (size_t g = lo; g < hi; ++g) c += starts_anagram(g) ? 1 : 0;
      first_rank[tid + 1] = c;
    });
    for (unsigned t = 0; t < used; ++t) first_rank[t + 1] += first_rank[t];
    const size_t nana = first_rank[used];
    ix.ana_key.resize(nana);
    ix.ana_inst_off.resize(nana);
    ix.ana_charcount.resize(nana);
    struct Seen {
      uint64_t mask[4] = {0, 0, 0, 0};
      uint32_t max_cc = 0;
    };
    std::vector<Seen> seen(nt_max);
    parallel_ranges(used, 1, [&](unsigned, uint64_t tlo, uint64_t thi) {
      for (uint64_t t = tlo; t < thi; ++t) {
        uint64_t r = first_rank[t];
        for (size_t g = range[t].first; g < range[t].second; ++g) {
          const VocabEntry& v = decoder[items[g].id];
          if (starts_anagram(g)) {
            const uint32_t cc = (uint32_t)v.syms.size();
            ix.ana_key[r] = items[g].key;
            ix.ana_inst_off[r] = (uint32_t)g;
            ix.ana_charcount[r] = (uint16_t)cc;
            seen[t].mask[cc >> 6] |= 1ull << (cc & 63);
            seen[t].max_cc = std::max(seen[t].max_cc, cc);
            ++r;
          }
          uint8_t* row = ix.inst_rows.data() + g * ix.norm_stride;
          row[0] = (uint8_t)v.syms.size();
          row[1] = v.first_lower ? ROW_FIRST_LOWER : 0;
          memcpy(row + 2, v.syms.data(), v.syms.size());
          ix.inst_vocab[g] = items[g].id;
          ix.inst_freq[g] = v.frequency;
        }
      }
    });
    for (unsigned t = 0; t < used; ++t) {
      for (int w = 0; w < 4; ++w) ix.charcount_mask[w] |= seen[t].mask[w];
      ix.max_charcount = std::max(ix.max_charcount, seen[t].max_cc);
    }
  }
  ix.ana_inst_off.push_back((uint32_t)items.size());
  for (uint32_t s = 0; s < 256; ++s)
    if (class_seen[s]) ix.active_classes.push_back((uint8_t)s);

  pt.lap("build: instance rows");
  // neighbour table: self postings, plus (sd = 1) one posting per distinct class of every anagram
  // The table is keyed by the linear multiset fingerprint mhash (device_types.h), not by the prime
  // product: mhash(C / p_x) = mhash(C) - class_rnd(x).
  struct Post {
    uint64_t fp;
    uint32_t ana;
    uint8_t cls;
  };
  // Generated on all cores (one list per thread, a range of anagrams each), then brought into the total order
  // (fp, anagram, class).  fp is a hash, so its top byte splits the postings into 256 evenly filled buckets that
  // are already in order among themselves: every thread scatters its own list into the buckets (offsets from the
  // per-thread bucket counts), then the buckets are sorted on all cores.  The order is total and equal postings
  // are identical, so the result is the one a single std::sort over one list gives.
  RawVec<Post> posts;
  {
    auto post_less = [](const Post& a, const Post& b) {
      if (a.fp != b.fp) return a.fp < b.fp;
      if (a.ana != b.ana) return a.ana < b.ana;
      return a.cls < b.cls;
    };
    const unsigned nt_max = host_threads();
    std::vector<std::vector<Post>> part(nt_max);
    std::vector<std::vector<uint64_t>> count(nt_max, std::vector<uint64_t>(256, 0));
    const unsigned used = parallel_ranges(ix.ana_key.size(), 1u << 14, [&](unsigned tid, uint64_t lo, uint64_t hi) {
      std::vector<Post>& mine = part[tid];
      std::vector<uint64_t>& cnt = count[tid];
      mine.reserve((hi - lo) * (sd ? 10 : 1));
      for (uint64_t r = lo; r < hi; ++r) {
        // symbols of this anagram from its first instance
        const uint8_t* row = ix.inst_rows.data() + (size_t)ix.ana_inst_off[r] * ix.norm_stride;
        uint64_t h = 0;
        for (uint32_t i = 0; i < row[0]; ++i) h += class_rnd(row[2 + i]);
        mine.push_back(Post{h, (uint32_t)r, POST_SELF});
        ++cnt[h >> 56];
        if (sd >= 1) {
          bool seen[256] = {false};
          for (uint32_t i = 0; i < row[0]; ++i) {
            uint8_t x = row[2 + i];
            if (seen[x]) continue;
            seen[x] = true;
            if (row[0] == 1) continue;  // the empty value is never a node (no empty leaves, src/iterators.rs:177)
            const uint64_t hx = h - class_rnd(x);
            mine.push_back(Post{hx, (uint32_t)r, x});
            ++cnt[hx >> 56];
          }
        }
      }
    });
    pt.lap("build: postings");
    // bucket b holds [first[b], first[b + 1]); thread t writes its share of bucket b from offset[t][b] on
    std::vector<uint64_t> first(257, 0);
    for (unsigned b = 0; b < 256; ++b) {
      uint64_t at = first[b];
      for (unsigned t = 0; t < used; ++t) {
        const uint64_t c = count[t][b];
        count[t][b] = at;  // becomes the thread's write cursor for this bucket
        at += c;
      }
      first[b + 1] = at;
    }
    const uint64_t total = first[256];
    posts.resize(total);
    prefault(posts.data(), total * sizeof(Post));
    parallel_ranges(used, 1, [&](unsigned, uint64_t tlo, uint64_t thi) {
      for (uint64_t t = tlo; t < thi; ++t) {
        std::vector<uint64_t>& cursor = count[t];
        for (const Post& q : part[t]) posts[cursor[q.fp >> 56]++] = q;
        std::vector<Post>().swap(part[t]);
      }
    });
    parallel_ranges(256, 1, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t b = lo; b < hi; ++b) std::sort(posts.begin() + first[b], posts.begin() + first[b + 1], post_less);
    });
  }
  pt.lap("build: sort postings");
  uint64_t groups = 0;
  for (size_t i = 0; i < posts.size(); ++i)
    if (i == 0 || posts[i].fp != posts[i - 1].fp) ++groups;
  ix.table_keys = groups;
  uint64_t slots = 1;
  while (slots < groups * 2) slots <<= 1;
  uint64_t words = 1;
  while (words * 2 < groups) words <<= 1;  // 1..2 keys per 64-bit word, 3 bits each: false positives < 0.1 %
  // A filter that outgrows the 126 MB L2 turns every probe into a DRAM sector read.  Large indexes trade
  // false positives (a wasted exact lookup each) for residency: up to ~8 keys per word (false positives
  // ~2-3 %) while the filter is larger than 128 MB.
  uint64_t bloom_max_bytes = 128ull << 20;
  if (const char* e = getenv("ANL_BLOOM_MAX_MB")) bloom_max_bytes = (uint64_t)std::max(1, atoi(e)) << 20;  // (tests: the dense branch on a small lexicon)
  while (words * 8 > bloom_max_bytes && words * 8 >= groups) words >>= 1;
  // the table is written at random: ask for huge pages before the first touch (a 4 KB page per TLB entry makes
  // every insertion a page walk once the table is hundreds of MB; harmless where THP is unavailable)
  pt.lap("build: count keys");
  ix.table.resize(slots);  // (RawVec: no fill; the parallel first touch below zeroes them = empty slots / words)
  ix.bloom.resize(words);
  advise_huge_pages(ix.table.data(), slots * sizeof(Slot));
  advise_huge_pages(ix.bloom.data(), words * sizeof(uint64_t));
  prefault(ix.table.data(), slots * sizeof(Slot));
  prefault(ix.bloom.data(), words * sizeof(uint64_t));
  ix.post_ana.resize(posts.size());
  ix.post_cls.resize(posts.size());
  prefault(ix.post_ana.data(), posts.size() * sizeof(uint32_t));
  prefault(ix.post_cls.data(), posts.size());
  pt.lap("build: allocate table");
  // Everything that does not depend on the insertion order runs on all cores first: the posting arrays (a plain
  // copy) and the Bloom words (OR is commutative; atomic because ranges share words).
  parallel_ranges(posts.size(), 1u << 16, [&](unsigned, uint64_t lo, uint64_t hi) {
    for (uint64_t t = lo; t < hi; ++t) {
      ix.post_ana[t] = posts[t].ana;
      ix.post_cls[t] = posts[t].cls;
      if (t == 0 || posts[t].fp != posts[t - 1].fp) {
        const uint64_t fp = posts[t].fp;
        __atomic_fetch_or(&ix.bloom[fp_index(fp, words - 1)], bloom_mask(fp), __ATOMIC_RELAXED);
      }
    }
  });
  pt.lap("build: postings arrays + Bloom filter");
  // The table itself is filled in key order on one thread: with linear probing the slot of a colliding key depends
  // on who came first, and the layout is kept reproducible.
  for (size_t i = 0; i < posts.size();) {
    size_t j = i;
    while (j < posts.size() && posts[j].fp == posts[i].fp) ++j;
    if (j - i > 0xFFFF) {
      *err = "posting list too long";
      return false;
    }
    if (i + 160 < posts.size())  // the home slot of a key some hundred postings ahead: a cache miss
      __builtin_prefetch(&ix.table[fp_index(posts[i + 160].fp, slots - 1)], 1);
    const uint64_t fp = posts[i].fp;
    uint64_t idx = fp_index(fp, slots - 1);
    while (ix.table[idx].post_cnt != 0) idx = (idx + 1) & (slots - 1);
    ix.table[idx] = Slot{fp, (uint32_t)i, (uint16_t)(j - i), 0};
    i = j;
  }
  pt.lap("build: table");
  index = std::move(ix);
  build_language_model();
  built = true;
  return true;
}

// ---- persistence of the built index -------------------------------------------------------------------------------
namespace {
const char kIndexMagic[8] = {'A', 'N', 'L', 'I', 'D', 'X', '0', '3'};  // (03: Bloom bits of a key inside one 32-bit half of its word)
struct IndexFileHeader {  // fixed-size, little-endian hosts only (x86-64 / aarch64)
  char magic[8];
  uint32_t header_bytes, slot_bytes, key_bytes, max_k;
  uint64_t fingerprint;
  uint32_t shard, n_shards, norm_stride, max_charcount, max_len, max_key_bits;
  int32_t sd;
  uint32_t reserved;
  uint64_t table_keys;
  uint64_t checksum;  // content_checksum() of every array that follows the header
  uint64_t charcount_mask[4];
  uint32_t prime_of[256];
};
// 64-bit checksum of a byte range on all cores: fixed 4 MB pieces hashed independently (8 bytes per step), the
// piece hashes folded in piece order -- independent of the thread count.
uint64_t fp_mix(uint64_t h, uint64_t x);
uint64_t bytes_checksum(const void* data, size_t bytes) {
  const size_t piece = (size_t)4 << 20;
  const size_t np = (bytes + piece - 1) / piece;
  std::vector<uint64_t> ph(np, 0);
  parallel_ranges(np, 4, [&](unsigned, uint64_t lo, uint64_t hi) {
    for (uint64_t q = lo; q < hi; ++q) {
      const unsigned char* p = (const unsigned char*)data + (size_t)q * piece;
      size_t n = std::min(piece, bytes - (size_t)q * piece);
      uint64_t h0 = 0x9E3779B97F4A7C15ULL ^ n, h1 = 0xC2B2AE3D27D4EB4FULL, h2 = 0x165667B19E3779F9ULL, h3 = 0xD6E8FEB86659FD93ULL;
      while (n >= 32) {  // four independent lanes: the multiplies pipeline
        uint64_t w[4];
        memcpy(w, p, 32);
        h0 = (h0 ^ w[0]) * 0xFF51AFD7ED558CCDULL; h0 ^= h0 >> 29;
        h1 = (h1 ^ w[1]) * 0xC4CEB9FE1A85EC53ULL; h1 ^= h1 >> 31;
        h2 = (h2 ^ w[2]) * 0x9FB21C651E98DF25ULL; h2 ^= h2 >> 30;
        h3 = (h3 ^ w[3]) * 0xD6E8FEB86659FD93ULL; h3 ^= h3 >> 28;
        p += 32;
        n -= 32;
      }
      while (n) {
        const size_t k = std::min<size_t>(n, 8);
        uint64_t w = 0;
        memcpy(&w, p, k);
        h0 = (h0 ^ w) * 0xFF51AFD7ED558CCDULL; h0 ^= h0 >> 29;
        p += k;
        n -= k;
      }
      ph[q] = fp_mix(fp_mix(fp_mix(h0, h1), h2), h3);
    }
  });
  uint64_t h = 0x414E4C43ULL ^ bytes;
  for (uint64_t v : ph) h = fp_mix(h, v);
  return h;
}
template <class V>
uint64_t array_checksum(uint64_t h, const V& v) {
  return fp_mix(fp_mix(h, v.size()), bytes_checksum(v.data(), v.size() * sizeof(typename V::value_type)));
}
uint64_t content_checksum(const HostIndex& ix) {
  uint64_t h = 0x414E4C58ULL;
  h = array_checksum(h, ix.ana_key);
  h = array_checksum(h, ix.ana_inst_off);
  h = array_checksum(h, ix.ana_charcount);
  h = array_checksum(h, ix.inst_vocab);
  h = array_checksum(h, ix.inst_freq);
  h = array_checksum(h, ix.inst_gid);
  h = array_checksum(h, ix.inst_rows);
  h = array_checksum(h, ix.table);
  h = array_checksum(h, ix.bloom);
  h = array_checksum(h, ix.post_ana);
  h = array_checksum(h, ix.post_cls);
  h = array_checksum(h, ix.active_classes);
  return h;
}
template <class V>
bool write_array(FILE* f, const V& v) {
  typedef typename V::value_type T;
  const uint64_t n = v.size(), b = sizeof(T);
  return fwrite(&n, 8, 1, f) == 1 && fwrite(&b, 8, 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == n);
}
template <class V>
bool read_array(FILE* f, V* v, uint64_t max_bytes) {
  typedef typename V::value_type T;
  uint64_t n = 0, b = 0;
  if (fread(&n, 8, 1, f) != 1 || fread(&b, 8, 1, f) != 1 || b != sizeof(T) || n > max_bytes / sizeof(T)) return false;
  const size_t bytes = (size_t)n * sizeof(T);
  if (bytes < ((size_t)8 << 20)) {
    v->resize(n);
    return n == 0 || fread(v->data(), sizeof(T), n, f) == n;
  }
  // large array: first touch and read on all cores (pread at this array's offset, 4 MB pieces)
  v->resize(n);
  if (!std::is_same<V, std::vector<T>>::value) prefault(v->data(), bytes);  // (a plain vector was just zero-filled)
  const long at = ftell(f);
  if (at < 0) return false;
  const int fd = fileno(f);
  const size_t piece = (size_t)4 << 20;
  std::atomic<bool> ok{true};
  parallel_ranges((bytes + piece - 1) / piece, 1, [&](unsigned, uint64_t lo, uint64_t hi) {
    size_t pos = (size_t)lo * piece;
    const size_t end = std::min(bytes, (size_t)hi * piece);
    while (pos < end) {
      const ssize_t got = pread(fd, (char*)v->data() + pos, end - pos, (off_t)at + (off_t)pos);
      if (got <= 0) {
        ok = false;
        return;
      }
      pos += (size_t)got;
    }
  });
  return ok && fseek(f, at + (long)bytes, SEEK_SET) == 0;
}
uint64_t fp_mix(uint64_t h, uint64_t x) {
  h ^= x + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
  h *= 0xFF51AFD7ED558CCDULL;
  return h ^ (h >> 32);
}
}  // namespace

// Test hook: per-array checksums of the built index (in file order), an order-independent digest of the table's
// slots and whether every slot is reachable by linear probing from its home position.
void HostModel::index_digest(uint64_t* out, size_t cap) const {
  uint64_t v[16] = {0};
  const HostIndex& ix = index;
  if (built) {
    v[0] = array_checksum(1, ix.ana_key);
    v[1] = array_checksum(1, ix.ana_inst_off);
    v[2] = array_checksum(1, ix.ana_charcount);
    v[3] = array_checksum(1, ix.inst_vocab);
    v[4] = array_checksum(1, ix.inst_freq);
    v[5] = array_checksum(1, ix.inst_gid);
    v[6] = array_checksum(1, ix.inst_rows);
    v[7] = array_checksum(1, ix.table);
    v[8] = array_checksum(1, ix.bloom);
    v[9] = array_checksum(1, ix.post_ana);
    v[10] = array_checksum(1, ix.post_cls);
    v[11] = array_checksum(1, ix.active_classes);
    const size_t slots = ix.table.size();
    uint64_t sum = 0, occupied = 0;
    bool reachable = true;
    for (size_t i = 0; i < slots; ++i) {
      const Slot& sl = ix.table[i];
      if (sl.post_cnt == 0) continue;
      ++occupied;
      sum += fp_mix(fp_mix(sl.fp, sl.post_off), sl.post_cnt);
      for (size_t j = fp_index(sl.fp, slots - 1); j != i; j = (j + 1) & (slots - 1))
        if (ix.table[j].post_cnt == 0) {
          reachable = false;  // an empty slot between the home position and the key: a probe would stop there
          break;
        }
    }
    v[12] = sum;
    v[13] = reachable ? 1 : 0;
    v[14] = occupied;
    v[15] = fp_mix(fp_mix(fp_mix(ix.norm_stride, ix.max_len), ix.max_charcount), fp_mix(ix.max_key_bits, ix.charcount_mask[0] ^ ix.charcount_mask[1]));
  }
  for (size_t i = 0; i < cap && i < 16; ++i) out[i] = v[i];
}

// Everything build_index reads: per vocabulary entry (in id order) its text, normalised symbols, frequency,
// vocabulary type and case flag, plus the alphabet size.
uint64_t HostModel::vocabulary_fingerprint() const {
  uint64_t h = fp_mix(0x414E4C49ULL, alphabet.size());
  for (const VocabEntry& v : decoder) {
    h = fp_mix(h, v.text.size());
    for (unsigned char c : v.text) h = fp_mix(h, c);
    h = fp_mix(h, v.syms.size());
    for (uint8_t s : v.syms) h = fp_mix(h, s);
    h = fp_mix(h, ((uint64_t)v.frequency << 32) | ((uint64_t)v.vocabtype << 1) | (v.first_lower ? 1 : 0));
  }
  return h;
}

bool HostModel::save_index(const std::string& path, std::string* err) const {
  if (!built) {
    *err = "Model has not been built yet! Call build() before save_index()";
    return false;
  }
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) {
    *err = "cannot open " + path + " for writing";
    return false;
  }
  const HostIndex& ix = index;
  IndexFileHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, kIndexMagic, 8);
  h.header_bytes = sizeof h;
  h.slot_bytes = sizeof(Slot);
  h.key_bytes = sizeof(Key192);
  h.max_k = ANL_MAX_K;
  h.fingerprint = vocabulary_fingerprint();
  h.shard = ix.shard;
  h.n_shards = ix.n_shards;
  h.norm_stride = ix.norm_stride;
  h.max_charcount = ix.max_charcount;
  h.max_len = ix.max_len;
  h.max_key_bits = ix.max_key_bits;
  h.sd = ix.sd;
  h.table_keys = ix.table_keys;
  h.checksum = content_checksum(ix);
  memcpy(h.charcount_mask, ix.charcount_mask, sizeof h.charcount_mask);
  memcpy(h.prime_of, ix.prime_of, sizeof h.prime_of);
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  ok = ok && write_array(f, ix.ana_key) && write_array(f, ix.ana_inst_off) && write_array(f, ix.ana_charcount) &&
       write_array(f, ix.inst_vocab) && write_array(f, ix.inst_freq) && write_array(f, ix.inst_gid) &&
       write_array(f, ix.inst_rows) && write_array(f, ix.table) && write_array(f, ix.bloom) && write_array(f, ix.post_ana) &&
       write_array(f, ix.post_cls) && write_array(f, ix.active_classes);
  ok = (fclose(f) == 0) && ok;
  if (!ok) *err = "short write to " + path;
  return ok;
}

bool HostModel::load_index(const std::string& path, std::string* err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    *err = "cannot open " + path;
    return false;
  }
  uint64_t file_bytes = 0;
  if (fseek(f, 0, SEEK_END) == 0) {
    const long end = ftell(f);
    file_bytes = end > 0 ? (uint64_t)end : 0;
  }
  rewind(f);
  IndexFileHeader h;
  HostIndex ix;
  auto bad = [&](const std::string& why) {
    fclose(f);
    *err = path + ": " + why;
    return false;
  };
  if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kIndexMagic, 8) != 0) return bad("not an analiticcl_b200 index file");
  if (h.header_bytes != sizeof h || h.slot_bytes != sizeof(Slot) || h.key_bytes != sizeof(Key192) || h.max_k != ANL_MAX_K)
    return bad("index file was written by a library with a different data layout");
  if (h.fingerprint != vocabulary_fingerprint())
    return bad("index file was built from a different vocabulary or alphabet (fingerprint mismatch)");
  // header fields the kernels size buffers and loops from
  if (h.sd < 0 || h.sd > 1 || h.n_shards < 1 || h.shard >= h.n_shards || h.max_len == 0 || h.max_len > (uint32_t)ANL_MAX_SYMBOLS ||
      h.max_charcount == 0 || h.max_charcount > h.max_len || h.max_key_bits == 0 || h.max_key_bits > 192 || h.norm_stride < 16 ||
      h.norm_stride % 16 != 0 || h.max_len + 2 > h.norm_stride || h.norm_stride > 256)
    return bad("index file is inconsistent (header fields out of range)");
  if (alphabet.size() + 1 > 168) return bad("alphabet has more classes than there are primes (168)");
  std::string verr;
  if (!check_variant_support(h.n_shards, &verr)) return bad(verr);  // what build_index refuses, a file must not bypass
  ix.shard = h.shard;
  ix.n_shards = h.n_shards;
  ix.norm_stride = h.norm_stride;
  ix.max_charcount = h.max_charcount;
  ix.max_len = h.max_len;
  ix.max_key_bits = h.max_key_bits;
  ix.sd = h.sd;
  ix.table_keys = h.table_keys;
  // derived tables are recomputed, not trusted: primes from the alphabet, the charcount mask from the anagrams below
  const uint32_t n_classes = (uint32_t)alphabet.size() + 1;  // incl. UNK
  for (uint32_t s = 0; s < n_classes; ++s) ix.prime_of[s] = kPrimes[s];
  if (memcmp(ix.prime_of, h.prime_of, sizeof ix.prime_of) != 0) return bad("index file is inconsistent (prime table)");
  const bool read_ok = read_array(f, &ix.ana_key, file_bytes) && read_array(f, &ix.ana_inst_off, file_bytes) &&
                       read_array(f, &ix.ana_charcount, file_bytes) && read_array(f, &ix.inst_vocab, file_bytes) &&
                       read_array(f, &ix.inst_freq, file_bytes) && read_array(f, &ix.inst_gid, file_bytes) &&
                       read_array(f, &ix.inst_rows, file_bytes) && read_array(f, &ix.table, file_bytes) &&
                       read_array(f, &ix.bloom, file_bytes) && read_array(f, &ix.post_ana, file_bytes) &&
                       read_array(f, &ix.post_cls, file_bytes) && read_array(f, &ix.active_classes, file_bytes);
  if (!read_ok) return bad("truncated or corrupt index file");
  if (content_checksum(ix) != h.checksum) return bad("truncated or corrupt index file (content checksum mismatch)");
  // structural checks: everything the kernels index with must be in range (a file with a valid checksum can still
  // come from a buggy or hostile writer)
  const size_t ninst = ix.inst_vocab.size(), nana = ix.ana_key.size();
  auto pow2 = [](size_t n) { return n != 0 && (n & (n - 1)) == 0; };
  bool sane = nana > 0 && ninst >= nana && ix.ana_inst_off.size() == nana + 1 && ix.ana_charcount.size() == nana &&
              ix.inst_freq.size() == ninst && (ix.inst_gid.empty() == (ix.n_shards == 1)) &&
              (ix.inst_gid.empty() || ix.inst_gid.size() == ninst) &&
              ix.inst_rows.size() == ninst * (size_t)ix.norm_stride && pow2(ix.table.size()) &&
              pow2(ix.bloom.size()) && ix.post_cls.size() == ix.post_ana.size() && ix.post_ana.size() < 0xFFFFFFF0ull &&
              ix.ana_inst_off[0] == 0 && ix.ana_inst_off[nana] == ninst && !ix.active_classes.empty() &&
              ix.active_classes.size() <= n_classes && ix.table_keys > 0 && ix.table_keys * 2 <= ix.table.size();
  if (!sane) return bad("index file is inconsistent");
  std::atomic<bool> ok{true};
  auto check = [&](uint64_t n, auto pred) {  // pred(i) for i in [0, n) on all cores
    parallel_ranges(n, 1u << 16, [&](unsigned, uint64_t lo, uint64_t hi) {
      bool good = true;
      for (uint64_t i = lo; i < hi && good; ++i) good = pred(i);
      if (!good) ok = false;
    });
    return ok.load();
  };
  // active classes: ascending symbols of the alphabet
  for (size_t i = 0; i < ix.active_classes.size(); ++i)
    if (ix.active_classes[i] >= n_classes || (i && ix.active_classes[i] <= ix.active_classes[i - 1])) ok = false;
  bool is_active[256] = {false};
  for (uint8_t c : ix.active_classes) is_active[c] = true;
  // anagrams: keys strictly ascending (has() binary-searches them), CSR offsets strictly ascending, charcounts in range
  ok = ok && check(nana, [&](uint64_t r) {
    return ix.ana_inst_off[r] < ix.ana_inst_off[r + 1] && ix.ana_inst_off[r + 1] <= ninst &&
           (r == 0 || key_less(ix.ana_key[r - 1], ix.ana_key[r])) && ix.ana_charcount[r] >= 1 &&
           ix.ana_charcount[r] <= ix.max_charcount && key_bits(ix.ana_key[r]) <= ix.max_key_bits;
  });
  // instances: vocabulary ids, row lengths and symbols; a shard's global gather ids ascend
  ok = ok && check(ninst, [&](uint64_t g) {
    const uint8_t* row = ix.inst_rows.data() + g * ix.norm_stride;
    if (ix.inst_vocab[g] >= decoder.size() || row[0] == 0 || row[0] > ix.max_len || (row[1] & ~ROW_FIRST_LOWER)) return false;
    for (uint32_t i = 0; i < row[0]; ++i)
      if (!is_active[row[2 + i]]) return false;
    return ix.inst_gid.empty() || g == 0 || ix.inst_gid[g - 1] < ix.inst_gid[g];
  });
  // every row of an anagram has that anagram's length
  ok = ok && check(nana, [&](uint64_t r) {
    for (uint32_t g = ix.ana_inst_off[r]; g < ix.ana_inst_off[r + 1]; ++g)
      if (ix.inst_rows[(size_t)g * ix.norm_stride] != ix.ana_charcount[r]) return false;
    return true;
  });
  // postings: anagram rank and class in range
  ok = ok && check(ix.post_ana.size(), [&](uint64_t t) {
    return ix.post_ana[t] < nana && (ix.post_cls[t] == POST_SELF || (ix.sd == 1 && is_active[ix.post_cls[t]]));
  });
  // table: posting ranges in range, occupancy as declared (a full table would make the linear probe spin forever)
  std::atomic<uint64_t> occupied{0};
  ok = ok && check(ix.table.size(), [&](uint64_t i) {
    const Slot& sl = ix.table[i];
    if (sl.post_cnt == 0) return true;
    occupied.fetch_add(1, std::memory_order_relaxed);
    return (uint64_t)sl.post_off + sl.post_cnt <= ix.post_ana.size();
  });
  if (!ok || occupied.load() != ix.table_keys) return bad("index file is inconsistent");
  for (uint16_t cc : ix.ana_charcount) ix.charcount_mask[cc >> 6] |= 1ull << (cc & 63);
  if (memcmp(ix.charcount_mask, h.charcount_mask, sizeof ix.charcount_mask) != 0)
    return bad("index file is inconsistent (charcount mask)");
  fclose(f);
  index = std::move(ix);
  build_language_model();
  built = true;
  return true;
}

bool HostModel::ensure_msets(uint32_t J, std::string* err) {
  if (J > (uint32_t)ANL_MAX_K) J = ANL_MAX_K;
  if (index.mset_built_j >= J && !(J == 0)) return true;
  if (J == 0) return true;
  const std::vector<uint8_t>& cls = index.active_classes;
  const size_t A = cls.size();
  // count first: C(A + j - 1, j) multisets of size j
  uint64_t total = 0;
  for (uint32_t j = 1; j <= J; ++j) {
    long double c = 1;
    for (uint32_t i = 1; i <= j; ++i) c = c * (A + j - i) / i;
    total += (uint64_t)(c + 0.5L);
  }
  if (total > (64ull << 20)) {
    *err = "insertion neighbourhood too large for this alphabet and anagram distance (" + std::to_string(total) +
           " multisets)";
    return false;
  }
  std::vector<MsetEntry> out;
  out.reserve(total);
  uint32_t ends[ANL_MAX_K + 1] = {0};
  for (uint32_t j = 1; j <= J; ++j) {
    // non-decreasing index tuples of length j over [0, A)
    std::vector<size_t> ids(j, 0);
    for (;;) {
      MsetEntry e;
      e.hsum = 0;
      e.j = (uint8_t)j;
      for (int t = 0; t < 6; ++t) e.cls[t] = 0xFF;
      for (uint32_t t = 0; t < j; ++t) {
        e.cls[t] = cls[ids[t]];
        e.hsum += class_rnd(cls[ids[t]]);
      }
      e.maxcls = cls[ids[j - 1]];
      out.push_back(e);
      int t = (int)j - 1;
      while (t >= 0 && ids[t] == A - 1) --t;
      if (t < 0) break;
      size_t v = ids[t] + 1;
      for (uint32_t u = (uint32_t)t; u < j; ++u) ids[u] = v;
    }
    ends[j] = (uint32_t)out.size();
  }
  for (uint32_t j = J + 1; j <= (uint32_t)ANL_MAX_K; ++j) ends[j] = ends[J];
  index.mset = std::move(out);
  memcpy(index.mset_end, ends, sizeof ends);
  index.mset_built_j = J;
  return true;
}

// ---- host-side queries -------------------------------------------------------------------------------------------
int64_t HostModel::vocab_id(const char* text, size_t len) const {
  auto it = encoder.find(std::string(text, len));
  return it == encoder.end() ? -1 : (int64_t)it->second;
}

bool HostModel::has(const char* text, size_t len) const {
  if (!built) return false;
  std::vector<uint8_t> syms;
  alphabet.encode(text, len, &syms);
  Key192 k;
  if (!key_of(syms.data(), syms.size(), &k)) return false;
  auto it = std::lower_bound(index.ana_key.begin(), index.ana_key.end(), k, key_less);
  if (it == index.ana_key.end() || !key_eq(*it, k)) return false;
  size_t r = it - index.ana_key.begin();
  for (uint32_t g = index.ana_inst_off[r]; g < index.ana_inst_off[r + 1]; ++g) {
    const std::string& t = decoder[index.inst_vocab[g]].text;
    if (t.size() == len && memcmp(t.data(), text, len) == 0) return true;
  }
  return false;
}

// Decodes UTF-8 into `out` (at most cap scalars); returns the count, or cap + 1 if it does not fit.
static size_t decode_small(const char* s, size_t n, char32_t* out, size_t cap) {
  size_t c = 0, i = 0;
  while (i < n) {
    unsigned l;
    const uint32_t cp = u8decode(s + i, n - i, &l);
    if (c >= cap) return cap + 1;
    out[c++] = cp;
    i += l;
  }
  return c;
}

// src/lib.rs:1733-1756.  Before paying for the edit script, an exact prefilter: every deletion
// (insertion) chunk of the script consists of characters of the input's (candidate's) "middle" --
// what is left after stripping the common prefix and suffix -- because the diff runs on the middles
// and the clean-up passes only merge equalities that lie between edits or rotate an edit over equal
// characters.  A pattern whose `-[..]` / `+[..]` instruction has no option made of such characters
// cannot match; if that rules out every confusable, the weight is 1.0 without computing the script.
template <class Ch, class Opt>
static bool confusable_possible(const std::vector<Confusable>& confusables, const Ch* ma, size_t la, const Ch* mb, size_t lb,
                                Opt options_of) {
  auto made_of = [](const auto& opt, const Ch* mid, size_t n) {
    for (auto c : opt) {
      const uint32_t cv = sizeof(c) == 1 ? (uint32_t)(unsigned char)c : (uint32_t)c;
      bool found = false;
      for (size_t i = 0; i < n; ++i) found = found || (uint32_t)mid[i] == cv;
      if (!found) return false;
    }
    return true;
  };
  for (const Confusable& c : confusables) {
    bool possible = true;
    for (const ConfusableInstr& ins : c.script) {
      if (ins.op == 0) continue;
      bool any = false;
      for (const auto& opt : options_of(ins))
        if (ins.op < 0 ? made_of(opt, ma, la) : made_of(opt, mb, lb)) {
          any = true;
          break;
        }
      if (!any) {
        possible = false;
        break;
      }
    }
    if (possible) return true;
  }
  return false;
}

static std::atomic<uint64_t> g_cw_calls{0}, g_cw_prefilter_pass{0}, g_cw_fast{0}, g_cw_full{0};
void confusable_stats(uint64_t* v) {
  v[0] = g_cw_calls.exchange(0);
  v[1] = g_cw_prefilter_pass.exchange(0);
  v[2] = g_cw_fast.exchange(0);
  v[3] = g_cw_full.exchange(0);
}
static const bool g_cw_count = getenv("ANL_PROFILE") != nullptr;

double HostModel::compute_confusable_weight(const char* input, size_t len, uint64_t candidate) const {
  double weight = 1.0;
  if (g_cw_count) g_cw_calls.fetch_add(1, std::memory_order_relaxed);
  if (candidate >= decoder.size() || confusables.empty()) return weight;
  const VocabEntry& ce = decoder[candidate];
  const std::string& cand = ce.text;
  bool input_ascii = true;
  for (size_t i = 0; i < len; ++i) input_ascii = input_ascii && ((unsigned char)input[i] < 0x80);
  if (input_ascii && ce.ascii) {
    // bytes are scalar values: work in place
    const unsigned char* a = reinterpret_cast<const unsigned char*>(input);
    const unsigned char* b = reinterpret_cast<const unsigned char*>(cand.data());
    const size_t na = len, nb = cand.size();
    size_t p = 0;
    while (p < na && p < nb && a[p] == b[p]) ++p;
    size_t s = 0;
    while (s < na - p && s < nb - p && a[na - 1 - s] == b[nb - 1 - s]) ++s;
    const size_t la = na - p - s, lb = nb - p - s;
    if (!confusable_possible(confusables, a + p, la, b + p, lb,
                             [](const ConfusableInstr& i) -> const std::vector<std::string>& { return i.options; }))
      return weight;
    if (g_cw_count) g_cw_prefilter_pass.fetch_add(1, std::memory_order_relaxed);
    if (la <= 1 && lb <= 1 && all_confusables_simple) {
      if (g_cw_count) g_cw_fast.fetch_add(1, std::memory_order_relaxed);
      // a single substitution / insertion / deletion: the script is [=prefix] [-x] [+y] [=suffix]; the
      // clean-up passes can only slide the one-character edit, which does not change the edit chunks
      EditView v[4];
      size_t nv = 0;
      if (p) v[nv++] = EditView{0, input, p};
      if (la) v[nv++] = EditView{-1, input + p, la};
      if (lb) v[nv++] = EditView{1, cand.data() + p, lb};
      if (s) v[nv++] = EditView{0, input + na - s, s};
      for (const Confusable& c : confusables)
        if (confusable_found_in_views(c, v, nv)) weight *= c.weight;
      return weight;
    }
  } else {
    static const size_t CAP = 128;
    char32_t a[CAP], b[CAP];
    const size_t na = decode_small(input, len, a, CAP), nb = decode_small(cand.data(), cand.size(), b, CAP);
    if (na <= CAP && nb <= CAP) {
      size_t p = 0;
      while (p < na && p < nb && a[p] == b[p]) ++p;
      size_t s = 0;
      while (s < na - p && s < nb - p && a[na - 1 - s] == b[nb - 1 - s]) ++s;
      if (!confusable_possible(confusables, a + p, na - p - s, b + p, nb - p - s,
                               [](const ConfusableInstr& i) -> const std::vector<std::u32string>& { return i.options32; }))
        return weight;
    }
  }
  if (g_cw_count) g_cw_full.fetch_add(1, std::memory_order_relaxed);
  {
    EditView v[64];
    size_t nv = 0;
    if (edit_views_fixed(input, len, cand.data(), cand.size(), v, &nv)) {
      for (const Confusable& c : confusables)
        if (confusable_found_in_views(c, v, nv)) weight *= c.weight;
      return weight;
    }
  }
  const std::vector<EditInstruction> script = shortest_edit_script(std::string(input, len), cand);
  for (const Confusable& c : confusables)
    if (confusable_found_in(c, script)) weight *= c.weight;
  return weight;
}

}  // namespace anl
