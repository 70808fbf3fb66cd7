// search.cpp -- the batch producer of find_all_matches (src/lib.rs:1790-1957, src/search.rs:190-336).
//
// Host-side segmentation: token boundaries, boundary strengths, n-gram spans per hard-delimited
// batch, redundant-match pruning.  Unlike the reference -- which calls find_variants once per
// segment from a rayon loop -- the segments of the whole text are looked up in at most two GPU
// batches: all unigrams first, then every higher-order segment that the unigram results do not make
// redundant (redundant_match only ever inspects unigram results, src/search.rs:317-336).
#include "search.h"

#include <algorithm>

#include "hostpool.h"
#include "unicode_tables.h"

namespace anl {

static inline uint32_t decode_at(const std::string& s, size_t i, unsigned* len) {
  const unsigned char c = (unsigned char)s[i];
  unsigned l = c < 0x80 ? 1 : ((c & 0xE0) == 0xC0 ? 2 : ((c & 0xF0) == 0xE0 ? 3 : ((c & 0xF8) == 0xF0 ? 4 : 1)));
  if (i + l > s.size()) l = (unsigned)(s.size() - i);
  *len = l;
  if (l == 1) return c;
  uint32_t cp = c & (0xFFu >> (l + 1));
  for (unsigned k = 1; k < l; ++k) cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3F);
  return cp;
}

// src/search.rs:190-235: a boundary is a maximal run of non-alphabetic characters; the text always
// ends with a boundary (possibly of length zero).  The scan runs in parallel over byte ranges; a run
// that crosses a range border is stitched back together afterwards.
// (Scratch vectors are thread_local: the pool's workers are persistent, so after the first call their
// capacity is already mapped and a 100 M-token stream does not page-fault hundreds of MB per call.)
const std::vector<Boundary>& find_boundaries(const std::string& text) {
  const size_t n = text.size();
  const unsigned nt_max = host_threads();
  std::vector<std::vector<Boundary>*> part(nt_max, nullptr);
  std::vector<std::pair<uint64_t, uint64_t>> range(nt_max, {0, 0});
  const unsigned used = parallel_ranges(n, 1u << 16, [&](unsigned t, uint64_t lo, uint64_t hi) {
    // move both ends to the start of a character (skip UTF-8 continuation bytes)
    while (lo < n && lo > 0 && ((unsigned char)text[lo] & 0xC0) == 0x80) ++lo;
    while (hi < n && ((unsigned char)text[hi] & 0xC0) == 0x80) ++hi;
    range[t] = {lo, hi};
    static thread_local std::vector<Boundary> scratch;
    scratch.clear();
    part[t] = &scratch;
    std::vector<Boundary>& out = scratch;
    out.reserve((size_t)(hi - lo) / 5 + 16);
    bool open = false;
    size_t start = 0;
    for (size_t i = lo; i < hi;) {
      unsigned l;
      const bool alpha = anl_unicode::is_alphabetic(decode_at(text, i, &l));
      if (open && alpha) {
        out.push_back(Boundary{start, i, BOUNDARY_NONE});
        open = false;
      } else if (!open && !alpha) {
        start = i;
        open = true;
      }
      i += l;
    }
    if (open) out.push_back(Boundary{start, (size_t)hi, BOUNDARY_NONE});  // may continue in the next range
  });
  static thread_local std::vector<Boundary> out_tl;
  std::vector<Boundary>& out = out_tl;
  out.clear();
  size_t total = 1;
  for (unsigned t = 0; t < used; ++t) total += part[t] ? part[t]->size() : 0;
  out.reserve(total);
  for (unsigned t = 0; t < used; ++t) {
    if (!part[t]) continue;
    for (const Boundary& b : *part[t]) {
      if (!out.empty() && out.back().end == b.begin)
        out.back().end = b.end;  // the same run, cut by a range border
      else
        out.push_back(b);
    }
  }
  if (out.empty() || out.back().end != n) out.push_back(Boundary{n, n, BOUNDARY_NONE});
  // src/search.rs:238-258: last or multi-byte boundary = hard; ' - _ = weak; else normal
  for (size_t i = 0; i < out.size(); ++i) {
    const size_t len = out[i].end - out[i].begin;
    if (i + 1 == out.size() || len > 1) {
      out[i].strength = BOUNDARY_HARD;
    } else {
      const char c = len == 1 ? text[out[i].begin] : 0;
      out[i].strength = (c == '\'' || c == '-' || c == '_') ? BOUNDARY_WEAK : BOUNDARY_NORMAL;
    }
  }
  return out;
}

// src/search.rs:262-312
void find_match_ngrams(const std::string& text, const Boundary* bounds, size_t nbounds, uint32_t order, size_t begin, size_t end,
                       std::vector<SegmentSpan>* out) {
  auto usable = [&](size_t b, size_t e) { return e > b && !(e - b == 1 && text[b] == ' '); };
  for (size_t i = 0; i + order - 1 < nbounds; ++i) {
    const Boundary& right = bounds[i + order - 1];
    if (right.begin > end) break;
    if (usable(begin, right.begin)) out->push_back(SegmentSpan{begin, right.begin, order});
    begin = bounds[i].end;
  }
  if (begin < end && usable(begin, end)) {
    // Match::internal_boundaries (src/search.rs:99-116) counts with a first/last window: the first
    // inner boundary only opens the window, every later one extends it.
    long first = -1;
    size_t last_plus1 = 0;
    for (size_t k = 0; k < nbounds; ++k) {
      if (bounds[k].begin > begin && bounds[k].end < end) {
        if (first < 0)
          first = (long)k;
        else
          last_plus1 = k + 1;
      }
    }
    const size_t inner = (first < 0 || (size_t)first >= last_plus1) ? 0 : last_plus1 - (size_t)first;
    if (inner == order) out->push_back(SegmentSpan{begin, end, order});
  }
}

void segment_text(const std::string& text, uint32_t max_ngram, SegmentedText* stp) {
  SegmentedText& st = *stp;
  st.segs.clear();
  st.batch_first.resize(1);
  st.batch_first[0] = 0;
  if (text.empty()) return;
  const std::vector<Boundary>& bounds = find_boundaries(text);
  // the batches: spans between hard boundaries (src/lib.rs:1822)
  struct Desc {
    size_t begin, begin_index, end_index;
  };
  // (a lambda naming a thread_local would see the executing thread's instance: bind it to a local reference)
  static thread_local std::vector<Desc> descs_tl;
  std::vector<Desc>& descs = descs_tl;
  descs.clear();
  {
    size_t begin = 0, begin_index = 0;
    for (size_t i = 0; i < bounds.size(); ++i) {
      if (bounds[i].strength != BOUNDARY_HARD || bounds[i].begin == begin) continue;
      descs.push_back(Desc{begin, begin_index, i});
      begin = bounds[i].end;
      begin_index = i + 1;
    }
  }
  const size_t nb = descs.size();
  const unsigned nt_max = host_threads();
  std::vector<std::vector<SegmentSpan>*> part(nt_max, nullptr);
  std::vector<std::pair<uint64_t, uint64_t>> range(nt_max, {0, 0});
  st.batch_first.resize(nb + 1);  // segments per batch first, then the exclusive prefix
  uint64_t* count = st.batch_first.data() + 1;
  const unsigned used = parallel_ranges(nb, 256, [&](unsigned t, uint64_t lo, uint64_t hi) {
    range[t] = {lo, hi};
    static thread_local std::vector<SegmentSpan> scratch;
    scratch.clear();
    part[t] = &scratch;
    std::vector<SegmentSpan>& out = scratch;
    if (hi > lo) out.reserve((descs[hi - 1].end_index - descs[lo].begin_index + 1) * (size_t)max_ngram + 16);
    for (uint64_t k = lo; k < hi; ++k) {
      const Desc& d = descs[k];
      const size_t before = out.size();
      for (uint32_t order = 1; order <= max_ngram; ++order)
        find_match_ngrams(text, bounds.data() + d.begin_index, d.end_index + 1 - d.begin_index, order, d.begin,
                          bounds[d.end_index].begin, &out);
      count[k] = out.size() - before;
    }
  });
  st.batch_first[0] = 0;
  for (size_t k = 0; k < nb; ++k) st.batch_first[k + 1] += st.batch_first[k];  // count[k] lives in batch_first[k + 1]
  st.segs.resize(st.batch_first[nb]);
  parallel_ranges(used, 1, [&](unsigned, uint64_t lo, uint64_t hi) {
    for (uint64_t t = lo; t < hi; ++t)
      if (part[t] && !part[t]->empty())
        std::copy(part[t]->begin(), part[t]->end(), st.segs.data() + st.batch_first[range[t].first]);
  });
}

std::vector<uint64_t> byte_to_codepoint_map(const std::string& text) {
  std::vector<uint64_t> map(text.size() + 1, 0);
  uint64_t cp = 0;
  for (size_t i = 0; i < text.size();) {
    unsigned l;
    decode_at(text, i, &l);
    for (unsigned k = 0; k < l; ++k) map[i + k] = cp;
    i += l;
    ++cp;
  }
  map[text.size()] = cp;
  return map;
}

}  // namespace anl
