// search.cpp -- the batch producer of find_all_matches (src/lib.rs:1790-1957, src/search.rs:190-336).
//
// Host-side segmentation: token boundaries, boundary strengths, n-gram spans per hard-delimited
// batch, redundant-match pruning.  Unlike the reference -- which calls find_variants once per
// segment from a rayon loop -- the segments of the whole text are looked up in at most two GPU
// batches: all unigrams first, then every higher-order segment that the unigram results do not make
// redundant (redundant_match only ever inspects unigram results, src/search.rs:317-336).
#include "search.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>

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
// (The per-part scratch vectors belong to the CALLING thread (thread_local arena indexed by part): after the
// first call their capacity is already mapped, so a 100 M-token stream does not page-fault hundreds of MB per
// call, and they outlive whichever thread ran the part -- when the pool is busy the parts run on short-lived
// threads whose own thread_locals are gone by the time the results are read.)
const std::vector<Boundary>& find_boundaries(const std::string& text) {
  const size_t n = text.size();
  const unsigned nt_max = host_threads();
  static thread_local std::vector<std::vector<Boundary>> arena_tl;
  std::vector<std::vector<Boundary>>& arena = arena_tl;  // (a lambda naming a thread_local would see the executing thread's)
  if (arena.size() < nt_max) arena.resize(nt_max);
  std::vector<std::vector<Boundary>*> part(nt_max, nullptr);
  std::vector<std::pair<uint64_t, uint64_t>> range(nt_max, {0, 0});
  const unsigned used = parallel_ranges(n, 1u << 16, [&](unsigned t, uint64_t lo, uint64_t hi) {
    // move both ends to the start of a character (skip UTF-8 continuation bytes)
    while (lo < n && lo > 0 && ((unsigned char)text[lo] & 0xC0) == 0x80) ++lo;
    while (hi < n && ((unsigned char)text[hi] & 0xC0) == 0x80) ++hi;
    range[t] = {lo, hi};
    std::vector<Boundary>& out = arena[t];
    out.clear();
    part[t] = &out;
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

// the batches: spans between hard boundaries (src/lib.rs:1822)
void list_batches(const std::vector<Boundary>& bounds, std::vector<BatchDesc>* out) {
  out->clear();
  size_t begin = 0, begin_index = 0;
  for (size_t i = 0; i < bounds.size(); ++i) {
    if (bounds[i].strength != BOUNDARY_HARD || bounds[i].begin == begin) continue;
    out->push_back(BatchDesc{begin, begin_index, i});
    begin = bounds[i].end;
    begin_index = i + 1;
  }
}

bool segment_on_device(size_t len) {
  if (const char* e = getenv("ANL_SEGMENT")) {
    if (e[0] == 'h') return false;
    if (e[0] == 'd') return len > 0 && len < 0x7FFFFFF0ull;
  }
  if (len == 0) return false;
  size_t min_len = 1u << 16;
  if (const char* e = getenv("ANL_SEGMENT_DEVICE_MIN")) min_len = (size_t)std::max(0ll, atoll(e));
  return len >= min_len && len < 0x7FFFFFF0ull;
}

bool segment_any(int device, const std::string& text, uint32_t max_ngram, Segmentation* out, std::string* err) {
  out->text_len = text.size();
  out->max_ngram = max_ngram;
  if (device >= 0 && segment_on_device(text.size()))
    return segment_text_device(device, text, max_ngram, &out->st, err, &out->bounds, &out->batches);
  segment_text(text, max_ngram, &out->st, &out->bounds, &out->batches);
  return true;
}

void segment_text(const std::string& text, uint32_t max_ngram, SegmentedText* stp, PodBuffer<Boundary>* bounds_out,
                  PodBuffer<BatchDesc>* batches_out) {
  SegmentedText& st = *stp;
  st.segs.clear();
  st.batch_first.resize(1);
  st.batch_first[0] = 0;
  if (bounds_out) bounds_out->clear();
  if (batches_out) batches_out->clear();
  if (text.empty()) return;
  const std::vector<Boundary>& bounds = find_boundaries(text);
  // (a lambda naming a thread_local would see the executing thread's instance: bind it to a local reference)
  static thread_local std::vector<BatchDesc> descs_tl;
  std::vector<BatchDesc>& descs = descs_tl;
  list_batches(bounds, &descs);
  const size_t nb = descs.size();
  const unsigned nt_max = host_threads();
  static thread_local std::vector<std::vector<SegmentSpan>> arena_tl;  // per-part scratch owned by the calling thread
  std::vector<std::vector<SegmentSpan>>& arena = arena_tl;
  if (arena.size() < nt_max) arena.resize(nt_max);
  std::vector<std::vector<SegmentSpan>*> part(nt_max, nullptr);
  std::vector<std::pair<uint64_t, uint64_t>> range(nt_max, {0, 0});
  st.batch_first.resize(nb + 1);  // segments per batch first, then the exclusive prefix
  uint64_t* count = st.batch_first.data() + 1;
  const unsigned used = parallel_ranges(nb, 256, [&](unsigned t, uint64_t lo, uint64_t hi) {
    range[t] = {lo, hi};
    std::vector<SegmentSpan>& out = arena[t];
    out.clear();
    part[t] = &out;
    if (hi > lo) out.reserve((descs[hi - 1].end_index - descs[lo].begin_index + 1) * (size_t)max_ngram + 16);
    for (uint64_t k = lo; k < hi; ++k) {
      const BatchDesc& d = descs[k];
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
  if (bounds_out) {
    bounds_out->resize(bounds.size());
    if (!bounds.empty()) memcpy(bounds_out->data(), bounds.data(), bounds.size() * sizeof(Boundary));
  }
  if (batches_out) {
    batches_out->resize(descs.size());
    if (!descs.empty()) memcpy(batches_out->data(), descs.data(), descs.size() * sizeof(BatchDesc));
  }
}

// most_likely_sequence, src/lib.rs:2088-2495, without language model and context rules.
//
// The reference hands a weighted FST to rustfst (start state + one state per boundary of the batch; a transition
// per (segment, variant) from the boundary before the segment to the boundary after it, cost n + (1 - score) in
// f32 with n = tokens covered; cost n + 1 for a unigram without variants, copied from the input; cost 100 for
// the epsilon fail-safe between neighbouring states) and, with nothing but the variant model to weigh, keeps the
// cheapest of the n shortest paths.  States are ordered by position, every transition points forward, so the
// cheapest path is a single sweep over the states: best[t] = min over incoming transitions of best[from] + cost,
// accumulated in f32 from the start state like the path weight rustfst reports.  Transitions are bucketed by
// their target state (counting sort) in (source state, segment, variant) order, and only a strictly lower cost
// replaces the incumbent: among equal-cost paths the earliest source state, then the earliest transition wins
// (rustfst's own choice among ties is not pinned by any reference test).
bool most_likely_sequence(const Boundary* bounds, size_t nbounds, size_t end_offset, const SegmentSpan* segs, size_t nsegs,
                          const BatchVariants& variants, std::vector<SequenceStep>* out) {
  out->clear();
  struct Arc {
    uint32_t from, seg;
    int32_t variant;  // -1 = out of vocabulary, -2 = epsilon
    float cost;
  };
  const size_t nstates = nbounds + 1;  // 0 = start, 1 + i = boundary i
  // boundary positions -> state: begin offsets and end offsets are both ascending, binary search replaces the
  // reference's scan over all boundaries per match (:2142-2148; offsets are unique, so "last hit" = "the hit")
  auto state_ending_at = [&](size_t pos) -> long {  // boundary whose end == pos
    size_t lo = 0, hi = nbounds;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (bounds[mid].end < pos) lo = mid + 1; else hi = mid;
    }
    return (lo < nbounds && bounds[lo].end == pos) ? (long)lo : -1;
  };
  auto state_starting_at = [&](size_t pos) -> long {  // boundary whose begin == pos
    size_t lo = 0, hi = nbounds;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (bounds[mid].begin < pos) lo = mid + 1; else hi = mid;
    }
    return (lo < nbounds && bounds[lo].begin == pos) ? (long)lo : -1;
  };
  // (scratch kept per thread: a text has one lattice per sentence)
  static thread_local std::vector<Arc> arcs, by_target;
  static thread_local std::vector<uint32_t> target, fill, order, cursor;
  static thread_local std::vector<float> best;
  static thread_local std::vector<int64_t> via;
  arcs.clear();
  target.clear();
  size_t labelled = 0;
  for (size_t k = 0; k < nsegs; ++k) {
    const long next = state_starting_at(segs[k].end);
    if (next < 0) continue;  // cannot happen for producer segments (the reference would panic)
    const long prev = state_ending_at(segs[k].begin);
    const long n = prev >= 0 ? next - prev : next + 1;
    const uint32_t from = prev >= 0 ? (uint32_t)prev + 1 : 0;
    const uint32_t cnt = variants.count[k];
    if (cnt > 0) {
      for (uint32_t j = 0; j < cnt; ++j) {
        const float cost = (float)n + (1.0f - (float)variants.score[variants.first[k] + j]);  // :2203-2204
        arcs.push_back(Arc{from, (uint32_t)k, (int32_t)j, cost});
        target.push_back((uint32_t)next + 1);
      }
      labelled += cnt;
    } else if (n == 1) {
      arcs.push_back(Arc{from, (uint32_t)k, -1, (float)n + 1.0f});  // :2223
      target.push_back((uint32_t)next + 1);
      ++labelled;
    }
  }
  if (labelled == 0) return false;
  for (size_t i = 0; i < nbounds; ++i) {  // fail-safe, :2249-2259
    arcs.push_back(Arc{(uint32_t)i, 0, -2, 100.0f});
    target.push_back((uint32_t)i + 1);
  }
  // bucket by target state; inside a bucket order by source state (stable counting sort over the arc list, which
  // is in (segment, variant) order) so that the sweep below sees candidates in (source, segment, variant) order
  fill.assign(nstates + 1, 0);
  for (uint32_t t : target) ++fill[t + 1];
  for (size_t s = 0; s < nstates; ++s) fill[s + 1] += fill[s];
  by_target.resize(arcs.size());
  {
    // counting sort by source state first (stable), then the stable scatter by target state
    cursor.assign(nstates + 1, 0);
    for (const Arc& a : arcs) ++cursor[a.from + 1];
    for (size_t s = 0; s < nstates; ++s) cursor[s + 1] += cursor[s];
    order.resize(arcs.size());
    for (size_t a = 0; a < arcs.size(); ++a) order[cursor[arcs[a].from]++] = (uint32_t)a;
    cursor.assign(fill.begin(), fill.end() - 1);
    for (uint32_t a : order) by_target[cursor[target[a]]++] = arcs[a];
  }
  const float INF = std::numeric_limits<float>::infinity();
  best.assign(nstates, INF);
  via.assign(nstates, -1);  // index into by_target
  best[0] = 0.0f;
  for (size_t t = 1; t < nstates; ++t) {
    for (uint32_t a = fill[t]; a < fill[t + 1]; ++a) {
      const Arc& arc = by_target[a];
      if (best[arc.from] == INF) continue;
      const float c = best[arc.from] + arc.cost;
      if (c < best[t]) {
        best[t] = c;
        via[t] = a;
      }
    }
  }
  long fin = -1;  // final states, :2119-2122
  for (size_t i = 0; i < nbounds; ++i)
    if ((bounds[i].begin == end_offset || bounds[i].end == end_offset) && (fin < 0 || best[i + 1] < best[fin])) fin = (long)i + 1;
  if (fin < 0 || best[fin] == INF) return false;
  for (size_t t = (size_t)fin; t != 0; t = by_target[via[t]].from) {
    const Arc& arc = by_target[via[t]];
    if (arc.variant != -2) out->push_back(SequenceStep{arc.seg, arc.variant});
  }
  std::reverse(out->begin(), out->end());
  return true;
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
