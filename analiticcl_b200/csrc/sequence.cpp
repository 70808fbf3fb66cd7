// sequence.cpp -- most_likely_sequence with the language model and the context rules
// (src/lib.rs:2088-2674, src/search.rs:338-524): host post-pass of find_all_matches.
//
// The reference hands a weighted FST to rustfst, asks for the `max_seq` shortest paths, scores each path with a bigram
// language model over the output tokens (lm_score) and with the context rules (test_context_rules) and keeps the
// sequence with the best weighted sum of the three normalised terms.  Here:
//   * the lattice is a DAG over the batch's boundaries, so the max_seq shortest paths come from per-state lists of the
//     best partial paths (each state merges the lists of its predecessors through a heap), no FST library;
//   * the language model is two open hash maps (unigram, bigram counts) keyed by vocabulary ids -- the reference's
//     lm_score_tokens only ever asks for bigrams and their unigram priors (:2643-2674);
//   * a context rule's pattern is compiled into a flat prefix program per position instead of a boxed tree.
// Tie rule (equal path costs; not pinned by any reference test): cost, then final state, then -- from the last arc
// backwards -- source state, arc in (segment, variant) order with the fail-safe arcs last, then the prefixes alike.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <queue>

#include "host_model.h"
#include "search.h"

namespace anl {

// ---- language model -------------------------------------------------------------------------------------------------
// src/lib.rs:2688-2751 into_ngram: the entry's text split on single spaces, each part looked up in the vocabulary
// (a part outside it counts as <unk> = 2; the reference always encodes with use_unk = true).  false: more than 5 parts.
bool HostModel::entry_tokens(uint64_t id, uint32_t* out, unsigned* n) const {
  const VocabEntry& e = decoder[id];
  *n = 0;
  if (e.tokencount > 5) return false;
  size_t pos = 0;
  std::string part;
  for (unsigned k = 0; k < e.tokencount; ++k) {
    const size_t sp = e.text.find(' ', pos);
    part.assign(e.text, pos, sp == std::string::npos ? std::string::npos : sp - pos);
    auto it = encoder.find(part);
    out[(*n)++] = it != encoder.end() ? (uint32_t)it->second : 2u;
    pos = sp == std::string::npos ? e.text.size() : sp + 1;
  }
  return true;
}

// "Constructing Language Model", src/lib.rs:246-295: every LM-typed entry is an n-gram with its frequency as count
void HostModel::build_language_model() {
  lm_unigram.clear();
  lm_bigram.clear();
  lm_ngrams = 0;
  if (decoder.size() > 0xFFFFFFFFull) return;  // (vocabulary ids are packed into 32 bits below; build_index refuses such a lexicon anyway)
  for (size_t id = 0; id < decoder.size(); ++id) {
    if (!(decoder[id].vocabtype & VT_LM)) continue;
    uint32_t t[5];
    unsigned n;
    if (!entry_tokens(id, t, &n)) continue;
    const uint32_t freq = decoder[id].frequency;
    if (n == 1) {
      auto r = lm_unigram.emplace(t[0], freq);
      if (!r.second) r.first->second += freq; else ++lm_ngrams;
    } else if (n == 2) {
      auto r = lm_bigram.emplace(((uint64_t)t[0] << 32) | t[1], freq);
      if (!r.second) r.first->second += freq; else ++lm_ngrams;
    } else {
      // higher orders are counted (have_lm) but never read: lm_score_tokens works on bigrams (:2649-2660)
      std::string key(reinterpret_cast<const char*>(t), n * sizeof(uint32_t));
      if (lm_higher.insert(key).second) ++lm_ngrams;
    }
  }
  lm_higher.clear();
}

static const float kTransitionSmoothingLogprob = -13.815510557964274f;  // src/search.rs:4

// src/lib.rs:2643-2674; token < 0 = out of vocabulary
void HostModel::lm_score_tokens(const int64_t* tokens, size_t n_tokens, float* logprob_out, double* perplexity) const {
  float logprob = 0.0f;
  uint64_t n = 0;
  for (size_t i = 1; i < n_tokens; ++i, ++n) {
    const int64_t a = tokens[i - 1], b = tokens[i];
    if (a < 0 || b < 0) {
      logprob += kTransitionSmoothingLogprob;
      continue;
    }
    auto joint = lm_bigram.find(((uint64_t)a << 32) | (uint64_t)b);
    if (joint == lm_bigram.end()) {
      logprob += kTransitionSmoothingLogprob;
      continue;
    }
    auto prior = lm_unigram.find((uint32_t)a);
    const uint32_t priorcount = prior != lm_unigram.end() ? prior->second : 1u;
    logprob += priorcount < joint->second ? logf((float)joint->second) : logf((float)joint->second / (float)priorcount);
  }
  *logprob_out = logprob;
  *perplexity = -1.0 / (double)n * (double)logprob;
}

// ---- context rules ---------------------------------------------------------------------------------------------------
namespace {
enum : uint8_t { OP_ANY, OP_NOLEX, OP_VOCAB, OP_FROMLEX, OP_NOT, OP_OR };

// PatternMatch::parse (src/search.rs:421-470) into a prefix program
bool compile_pattern(const HostModel& hm, const std::string& raw, std::vector<RuleOp>* code, std::string* err) {
  const std::string s = trim_unicode(raw);
  if (s == "?") {
    code->push_back(RuleOp{OP_ANY, 0, 0, 0});
  } else if (s == "^") {
    code->push_back(RuleOp{OP_NOLEX, 0, 0, 0});
  } else if (s.size() >= 3 && s[0] == '!' && s[1] == '(' && s.back() == ')') {  // negation over a disjunction
    code->push_back(RuleOp{OP_NOT, 0, 1, 0});
    return compile_pattern(hm, s.substr(2, s.size() - 3), code, err);
  } else if (s.find('|') != std::string::npos) {
    const size_t head = code->size();
    code->push_back(RuleOp{OP_OR, 0, 0, 0});
    uint16_t alternatives = 0;
    size_t pos = 0;
    for (;;) {
      const size_t bar = s.find('|', pos);
      if (!compile_pattern(hm, s.substr(pos, bar == std::string::npos ? std::string::npos : bar - pos), code, err)) return false;
      ++alternatives;
      if (bar == std::string::npos) break;
      pos = bar + 1;
    }
    (*code)[head].n = alternatives;
  } else if (!s.empty() && s[0] == '!') {
    code->push_back(RuleOp{OP_NOT, 0, 1, 0});
    return compile_pattern(hm, s.substr(1), code, err);
  } else if (!s.empty() && s[0] == '@') {
    const std::string source = s.substr(1), rel = "/" + source;
    for (size_t i = 0; i < hm.lexicons.size(); ++i) {
      const std::string& lx = hm.lexicons[i];
      if (lx == source || (lx.size() >= rel.size() && lx.compare(lx.size() - rel.size(), rel.size(), rel) == 0)) {
        code->push_back(RuleOp{OP_FROMLEX, (uint8_t)i, 0, 0});
        return true;
      }
    }
    *err = "WARNING: Context rule references lexicon or variant list '" + source + "' but this source was not loaded";
    return false;
  } else {
    auto it = hm.encoder.find(s);
    if (it == hm.encoder.end()) {
      *err = "WARNING: Context rule references word '" + s + "' but this word does not occur in any lexicon";
      return false;
    }
    code->push_back(RuleOp{OP_VOCAB, 0, 0, it->second});
  }
  return true;
}

struct SeqItem {
  uint64_t vocab_id;  // 0 = out of vocabulary
  uint32_t lexindex;
};
// PatternMatch::matches (src/search.rs:374-419): evaluates the expression at code[*pc] and leaves *pc behind it
bool eval_pattern(const RuleOp* code, size_t* pc, const SeqItem* seq, size_t n, size_t index) {
  const RuleOp op = code[(*pc)++];
  switch (op.kind) {
    case OP_ANY: return true;
    case OP_NOLEX: return index < n && (seq[index].lexindex == 0 || seq[index].vocab_id == 0);
    case OP_VOCAB: return index < n && seq[index].vocab_id == op.vocab_id;
    case OP_FROMLEX: return index < n && ((seq[index].lexindex >> (op.lexicon & 31)) & 1u);
    case OP_NOT: return !eval_pattern(code, pc, seq, n, index);
    default: {  // OP_OR: every alternative is walked (the program counter must end behind the whole expression)
      bool any = false;
      for (uint16_t k = 0; k < op.n; ++k) any = eval_pattern(code, pc, seq, n, index) || any;
      return any;
    }
  }
}
bool parse_u8(const std::string& s, uint8_t* out) {  // str::parse::<u8>()
  size_t i = (!s.empty() && s[0] == '+') ? 1 : 0;
  if (i >= s.size()) return false;
  uint32_t v = 0;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (uint32_t)(s[i] - '0');
    if (v > 255) return false;
  }
  *out = (uint8_t)v;
  return true;
}
std::vector<std::string> split_on(const std::string& s, char sep) {
  std::vector<std::string> out;
  size_t pos = 0;
  for (;;) {
    const size_t k = s.find(sep, pos);
    out.push_back(s.substr(pos, k == std::string::npos ? std::string::npos : k - pos));
    if (k == std::string::npos) break;
    pos = k + 1;
  }
  return out;
}
}  // namespace

// src/lib.rs:658-765
bool HostModel::add_contextrule(const std::string& pattern, float score, const std::vector<std::string>& tag_names,
                                const std::vector<std::string>& tagoffsets, std::string* err) {
  ContextRule rule;
  rule.score = score;
  for (const std::string& expr : split_on(pattern, ';')) {
    rule.start.push_back((uint32_t)rule.code.size());
    std::string perr;
    if (!compile_pattern(*this, expr, &rule.code, &perr)) {
      *err = "Error parsing context rule: " + perr;
      return false;
    }
  }
  const uint8_t plen = (uint8_t)rule.start.size();
  bool empty_tag = false;
  for (const std::string& t : tag_names) {  // (tags are registered before the reference notices an empty one)
    if (t.empty()) empty_tag = true;
    size_t pos = 0;
    while (pos < tags.size() && tags[pos] != t) ++pos;
    if (pos == tags.size()) tags.push_back(t);
    rule.tag.push_back((uint16_t)pos);
  }
  if (empty_tag) {
    *err = "tag is empty";
    return false;
  }
  const char* bad = nullptr;
  for (const std::string& s : tagoffsets) {
    const std::vector<std::string> f = split_on(s, ':');
    uint8_t begin = 0, length = 0;
    if (!f[0].empty() && !parse_u8(f[0], &begin)) bad = "tag offset should be an integer";
    if (f.size() < 2 || f[1].empty())
      length = (uint8_t)(plen - begin);
    else if (!parse_u8(f[1], &length))
      bad = "tag length should be an integer";
    rule.tagoffset.push_back({begin, length});
  }
  if (bad) {
    *err = bad;
    return false;
  }
  while (rule.tagoffset.size() < rule.tag.size()) rule.tagoffset.push_back({0, plen});
  if (!rule.start.empty()) context_rules.push_back(std::move(rule));
  return true;
}

// src/lib.rs:570-656
bool HostModel::read_contextrules(const std::string& filename, std::string* err) {
  std::ifstream f(filename, std::ios::binary);
  if (!f) {
    *err = "cannot open context rules file " + filename;
    return false;
  }
  auto items = [](const std::string& field) {
    std::vector<std::string> out;
    for (const std::string& w : split_on(field, ';')) {
      const std::string t = trim_unicode(w);
      if (!t.empty()) out.push_back(t);
    }
    return out;
  };
  std::string line;
  size_t linenr = 0;
  while (std::getline(f, line)) {
    ++linenr;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    const std::vector<std::string> fields = split_on(line, '\t');
    const std::string where = " (" + filename + ", line " + std::to_string(linenr) + ")";
    if (fields.size() < 2) {
      *err = "Expected at least two columns in context rules file " + filename + ", line " + std::to_string(linenr);
      return false;
    }
    if (fields[0].empty()) continue;
    char* end = nullptr;
    const float score = strtof(fields[1].c_str(), &end);
    if (fields[1].empty() || *end != '\0') {
      *err = "context rule score should be a floating point value above or below 1.0, got " + fields[1] + where;
      return false;
    }
    std::vector<std::string> tag = fields.size() > 2 ? items(fields[2]) : std::vector<std::string>();
    std::vector<std::string> tagoffset = fields.size() > 3 ? items(fields[3]) : std::vector<std::string>();
    if (tag.size() == 1 && tagoffset.empty()) {
      tagoffset.push_back("0:");
    } else if (tag.size() != tagoffset.size()) {
      *err = "Multiple tags are specified for a context rule, expected the same number of tag offsets! (semicolon separated)" + where;
      return false;
    }
    std::string rerr;
    if (!add_contextrule(fields[0], score, tag, tagoffset, &rerr)) {
      *err = "Error adding context rule: " + rerr + where;
      return false;
    }
  }
  return true;
}

// ---- the sequence ----------------------------------------------------------------------------------------------------
namespace {
struct Arc {
  uint32_t from, to, seg;
  int32_t variant;  // -1 = out of vocabulary, -2 = epsilon (fail-safe)
  float cost;
};
struct Partial {  // one of the best partial paths ending in a state
  float cost;
  uint32_t arc;   // last arc (index into the arc list, which is in tie order)
  uint32_t rank;  // which of the source state's partial paths it extends
};
struct HeapItem {
  float cost;
  uint32_t from, arc, rank;
  bool operator>(const HeapItem& o) const {
    if (cost != o.cost) return cost > o.cost;
    if (from != o.from) return from > o.from;
    if (arc != o.arc) return arc > o.arc;
    return rank > o.rank;
  }
};
struct PatternHit {  // PatternMatchResult, src/search.rs:367-372
  float score;
  int32_t tag;
  uint8_t seqnr;
};
}  // namespace

bool most_likely_sequence_full(const HostModel* hm, const std::string& text, const Boundary* bounds, size_t nbounds, size_t end_offset,
                               const SegmentSpan* segs, size_t nsegs, const BatchVariants& variants, const SequenceWeights& w,
                               std::vector<SequenceStep>* out, std::vector<StepTags>* out_tags) {
  out->clear();
  if (out_tags) out_tags->clear();
  const bool use_lm = hm && hm->have_lm() && w.lm_weight > 0.0f;  // :2336
  const bool use_rules = hm && !hm->context_rules.empty();        // :2345
  bool ids_missing = false;  // (a batch without a single variant needs no ids)
  if (!variants.vocab_id)
    for (size_t k = 0; k < nsegs; ++k) ids_missing = ids_missing || variants.count[k] > 0;
  if ((!use_lm && !use_rules) || ids_missing) {
    // nothing but the variant model to weigh: the best of the max_seq shortest paths is the shortest path
    return most_likely_sequence(bounds, nbounds, end_offset, segs, nsegs, variants, out);
  }
  const size_t nstates = nbounds + 1;  // 0 = start, 1 + i = boundary i
  auto state_ending_at = [&](size_t pos) -> long {
    size_t lo = 0, hi = nbounds;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (bounds[mid].end < pos) lo = mid + 1; else hi = mid;
    }
    return (lo < nbounds && bounds[lo].end == pos) ? (long)lo : -1;
  };
  auto state_starting_at = [&](size_t pos) -> long {
    size_t lo = 0, hi = nbounds;
    while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (bounds[mid].begin < pos) lo = mid + 1; else hi = mid;
    }
    return (lo < nbounds && bounds[lo].begin == pos) ? (long)lo : -1;
  };
  std::vector<Arc> arcs;
  size_t labelled = 0;
  for (size_t k = 0; k < nsegs; ++k) {  // :2133-2246
    const long next = state_starting_at(segs[k].end);
    if (next < 0) continue;
    const long prev = state_ending_at(segs[k].begin);
    const long n = prev >= 0 ? next - prev : next + 1;
    const uint32_t from = prev >= 0 ? (uint32_t)prev + 1 : 0;
    const uint32_t cnt = variants.count[k];
    for (uint32_t j = 0; j < cnt; ++j)
      arcs.push_back(Arc{from, (uint32_t)next + 1, (uint32_t)k, (int32_t)j, (float)n + (1.0f - (float)variants.score[variants.first[k] + j])});
    labelled += cnt;
    if (cnt == 0 && n == 1) {
      arcs.push_back(Arc{from, (uint32_t)next + 1, (uint32_t)k, -1, (float)n + 1.0f});
      ++labelled;
    }
  }
  if (labelled == 0) return false;  // :2261-2267
  for (size_t i = 0; i < nbounds; ++i) arcs.push_back(Arc{(uint32_t)i, (uint32_t)i + 1, 0, -2, 100.0f});  // :2249-2259
  // incoming arcs per state (ascending arc index inside a state's list)
  std::vector<uint32_t> in_first(nstates + 1, 0), in_arc(arcs.size());
  for (const Arc& a : arcs) ++in_first[a.to + 1];
  for (size_t s = 0; s < nstates; ++s) in_first[s + 1] += in_first[s];
  {
    std::vector<uint32_t> cur(in_first.begin(), in_first.end() - 1);
    for (uint32_t a = 0; a < arcs.size(); ++a) in_arc[cur[arcs[a].to]++] = a;
  }
  const size_t K = (size_t)std::min<uint64_t>(w.max_seq, 1u << 20);
  if (K == 0) return false;
  // the K best partial paths per state: K-way merge of the predecessors' lists
  std::vector<std::vector<Partial>> best(nstates);
  best[0].push_back(Partial{0.0f, 0xFFFFFFFFu, 0});
  std::priority_queue<HeapItem, std::vector<HeapItem>, std::greater<HeapItem>> heap;
  for (size_t t = 1; t < nstates; ++t) {
    for (uint32_t p = in_first[t]; p < in_first[t + 1]; ++p) {
      const Arc& a = arcs[in_arc[p]];
      if (!best[a.from].empty()) heap.push(HeapItem{best[a.from][0].cost + a.cost, a.from, in_arc[p], 0});
    }
    std::vector<Partial>& list = best[t];
    while (!heap.empty() && list.size() < K) {
      const HeapItem h = heap.top();
      heap.pop();
      list.push_back(Partial{h.cost, h.arc, h.rank});
      const std::vector<Partial>& src = best[h.from];
      if (h.rank + 1 < src.size()) heap.push(HeapItem{src[h.rank + 1].cost + arcs[h.arc].cost, h.from, h.arc, h.rank + 1});
    }
    while (!heap.empty()) heap.pop();
  }
  // final states (:2113-2124) and their merged list
  struct Final {
    float cost;
    uint32_t state, rank;
  };
  std::vector<Final> finals;
  for (size_t i = 0; i < nbounds; ++i)
    if (bounds[i].begin == end_offset || bounds[i].end == end_offset)
      for (uint32_t r = 0; r < best[i + 1].size(); ++r) finals.push_back(Final{best[i + 1][r].cost, (uint32_t)i + 1, r});
  std::sort(finals.begin(), finals.end(), [](const Final& a, const Final& b) {
    if (a.cost != b.cost) return a.cost < b.cost;
    if (a.state != b.state) return a.state < b.state;
    return a.rank < b.rank;
  });
  if (finals.size() > K) finals.resize(K);
  if (finals.empty()) return false;

  // per boundary of the batch: its tokens for the language model (:2603-2620), resolved once
  std::vector<std::vector<int64_t>> boundary_tokens;
  if (use_lm) {
    boundary_tokens.resize(nbounds);
    for (size_t i = 0; i < nbounds; ++i) {
      const std::string t = trim_unicode(text.substr(bounds[i].begin, bounds[i].end - bounds[i].begin));
      if (t.empty()) continue;
      auto it = hm->encoder.find(t);
      if (it == hm->encoder.end()) {
        boundary_tokens[i].push_back(-1);
      } else {
        uint32_t tk[5];
        unsigned n;
        if (hm->entry_tokens(it->second, tk, &n))
          for (unsigned q = 0; q < n; ++q) boundary_tokens[i].push_back(tk[q]);
      }
    }
  }
  struct Candidate {
    std::vector<uint32_t> path;  // arcs, start -> final, fail-safe arcs left out
    float variant_cost;
    double perplexity = 0.0, context_score = 1.0;
    std::vector<std::vector<PatternHit>> hits;
  };
  std::vector<Candidate> cands(finals.size());
  double best_perplexity = 999999.0, best_context = 0.0;  // :2322-2324
  float best_cost = (float)(nbounds - 1) * 2.0f;
  std::vector<int64_t> tokens;
  std::vector<SeqItem> seq;
  for (size_t c = 0; c < finals.size(); ++c) {
    Candidate& cd = cands[c];
    cd.variant_cost = finals[c].cost;
    for (uint32_t st = finals[c].state, rank = finals[c].rank; st != 0;) {
      const Partial& p = best[st][rank];
      if (arcs[p.arc].variant != -2) cd.path.push_back(p.arc);
      st = arcs[p.arc].from;
      rank = p.rank;
    }
    std::reverse(cd.path.begin(), cd.path.end());
    seq.clear();
    for (uint32_t a : cd.path) {
      const Arc& arc = arcs[a];
      const uint64_t vid = arc.variant >= 0 ? variants.vocab_id[variants.first[arc.seg] + arc.variant] : 0;
      const uint32_t lex = (vid != 0 && vid < hm->decoder.size()) ? hm->decoder[vid].lexindex : 0;
      seq.push_back(SeqItem{vid, lex});
    }
    if (use_lm) {  // lm_score, :2570-2640
      tokens.clear();
      tokens.push_back(0);  // <bos>
      for (size_t i = 0; i < cd.path.size(); ++i) {
        if (seq[i].vocab_id == 0) {
          tokens.push_back(-1);
        } else {
          uint32_t tk[5];
          unsigned n;
          if (hm->entry_tokens(seq[i].vocab_id, tk, &n))
            for (unsigned q = 0; q < n; ++q) tokens.push_back(tk[q]);
        }
        const std::vector<int64_t>& bt = boundary_tokens[arcs[cd.path[i]].to - 1];
        tokens.insert(tokens.end(), bt.begin(), bt.end());
      }
      tokens.push_back(1);  // <eos>
      float logprob;
      hm->lm_score_tokens(tokens.data(), tokens.size(), &logprob, &cd.perplexity);
      if (cd.perplexity < best_perplexity) best_perplexity = cd.perplexity;
    }
    if (use_rules) {  // test_context_rules, :2501-2566
      cd.hits.assign(seq.size(), {});
      bool found = false;
      for (size_t begin = 0; begin < seq.size(); ++begin) {
        for (const ContextRule& rule : hm->context_rules) {
          const size_t plen = rule.start.size();
          if (begin + plen > seq.size()) continue;
          bool ok = true;
          for (size_t cursor = 0; cursor < plen && ok; ++cursor) {
            size_t pc = rule.start[cursor];
            ok = cd.hits[begin + cursor].empty() && eval_pattern(rule.code.data(), &pc, seq.data(), seq.size(), begin + cursor);
          }
          if (!ok) continue;
          found = true;
          for (size_t cursor = 0; cursor < plen; ++cursor) {
            std::vector<PatternHit>& h = cd.hits[begin + cursor];
            h.clear();
            if (rule.tag.empty()) {
              h.push_back(PatternHit{rule.score, -1, (uint8_t)cursor});
            } else {
              const size_t nt = std::min(rule.tag.size(), rule.tagoffset.size());
              for (size_t q = 0; q < nt; ++q) {
                const unsigned b = rule.tagoffset[q].first, l = rule.tagoffset[q].second, cu = (uint8_t)cursor;
                if (cu >= b && cu < b + l) h.push_back(PatternHit{rule.score, (int32_t)rule.tag[q], (uint8_t)(cu - b)});
              }
            }
          }
        }
      }
      if (found) {
        float sum = 0.0f;
        for (const auto& h : cd.hits) sum += h.empty() ? 1.0f : h[0].score;
        cd.context_score = (double)sum / (double)seq.size();
      }
    }
    if (cd.variant_cost < best_cost) best_cost = cd.variant_cost;
    if (cd.context_score > best_context) best_context = cd.context_score;
  }
  // :2381-2425
  const bool lm_in_sum = hm->have_lm() && w.lm_weight != 0.0f;
  const bool shortcut = !lm_in_sum && (!use_rules || w.contextrules_weight == 0.0f);
  const double wl = w.lm_weight, wv = w.variantmodel_weight, wc = w.contextrules_weight;
  const Candidate* chosen = nullptr;
  double chosen_score = -99999999.0;
  for (const Candidate& cd : cands) {
    const double norm_lm = use_lm ? std::log(best_perplexity / cd.perplexity) : 0.0;
    const double norm_variant = std::log((double)best_cost / (double)cd.variant_cost);
    const double norm_context = std::log(cd.context_score / best_context);
    const double score = shortcut ? norm_variant : (wl * norm_lm + wv * norm_variant + wc * norm_context) / (wl + wv + wc);
    if (score > chosen_score || !chosen) {
      chosen_score = score;
      chosen = &cd;
    }
  }
  for (size_t i = 0; i < chosen->path.size(); ++i) {
    const Arc& arc = arcs[chosen->path[i]];
    out->push_back(SequenceStep{arc.seg, arc.variant});
    if (out_tags) {
      StepTags st;
      if (use_rules)
        for (const PatternHit& h : chosen->hits[i])
          if (h.tag >= 0) {
            st.tag.push_back((uint16_t)h.tag);
            st.seqnr.push_back(h.seqnr);
          }
      out_tags->push_back(std::move(st));
    }
  }
  return true;
}

}  // namespace anl
