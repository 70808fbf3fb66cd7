// editscript.cpp -- host post-pass for confusable rescoring.
//
// The reference calls sesdiff::shortest_edit_script(input, candidate, false, false, false)
// (src/lib.rs:1736) and matches the result against its confusable patterns
// (src/confusables.rs:13-129).  sesdiff 0.3.1 / dissimilar are third-party crates that are NOT
// vendored in the reference tree (Cargo.toml:29), so this file restates the published algorithm
// they implement -- Myers' bisecting diff with the diff-match-patch clean-up passes (merge,
// semantic, lossless semantic shift, overlap extraction) over Unicode scalar values.  Parity with
// the crates is pinned only by the reference's four confusable tests (tests/main.rs:914-1020).
//
// This runs on the host by design: at most max_matches short string pairs per query, pointer-chasing
// string work with no data parallelism worth a kernel.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "host_model.h"
#include "unicode_tables.h"

namespace anl {
namespace {

typedef std::u32string Str;
enum Op { DEL = -1, EQ = 0, INS = 1 };
struct Chunk {
  Op op;
  Str text;
};
typedef std::vector<Chunk> Script;

Str decode(const std::string& s) {
  Str out;
  size_t i = 0;
  while (i < s.size()) {
    unsigned char c = (unsigned char)s[i];
    unsigned l = c < 0x80 ? 1 : ((c & 0xE0) == 0xC0 ? 2 : ((c & 0xF0) == 0xE0 ? 3 : ((c & 0xF8) == 0xF0 ? 4 : 1)));
    if (i + l > s.size()) l = (unsigned)(s.size() - i);
    char32_t cp = l == 1 ? c : (c & (0xFFu >> (l + 1)));
    for (unsigned k = 1; k < l; ++k) cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3F);
    out.push_back(cp);
    i += l;
  }
  return out;
}
std::string encode(const Str& v) {
  std::string out;
  for (char32_t cp : v) {
    if (cp < 0x80) {
      out += (char)cp;
    } else if (cp < 0x800) {
      out += (char)(0xC0 | (cp >> 6));
      out += (char)(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
      out += (char)(0xE0 | (cp >> 12));
      out += (char)(0x80 | ((cp >> 6) & 0x3F));
      out += (char)(0x80 | (cp & 0x3F));
    } else {
      out += (char)(0xF0 | (cp >> 18));
      out += (char)(0x80 | ((cp >> 12) & 0x3F));
      out += (char)(0x80 | ((cp >> 6) & 0x3F));
      out += (char)(0x80 | (cp & 0x3F));
    }
  }
  return out;
}

size_t prefix_len(const Str& a, const Str& b) {
  size_t n = std::min(a.size(), b.size()), i = 0;
  while (i < n && a[i] == b[i]) ++i;
  return i;
}
size_t suffix_len(const Str& a, const Str& b) {
  size_t n = std::min(a.size(), b.size()), i = 0;
  while (i < n && a[a.size() - 1 - i] == b[b.size() - 1 - i]) ++i;
  return i;
}
bool ends_with(const Str& s, const Str& t) { return s.size() >= t.size() && s.compare(s.size() - t.size(), t.size(), t) == 0; }
bool starts_with(const Str& s, const Str& t) { return s.size() >= t.size() && s.compare(0, t.size(), t) == 0; }

// length of the longest suffix of `a` that is a prefix of `b`
size_t overlap_len(Str a, Str b) {
  if (a.empty() || b.empty()) return 0;
  if (a.size() > b.size())
    a = a.substr(a.size() - b.size());
  else if (a.size() < b.size())
    b = b.substr(0, a.size());
  const size_t n = a.size();
  if (a == b) return n;
  size_t best = 0, len = 1;
  for (;;) {
    Str pat = a.substr(n - len);
    size_t found = b.find(pat);
    if (found == Str::npos) return best;
    len += found;
    if (found == 0 || a.substr(n - len) == b.substr(0, len)) {
      best = len;
      ++len;
    }
    if (len > n) return best;
  }
}

void merge_pass(Script& d);
Script diff(const Str& a, const Str& b);

// Myers O(ND) middle snake, then recurse on both halves.
Script bisect(const Str& a, const Str& b) {
  const long n = (long)a.size(), m = (long)b.size();
  const long maxd = (n + m + 1) / 2, off = maxd, vlen = 2 * maxd;
  std::vector<long> vf(vlen, -1), vr(vlen, -1);
  vf[off + 1] = 0;
  vr[off + 1] = 0;
  const long delta = n - m;
  const bool odd = (delta % 2) != 0;
  long fs = 0, fe = 0, rs = 0, re = 0;
  auto split = [&](long x, long y) {
    Script left = diff(a.substr(0, x), b.substr(0, y));
    Script right = diff(a.substr(x), b.substr(y));
    left.insert(left.end(), right.begin(), right.end());
    return left;
  };
  for (long d = 0; d < maxd; ++d) {
    for (long k = -d + fs; k <= d - fe; k += 2) {
      const long ko = off + k;
      long x = (k == -d || (k != d && vf[ko - 1] < vf[ko + 1])) ? vf[ko + 1] : vf[ko - 1] + 1;
      long y = x - k;
      while (x < n && y < m && a[x] == b[y]) {
        ++x;
        ++y;
      }
      vf[ko] = x;
      if (x > n) {
        fe += 2;
      } else if (y > m) {
        fs += 2;
      } else if (odd) {
        const long ro = off + delta - k;
        if (ro >= 0 && ro < vlen && vr[ro] != -1 && x >= n - vr[ro]) return split(x, y);
      }
    }
    for (long k = -d + rs; k <= d - re; k += 2) {
      const long ko = off + k;
      long x = (k == -d || (k != d && vr[ko - 1] < vr[ko + 1])) ? vr[ko + 1] : vr[ko - 1] + 1;
      long y = x - k;
      while (x < n && y < m && a[n - x - 1] == b[m - y - 1]) {
        ++x;
        ++y;
      }
      vr[ko] = x;
      if (x > n) {
        re += 2;
      } else if (y > m) {
        rs += 2;
      } else if (!odd) {
        const long fo = off + delta - k;
        if (fo >= 0 && fo < vlen && vf[fo] != -1) {
          const long x1 = vf[fo], y1 = off + x1 - fo;
          if (x1 >= n - x) return split(x1, y1);
        }
      }
    }
  }
  return Script{{DEL, a}, {INS, b}};
}

Script middle(const Str& a, const Str& b) {
  if (a.empty() && b.empty()) return {};
  if (a.empty()) return Script{{INS, b}};
  if (b.empty()) return Script{{DEL, a}};
  const Str& lng = a.size() > b.size() ? a : b;
  const Str& sht = a.size() > b.size() ? b : a;
  size_t at = lng.find(sht);
  if (at != Str::npos) {
    const Op op = a.size() > b.size() ? DEL : INS;
    return Script{{op, lng.substr(0, at)}, {EQ, sht}, {op, lng.substr(at + sht.size())}};
  }
  if (sht.size() == 1) return Script{{DEL, a}, {INS, b}};
  return bisect(a, b);
}

Script diff(const Str& a, const Str& b) {
  const size_t p = prefix_len(a, b);
  const Str a1 = a.substr(p), b1 = b.substr(p);
  const size_t s = suffix_len(a1, b1);
  Script out = middle(a1.substr(0, a1.size() - s), b1.substr(0, b1.size() - s));
  if (p) out.insert(out.begin(), Chunk{EQ, a.substr(0, p)});
  if (s) out.push_back(Chunk{EQ, a1.substr(a1.size() - s)});
  merge_pass(out);
  return out;
}

// Reorder and merge like edit sections; factor out common affixes; slide single edits.
void merge_pass(Script& d) {
  bool again = true;
  while (again) {
    d.push_back(Chunk{EQ, Str()});
    size_t i = 0, ndel = 0, nins = 0;
    Str tdel, tins;
    while (i < d.size()) {
      if (d[i].op == INS) {
        ++nins;
        tins += d[i].text;
        ++i;
      } else if (d[i].op == DEL) {
        ++ndel;
        tdel += d[i].text;
        ++i;
      } else {
        if (ndel + nins > 1) {
          if (ndel && nins) {
            size_t c = prefix_len(tins, tdel);
            if (c) {
              const size_t before = i - ndel - nins;
              if (before > 0 && d[before - 1].op == EQ) {
                d[before - 1].text += tins.substr(0, c);
              } else {
                d.insert(d.begin(), Chunk{EQ, tins.substr(0, c)});
                ++i;
              }
              tins = tins.substr(c);
              tdel = tdel.substr(c);
            }
            c = suffix_len(tins, tdel);
            if (c) {
              d[i].text = tins.substr(tins.size() - c) + d[i].text;
              tins = tins.substr(0, tins.size() - c);
              tdel = tdel.substr(0, tdel.size() - c);
            }
          }
          i -= ndel + nins;
          d.erase(d.begin() + i, d.begin() + i + ndel + nins);
          if (!tdel.empty()) d.insert(d.begin() + i++, Chunk{DEL, tdel});
          if (!tins.empty()) d.insert(d.begin() + i++, Chunk{INS, tins});
          ++i;
        } else if (i > 0 && d[i - 1].op == EQ) {
          d[i - 1].text += d[i].text;
          d.erase(d.begin() + i);
        } else {
          ++i;
        }
        ndel = nins = 0;
        tdel.clear();
        tins.clear();
      }
    }
    if (d.back().text.empty()) d.pop_back();
    again = false;
    for (size_t k = 1; k + 1 < d.size(); ++k) {
      if (d[k - 1].op != EQ || d[k + 1].op != EQ) continue;
      if (ends_with(d[k].text, d[k - 1].text)) {
        d[k].text = d[k - 1].text + d[k].text.substr(0, d[k].text.size() - d[k - 1].text.size());
        d[k + 1].text = d[k - 1].text + d[k + 1].text;
        d.erase(d.begin() + k - 1);
        again = true;
      } else if (starts_with(d[k].text, d[k + 1].text)) {
        d[k - 1].text += d[k + 1].text;
        d[k].text = d[k].text.substr(d[k + 1].text.size()) + d[k + 1].text;
        d.erase(d.begin() + k + 1);
        again = true;
      }
    }
  }
}

bool cp_space(char32_t c) {
  return c == ' ' || (c >= 9 && c <= 13) || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
         c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
bool cp_alnum(char32_t c) { return anl_unicode::is_alphabetic(c) || (c >= '0' && c <= '9'); }

// Boundary quality between two strings (6 = edge ... 0 = inside a word).
int boundary_score(const Str& one, const Str& two) {
  if (one.empty() || two.empty()) return 6;
  const char32_t c1 = one.back(), c2 = two.front();
  const bool na1 = !cp_alnum(c1), na2 = !cp_alnum(c2);
  const bool ws1 = na1 && cp_space(c1), ws2 = na2 && cp_space(c2);
  const bool lb1 = ws1 && (c1 == '\n' || c1 == '\r'), lb2 = ws2 && (c2 == '\n' || c2 == '\r');
  auto tail_blank = [](const Str& s) {
    const size_t n = s.size();
    return (n >= 2 && s[n - 1] == '\n' && s[n - 2] == '\n') ||
           (n >= 3 && s[n - 1] == '\n' && s[n - 2] == '\r' && s[n - 3] == '\n');
  };
  auto head_blank = [](const Str& s) {
    const size_t n = s.size();
    if (n >= 2 && s[0] == '\n' && s[1] == '\n') return true;
    if (n >= 3 && s[0] == '\n' && s[1] == '\r' && s[2] == '\n') return true;
    if (n >= 3 && s[0] == '\r' && s[1] == '\n' && s[2] == '\n') return true;
    return n >= 4 && s[0] == '\r' && s[1] == '\n' && s[2] == '\r' && s[3] == '\n';
  };
  if ((lb1 && tail_blank(one)) || (lb2 && head_blank(two))) return 5;
  if (lb1 || lb2) return 4;
  if (na1 && !ws1 && ws2) return 3;
  if (ws1 || ws2) return 2;
  if (na1 || na2) return 1;
  return 0;
}

// Slide an edit that is surrounded by equalities sideways to the best boundary.
void lossless_shift(Script& d) {
  for (size_t k = 1; k + 1 < d.size(); ++k) {
    if (d[k - 1].op != EQ || d[k + 1].op != EQ) continue;
    Str e1 = d[k - 1].text, ed = d[k].text, e2 = d[k + 1].text;
    const size_t c = suffix_len(e1, ed);
    if (c) {
      const Str tail = ed.substr(ed.size() - c);
      e1 = e1.substr(0, e1.size() - c);
      ed = tail + ed.substr(0, ed.size() - c);
      e2 = tail + e2;
    }
    Str b1 = e1, bd = ed, b2 = e2;
    int best = boundary_score(e1, ed) + boundary_score(ed, e2);
    while (!ed.empty() && !e2.empty() && ed[0] == e2[0]) {
      e1 += ed[0];
      ed = ed.substr(1) + e2[0];
      e2 = e2.substr(1);
      const int sc = boundary_score(e1, ed) + boundary_score(ed, e2);
      if (sc >= best) {
        best = sc;
        b1 = e1;
        bd = ed;
        b2 = e2;
      }
    }
    if (d[k - 1].text != b1) {
      if (!b1.empty()) {
        d[k - 1].text = b1;
      } else {
        d.erase(d.begin() + k - 1);
        --k;
      }
      d[k].text = bd;
      if (!b2.empty()) {
        d[k + 1].text = b2;
      } else {
        d.erase(d.begin() + k + 1);
        --k;
      }
    }
  }
}

// Remove equalities that are no longer than the edits on both of their sides.
void semantic_pass(Script& d) {
  bool changed = false;
  std::vector<size_t> eqs;
  bool have = false;
  Str lasteq;
  long i = 0;
  size_t ins1 = 0, del1 = 0, ins2 = 0, del2 = 0;
  while (i < (long)d.size()) {
    if (d[i].op == EQ) {
      eqs.push_back((size_t)i);
      ins1 = ins2;
      del1 = del2;
      ins2 = del2 = 0;
      lasteq = d[i].text;
      have = true;
    } else {
      (d[i].op == INS ? ins2 : del2) += d[i].text.size();
      if (have && lasteq.size() <= std::max(ins1, del1) && lasteq.size() <= std::max(ins2, del2)) {
        const size_t at = eqs.back();
        d.insert(d.begin() + at, Chunk{DEL, lasteq});
        d[at + 1].op = INS;
        eqs.pop_back();
        if (!eqs.empty()) eqs.pop_back();
        i = eqs.empty() ? -1 : (long)eqs.back();
        ins1 = del1 = ins2 = del2 = 0;
        have = false;
        changed = true;
      }
    }
    ++i;
  }
  if (changed) merge_pass(d);
  lossless_shift(d);
  // a deletion followed by an insertion that overlap: pull the overlap out as an equality
  for (size_t k = 1; k < d.size(); ++k) {
    if (d[k - 1].op == DEL && d[k].op == INS) {
      const Str del = d[k - 1].text, ins = d[k].text;
      const size_t o1 = overlap_len(del, ins), o2 = overlap_len(ins, del);
      if (o1 >= o2) {
        if (o1 * 2 >= del.size() || o1 * 2 >= ins.size()) {
          d.insert(d.begin() + k, Chunk{EQ, ins.substr(0, o1)});
          d[k - 1].text = del.substr(0, del.size() - o1);
          d[k + 1].text = ins.substr(o1);
          ++k;
        }
      } else if (o2 * 2 >= del.size() || o2 * 2 >= ins.size()) {
        d.insert(d.begin() + k, Chunk{EQ, del.substr(0, o2)});
        d[k - 1] = Chunk{INS, ins.substr(0, ins.size() - o2)};
        d[k + 1] = Chunk{DEL, del.substr(o2)};
        ++k;
      }
      ++k;
    }
  }
}

}  // namespace

std::vector<EditInstruction> shortest_edit_script(const std::string& src, const std::string& dst) {
  Script d = diff(decode(src), decode(dst));
  semantic_pass(d);
  merge_pass(d);
  std::vector<EditInstruction> out;
  for (const Chunk& c : d)
    if (!c.text.empty()) out.push_back(EditInstruction{(int)c.op, encode(c.text)});
  return out;
}

// Confusable::new (src/confusables.rs:13-45) + sesdiff's pattern syntax: a sequence of `=[..]`,
// `+[..]`, `-[..]`; `|` separates alternatives inside the brackets; leading `^` / trailing `$` anchor.
bool parse_confusable(const std::string& editscript, double weight, Confusable* out) {
  if (editscript.empty()) return false;
  out->strictbegin = editscript.front() == '^';
  out->strictend = editscript.back() == '$';
  out->weight = weight;
  out->script.clear();
  size_t i = out->strictbegin ? 1 : 0;
  const size_t end = editscript.size() - (out->strictend ? 1 : 0);
  if (end < i) return false;
  while (i < end) {
    ConfusableInstr ins;
    switch (editscript[i]) {
      case '=': ins.op = 0; break;
      case '+': ins.op = 1; break;
      case '-': ins.op = -1; break;
      default: return false;
    }
    if (i + 1 >= end || editscript[i + 1] != '[') return false;
    const size_t close = editscript.find(']', i + 2);
    if (close == std::string::npos || close >= end) return false;
    const std::string body = editscript.substr(i + 2, close - i - 2);
    size_t s = 0;
    for (;;) {
      const size_t bar = body.find('|', s);
      ins.options.push_back(body.substr(s, bar == std::string::npos ? std::string::npos : bar - s));
      if (bar == std::string::npos) break;
      s = bar + 1;
    }
    for (const std::string& o : ins.options) ins.options32.push_back(decode(o));
    out->script.push_back(ins);
    i = close + 1;
  }
  out->simple = !out->strictbegin && !out->strictend;
  for (const ConfusableInstr& ci : out->script) out->simple = out->simple && ci.op != 0;
  return !out->script.empty();
}

// Confusable::found_in (src/confusables.rs:47-128)
bool confusable_found_in_views(const Confusable& c, const EditView* ref, size_t nref) {
  const size_t l = c.script.size();
  size_t matches = 0;
  auto sfx = [](const EditView& r, const std::string& t) {
    return r.n >= t.size() && memcmp(r.p + r.n - t.size(), t.data(), t.size()) == 0;
  };
  auto pfx = [](const EditView& r, const std::string& t) { return r.n >= t.size() && memcmp(r.p, t.data(), t.size()) == 0; };
  auto eq = [](const EditView& r, const std::string& t) { return r.n == t.size() && memcmp(r.p, t.data(), r.n) == 0; };
  for (size_t i = 0; i < nref; ++i) {
    if (matches >= l) continue;
    const ConfusableInstr& ins = c.script[matches];
    bool found = false;
    if (ins.op == ref[i].op) {
      for (const std::string& s : ins.options) {
        bool ok;
        if (ins.op != 0)
          ok = sfx(ref[i], s);
        else if (matches == 0 && matches == l - 1)
          ok = eq(ref[i], s);
        else if (matches == 0)
          ok = sfx(ref[i], s);
        else if (matches == l - 1)
          ok = pfx(ref[i], s);
        else
          ok = eq(ref[i], s);
        if (ok) {
          found = true;
          break;
        }
      }
    }
    if (!found) {
      matches = 0;
      if (c.strictbegin) return false;
    } else if (++matches == l) {
      return c.strictend ? i == nref - 1 : true;
    }
  }
  return false;
}

bool confusable_found_in(const Confusable& c, const std::vector<EditInstruction>& ref) {
  EditView small[16];
  std::vector<EditView> big;
  EditView* v = small;
  if (ref.size() > 16) {
    big.resize(ref.size());
    v = big.data();
  }
  for (size_t i = 0; i < ref.size(); ++i) v[i] = EditView{ref[i].op, ref[i].text.data(), ref[i].text.size()};
  return confusable_found_in_views(c, v, ref.size());
}

}  // namespace anl
