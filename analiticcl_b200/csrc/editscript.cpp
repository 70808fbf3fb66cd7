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

// the fixed-capacity implementation, instantiated here for Unicode scalar values (host only)
#define ESF_FN inline
#include "editscript_fixed.h"

namespace anl {
namespace {

// A diff is an ordered partition of both strings, so a script is stored as (op, length) pairs; the
// text of a chunk is implied by its position: EQ consumes `len` scalars of both strings, DEL of the
// source, INS of the destination.  Every pass below works on lengths plus comparisons inside the
// two base arrays: no allocation, no copying of text.
enum Op : int8_t { DEL = -1, EQ = 0, INS = 1 };
struct Seg {
  Op op;
  int len;
};
typedef std::vector<Seg> Script;
typedef const char32_t* Txt;

struct Ctx {
  Txt a;  // source scalars
  Txt b;  // destination scalars
};

size_t decode(const char* s, size_t n, std::vector<char32_t>* out) {
  out->clear();
  size_t i = 0;
  while (i < n) {
    unsigned char c = (unsigned char)s[i];
    unsigned l = c < 0x80 ? 1 : ((c & 0xE0) == 0xC0 ? 2 : ((c & 0xF0) == 0xE0 ? 3 : ((c & 0xF8) == 0xF0 ? 4 : 1)));
    if (i + l > n) l = (unsigned)(n - i);
    char32_t cp = l == 1 ? c : (c & (0xFFu >> (l + 1)));
    for (unsigned k = 1; k < l; ++k) cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3F);
    out->push_back(cp);
    i += l;
  }
  return out->size();
}
std::u32string decode32(const std::string& s) {
  std::vector<char32_t> v;
  decode(s.data(), s.size(), &v);
  return std::u32string(v.begin(), v.end());
}
void append_utf8(std::string* out, Txt t, int n) {
  for (int i = 0; i < n; ++i) {
    const char32_t cp = t[i];
    if (cp < 0x80) {
      *out += (char)cp;
    } else if (cp < 0x800) {
      *out += (char)(0xC0 | (cp >> 6));
      *out += (char)(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
      *out += (char)(0xE0 | (cp >> 12));
      *out += (char)(0x80 | ((cp >> 6) & 0x3F));
      *out += (char)(0x80 | (cp & 0x3F));
    } else {
      *out += (char)(0xF0 | (cp >> 18));
      *out += (char)(0x80 | ((cp >> 12) & 0x3F));
      *out += (char)(0x80 | ((cp >> 6) & 0x3F));
      *out += (char)(0x80 | (cp & 0x3F));
    }
  }
}

inline int common_prefix(Txt x, int nx, Txt y, int ny) {
  const int n = std::min(nx, ny);
  int i = 0;
  while (i < n && x[i] == y[i]) ++i;
  return i;
}
inline int common_suffix(Txt x, int nx, Txt y, int ny) {
  const int n = std::min(nx, ny);
  int i = 0;
  while (i < n && x[nx - 1 - i] == y[ny - 1 - i]) ++i;
  return i;
}
inline bool same(Txt x, Txt y, int n) {
  for (int i = 0; i < n; ++i)
    if (x[i] != y[i]) return false;
  return true;
}
// first occurrence of needle in hay, or -1
inline int find_in(Txt hay, int nh, Txt needle, int nn) {
  if (nn == 0) return 0;
  for (int i = 0; i + nn <= nh; ++i)
    if (same(hay + i, needle, nn)) return i;
  return -1;
}
// length of the longest suffix of x that is a prefix of y
inline int overlap_len(Txt x, int nx, Txt y, int ny) {
  for (int l = std::min(nx, ny); l >= 1; --l)
    if (same(x + nx - l, y, l)) return l;
  return 0;
}

// start positions (in a and b) of segment k of a script whose first segment starts at (a0, b0)
inline void seg_pos(const Script& d, size_t k, int a0, int b0, int* pa, int* pb) {
  for (size_t i = 0; i < k; ++i) {
    if (d[i].op != INS) a0 += d[i].len;
    if (d[i].op != DEL) b0 += d[i].len;
  }
  *pa = a0;
  *pb = b0;
}

void merge_pass(const Ctx& c, Script& d, int a0, int b0);
void diff(const Ctx& c, int alo, int ahi, int blo, int bhi, Script* out);

// Myers O(ND) middle snake on a[alo,ahi) x b[blo,bhi), then recurse on both halves.
void bisect(const Ctx& c, int alo, int ahi, int blo, int bhi, Script* out) {
  const long n = ahi - alo, m = bhi - blo;
  const long maxd = (n + m + 1) / 2, off = maxd, vlen = 2 * maxd;
  std::vector<long> vf(vlen, -1), vr(vlen, -1);
  vf[off + 1] = 0;
  vr[off + 1] = 0;
  const long delta = n - m;
  const bool odd = (delta % 2) != 0;
  long fs = 0, fe = 0, rs = 0, re = 0;
  Txt a = c.a + alo;
  Txt b = c.b + blo;
  for (long dd = 0; dd < maxd; ++dd) {
    for (long k = -dd + fs; k <= dd - fe; k += 2) {
      const long ko = off + k;
      long x = (k == -dd || (k != dd && vf[ko - 1] < vf[ko + 1])) ? vf[ko + 1] : vf[ko - 1] + 1;
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
        if (ro >= 0 && ro < vlen && vr[ro] != -1 && x >= n - vr[ro]) {
          diff(c, alo, alo + (int)x, blo, blo + (int)y, out);
          diff(c, alo + (int)x, ahi, blo + (int)y, bhi, out);
          return;
        }
      }
    }
    for (long k = -dd + rs; k <= dd - re; k += 2) {
      const long ko = off + k;
      long x = (k == -dd || (k != dd && vr[ko - 1] < vr[ko + 1])) ? vr[ko + 1] : vr[ko - 1] + 1;
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
          if (x1 >= n - x) {
            diff(c, alo, alo + (int)x1, blo, blo + (int)y1, out);
            diff(c, alo + (int)x1, ahi, blo + (int)y1, bhi, out);
            return;
          }
        }
      }
    }
  }
  out->push_back(Seg{DEL, (int)n});
  out->push_back(Seg{INS, (int)m});
}

// the diff of two strings without a common prefix or suffix
void middle(const Ctx& c, int alo, int ahi, int blo, int bhi, Script* out) {
  const int n = ahi - alo, m = bhi - blo;
  if (n == 0 && m == 0) return;
  if (n == 0) {
    out->push_back(Seg{INS, m});
    return;
  }
  if (m == 0) {
    out->push_back(Seg{DEL, n});
    return;
  }
  const bool a_longer = n > m;
  const int at = a_longer ? find_in(c.a + alo, n, c.b + blo, m) : find_in(c.b + blo, m, c.a + alo, n);
  if (at >= 0) {
    const Op op = a_longer ? DEL : INS;
    const int lng = a_longer ? n : m, sht = a_longer ? m : n;
    out->push_back(Seg{op, at});  // may be empty, like the restated algorithm
    out->push_back(Seg{EQ, sht});
    out->push_back(Seg{op, lng - at - sht});
    return;
  }
  if (std::min(n, m) == 1) {
    out->push_back(Seg{DEL, n});
    out->push_back(Seg{INS, m});
    return;
  }
  bisect(c, alo, ahi, blo, bhi, out);
}

void diff(const Ctx& c, int alo, int ahi, int blo, int bhi, Script* out) {
  const int p = common_prefix(c.a + alo, ahi - alo, c.b + blo, bhi - blo);
  const int s = common_suffix(c.a + alo + p, ahi - alo - p, c.b + blo + p, bhi - blo - p);
  Script local;
  if (p) local.push_back(Seg{EQ, p});
  middle(c, alo + p, ahi - s, blo + p, bhi - s, &local);
  if (s) local.push_back(Seg{EQ, s});
  merge_pass(c, local, alo, blo);
  out->insert(out->end(), local.begin(), local.end());
}

// Reorder and merge like edit sections; factor out common affixes; slide single edits.
void merge_pass(const Ctx& c, Script& d, int a0, int b0) {
  bool again = true;
  while (again) {
    d.push_back(Seg{EQ, 0});
    size_t i = 0, ndel = 0, nins = 0;
    int dl = 0, il = 0;          // accumulated deletion / insertion lengths of the current run
    int ra = a0, rb = b0;        // positions where the current run starts
    int pa = a0, pb = b0;        // positions of segment i
    while (i < d.size()) {
      if (d[i].op == INS) {
        ++nins;
        il += d[i].len;
        pb += d[i].len;
        ++i;
      } else if (d[i].op == DEL) {
        ++ndel;
        dl += d[i].len;
        pa += d[i].len;
        ++i;
      } else {
        if (ndel + nins > 1) {
          if (ndel && nins) {
            int cp = common_prefix(c.b + rb, il, c.a + ra, dl);
            if (cp) {
              const size_t before = i - ndel - nins;
              if (before > 0 && d[before - 1].op == EQ) {
                d[before - 1].len += cp;
              } else {
                d.insert(d.begin(), Seg{EQ, cp});
                ++i;
              }
              ra += cp;
              rb += cp;
              il -= cp;
              dl -= cp;
            }
            cp = common_suffix(c.b + rb, il, c.a + ra, dl);
            if (cp) {
              d[i].len += cp;
              il -= cp;
              dl -= cp;
            }
          }
          i -= ndel + nins;
          d.erase(d.begin() + i, d.begin() + i + ndel + nins);
          if (dl) d.insert(d.begin() + i++, Seg{DEL, dl});
          if (il) d.insert(d.begin() + i++, Seg{INS, il});
          // positions of the equality now at index i
          pa = ra + dl;
          pb = rb + il;
          pa += d[i].len;
          pb += d[i].len;
          ++i;
        } else if (i > 0 && d[i - 1].op == EQ) {
          d[i - 1].len += d[i].len;
          pa += d[i].len;
          pb += d[i].len;
          d.erase(d.begin() + i);
        } else {
          pa += d[i].len;
          pb += d[i].len;
          ++i;
        }
        ndel = nins = 0;
        dl = il = 0;
        ra = pa;
        rb = pb;
      }
    }
    if (d.back().len == 0) d.pop_back();
    again = false;
    for (size_t k = 1; k + 1 < d.size(); ++k) {
      if (d[k - 1].op != EQ || d[k + 1].op != EQ) continue;
      int ka, kb;
      seg_pos(d, k, a0, b0, &ka, &kb);
      // text of the edit and of its neighbours, all inside the edit's own string
      const bool del = d[k].op == DEL;
      Txt base = del ? c.a : c.b;
      const int pos = del ? ka : kb;
      const int len = d[k].len, lp = d[k - 1].len, ln = d[k + 1].len;
      // the equalities read through the *other* string are identical, so one string suffices
      if (len >= lp && same(base + pos + len - lp, base + pos - lp, lp)) {
        // edit ends with the previous equality: shift the edit left over it
        d[k + 1].len += lp;
        d.erase(d.begin() + k - 1);
        again = true;
      } else if (len >= ln && same(base + pos, base + pos + len, ln)) {
        // edit starts with the next equality: shift the edit right over it
        d[k - 1].len += ln;
        d.erase(d.begin() + k + 1);
        again = true;
      }
    }
  }
}

bool cp_space(char32_t ch) {
  return ch == ' ' || (ch >= 9 && ch <= 13) || ch == 0x85 || ch == 0xA0 || ch == 0x1680 || (ch >= 0x2000 && ch <= 0x200A) ||
         ch == 0x2028 || ch == 0x2029 || ch == 0x202F || ch == 0x205F || ch == 0x3000;
}
bool cp_alnum(char32_t ch) { return anl_unicode::is_alphabetic(ch) || (ch >= '0' && ch <= '9'); }

// Boundary quality between two strings (6 = edge ... 0 = inside a word).
int boundary_score(Txt one, int n1, Txt two, int n2) {
  if (n1 == 0 || n2 == 0) return 6;
  const char32_t c1 = one[n1 - 1], c2 = two[0];
  const bool na1 = !cp_alnum(c1), na2 = !cp_alnum(c2);
  const bool ws1 = na1 && cp_space(c1), ws2 = na2 && cp_space(c2);
  const bool lb1 = ws1 && (c1 == '\n' || c1 == '\r'), lb2 = ws2 && (c2 == '\n' || c2 == '\r');
  auto tail_blank = [](Txt s, int n) {
    return (n >= 2 && s[n - 1] == '\n' && s[n - 2] == '\n') || (n >= 3 && s[n - 1] == '\n' && s[n - 2] == '\r' && s[n - 3] == '\n');
  };
  auto head_blank = [](Txt s, int n) {
    if (n >= 2 && s[0] == '\n' && s[1] == '\n') return true;
    if (n >= 3 && s[0] == '\n' && s[1] == '\r' && s[2] == '\n') return true;
    if (n >= 3 && s[0] == '\r' && s[1] == '\n' && s[2] == '\n') return true;
    return n >= 4 && s[0] == '\r' && s[1] == '\n' && s[2] == '\r' && s[3] == '\n';
  };
  if ((lb1 && tail_blank(one, n1)) || (lb2 && head_blank(two, n2))) return 5;
  if (lb1 || lb2) return 4;
  if (na1 && !ws1 && ws2) return 3;
  if (ws1 || ws2) return 2;
  if (na1 || na2) return 1;
  return 0;
}

// Slide an edit that is surrounded by equalities sideways to the best boundary.
void lossless_shift(const Ctx& c, Script& d) {
  for (size_t k = 1; k + 1 < d.size(); ++k) {
    if (d[k - 1].op != EQ || d[k + 1].op != EQ) continue;
    int ka, kb;
    seg_pos(d, k, 0, 0, &ka, &kb);
    const bool del = d[k].op == DEL;
    Txt base = del ? c.a : c.b;
    int pos = del ? ka : kb;  // start of the edit inside its own string
    const int m = d[k].len;
    int l1 = d[k - 1].len, l2 = d[k + 1].len;
    // the window [pos - l1, pos + m + l2) of `base` is  equality1 | edit | equality2
    const int cs = common_suffix(base + pos - l1, l1, base + pos, m);
    pos -= cs;
    l1 -= cs;
    l2 += cs;
    int best_l1 = l1, best_l2 = l2;
    int best = boundary_score(base + pos - l1, l1, base + pos, m) + boundary_score(base + pos, m, base + pos + m, l2);
    while (m > 0 && l2 > 0 && base[pos] == base[pos + m]) {
      ++pos;
      ++l1;
      --l2;
      const int sc = boundary_score(base + pos - l1, l1, base + pos, m) + boundary_score(base + pos, m, base + pos + m, l2);
      if (sc >= best) {
        best = sc;
        best_l1 = l1;
        best_l2 = l2;
      }
    }
    if (d[k - 1].len != best_l1) {
      // (equal lengths mean equal text here: the window is fixed and only the split moves)
      size_t kk = k;
      if (best_l1) {
        d[kk - 1].len = best_l1;
      } else {
        d.erase(d.begin() + kk - 1);
        --kk;
      }
      if (best_l2) {
        d[kk + 1].len = best_l2;
      } else {
        d.erase(d.begin() + kk + 1);
        --kk;
      }
      k = kk;
    }
  }
}

// Remove equalities that are no longer than the edits on both of their sides.
void semantic_pass(const Ctx& c, Script& d) {
  bool changed = false;
  std::vector<size_t> eqs;
  bool have = false;
  int lasteq = 0;
  long i = 0;
  int ins1 = 0, del1 = 0, ins2 = 0, del2 = 0;
  while (i < (long)d.size()) {
    if (d[i].op == EQ) {
      eqs.push_back((size_t)i);
      ins1 = ins2;
      del1 = del2;
      ins2 = del2 = 0;
      lasteq = d[i].len;
      have = true;
    } else {
      (d[i].op == INS ? ins2 : del2) += d[i].len;
      if (have && lasteq <= std::max(ins1, del1) && lasteq <= std::max(ins2, del2)) {
        const size_t at = eqs.back();
        d.insert(d.begin() + at, Seg{DEL, lasteq});
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
  if (changed) merge_pass(c, d, 0, 0);
  lossless_shift(c, d);
  // a deletion followed by an insertion that overlap: pull the overlap out as an equality
  for (size_t k = 1; k < d.size(); ++k) {
    if (d[k - 1].op == DEL && d[k].op == INS) {
      int ka, kb;
      seg_pos(d, k - 1, 0, 0, &ka, &kb);
      const int dl = d[k - 1].len, il = d[k].len;
      Txt del = c.a + ka;
      Txt ins = c.b + kb;
      const int o1 = overlap_len(del, dl, ins, il), o2 = overlap_len(ins, il, del, dl);
      if (o1 >= o2) {
        if (o1 * 2 >= dl || o1 * 2 >= il) {
          d.insert(d.begin() + k, Seg{EQ, o1});
          d[k - 1].len = dl - o1;
          d[k + 1].len = il - o1;
          ++k;
        }
      } else if (o2 * 2 >= dl || o2 * 2 >= il) {
        d.insert(d.begin() + k, Seg{EQ, o2});
        d[k - 1] = Seg{INS, il - o2};
        d[k + 1] = Seg{DEL, dl - o2};
        ++k;
      }
      ++k;
    }
  }
}

}  // namespace

std::vector<EditInstruction> shortest_edit_script(const std::string& src, const std::string& dst) {
  static thread_local std::vector<char32_t> ua, ub;
  const int na = (int)decode(src.data(), src.size(), &ua), nb = (int)decode(dst.data(), dst.size(), &ub);
  Ctx c{ua.data(), ub.data()};
  Script d;
  d.reserve(16);
  diff(c, 0, na, 0, nb, &d);
  semantic_pass(c, d);
  merge_pass(c, d, 0, 0);
  std::vector<EditInstruction> out;
  out.reserve(d.size());
  int pa = 0, pb = 0;
  for (const Seg& s : d) {
    if (s.len > 0) {
      EditInstruction e;
      e.op = (int)s.op;
      append_utf8(&e.text, s.op == INS ? c.b + pb : c.a + pa, s.len);
      out.push_back(std::move(e));
    }
    if (s.op != INS) pa += s.len;
    if (s.op != DEL) pb += s.len;
  }
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
    for (const std::string& o : ins.options) ins.options32.push_back(decode32(o));
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


// The allocation-free route for short strings (what the host post-pass uses): edit script over Unicode
// scalar values by the fixed-capacity implementation, returned as byte views into the two UTF-8 strings.
// false = outside its limits (more than esf::MAXLEN scalars, internal capacity): use shortest_edit_script.
namespace {
struct UnicodeClass {
  static bool alnum(uint32_t c) { return cp_alnum((char32_t)c); }
  static bool space(uint32_t c) { return cp_space((char32_t)c); }
};
// decodes into scalars + the byte offset of each scalar (n + 1 entries); false if more than MAXLEN scalars
bool decode_with_offsets(const char* s, size_t n, char32_t* out, uint16_t* off, int* count) {
  int c = 0;
  size_t i = 0;
  while (i < n) {
    if (c >= esf::MAXLEN) return false;
    const unsigned char ch = (unsigned char)s[i];
    unsigned l = ch < 0x80 ? 1 : ((ch & 0xE0) == 0xC0 ? 2 : ((ch & 0xF0) == 0xE0 ? 3 : ((ch & 0xF8) == 0xF0 ? 4 : 1)));
    if (i + l > n) l = (unsigned)(n - i);
    char32_t cp = l == 1 ? ch : (ch & (0xFFu >> (l + 1)));
    for (unsigned k = 1; k < l; ++k) cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3F);
    off[c] = (uint16_t)i;
    out[c++] = cp;
    i += l;
  }
  off[c] = (uint16_t)n;
  *count = c;
  return true;
}
}  // namespace

bool edit_views_fixed(const char* src, size_t src_len, const char* dst, size_t dst_len, EditView* out, size_t* nout) {
  if (src_len > 4 * (size_t)esf::MAXLEN || dst_len > 4 * (size_t)esf::MAXLEN) return false;
  char32_t a[esf::MAXLEN], b[esf::MAXLEN];
  uint16_t oa[esf::MAXLEN + 1], ob[esf::MAXLEN + 1];
  int na = 0, nb = 0;
  if (!decode_with_offsets(src, src_len, a, oa, &na) || !decode_with_offsets(dst, dst_len, b, ob, &nb)) return false;
  esf::View v[esf::MAXSEG];
  const int nv = esf::shortest_edit_script_t<UnicodeClass, char32_t>(a, na, b, nb, v);
  if (nv < 0) return false;
  for (int i = 0; i < nv; ++i) {
    const bool ins = v[i].op > 0;
    const uint16_t* o = ins ? ob : oa;
    const char* base = ins ? dst : src;
    out[i] = EditView{v[i].op, base + o[v[i].pos], (size_t)(o[v[i].pos + v[i].len] - o[v[i].pos])};
  }
  *nout = (size_t)nv;
  return true;
}

}  // namespace anl
