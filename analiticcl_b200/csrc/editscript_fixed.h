// editscript_fixed.h -- fixed-capacity, allocation-free edit script + confusable matching.
//
// The same restated algorithm as editscript.cpp (sesdiff::shortest_edit_script as called from
// src/lib.rs:1736: Myers' bisecting diff + the diff-match-patch clean-up passes; then
// Confusable::found_in, src/confusables.rs:47-128), written over plain arrays so that it compiles
// for the device: the confusable kernel (confusables.cu) runs one (input, candidate) pair per thread
// with everything in registers / local memory.  It covers pure-ASCII pairs (bytes are Unicode scalar
// values) of at most MAXLEN characters; anything else -- and any internal capacity overflow -- is
// reported as "not settled" and left to the host post-pass (editscript.cpp), never approximated.
//
// The recursion diff -> bisect -> diff of the host version becomes an explicit frame stack: a
// sub-problem's script is always the tail of the one flat segment array, so its clean-up pass
// (merge_pass on the tail) reproduces the host's post-order of nested merge passes exactly.
// tests/test_editscript_parity.py checks this file (compiled for the host) against editscript.cpp
// and the oracle on random and workload pairs; the GPU parity tests check the device build.
#pragma once
#include <stdint.h>

#include "device_types.h"

#if !defined(ESF_FN)
#if defined(__CUDACC__)
#define ESF_FN __host__ __device__ inline
#else
#define ESF_FN inline
#endif
#endif

namespace anl {
namespace esf {

constexpr int MAXLEN = 64;    // longest string handled (characters == bytes, ASCII only)
constexpr int MAXSEG = 64;    // capacity of the segment array
constexpr int MAXFRAME = 32;  // capacity of the frame stack

constexpr int8_t DEL = -1, EQ = 0, INS = 1;
struct Seg {
  int8_t op;
  uint8_t len;
};
struct Frame {
  uint8_t alo, ahi, blo, bhi;  // stage 0: the sub-problem a[alo,ahi) x b[blo,bhi); stage 1: (alo, blo) = origin
  uint8_t start, suf, stage, pad;
};
template <class T>
struct State {
  const T* a;  // source
  const T* b;  // destination
  int nd;
  bool overflow;
  Seg d[MAXSEG];
};
struct View {  // one instruction of the final script: its text is (op == INS ? b : a)[pos, pos + len)
  int8_t op;
  uint8_t len;
  uint8_t pos;
  uint8_t pad;
};

ESF_FN int imin(int x, int y) { return x < y ? x : y; }
ESF_FN int imax(int x, int y) { return x > y ? x : y; }

template <class T>
ESF_FN void push(State<T>& S, int8_t op, int len) {
  if (S.nd >= MAXSEG) {
    S.overflow = true;
    return;
  }
  S.d[S.nd].op = op;
  S.d[S.nd].len = (uint8_t)len;
  ++S.nd;
}
template <class T>
ESF_FN void insert_at(State<T>& S, int idx, int8_t op, int len) {
  if (S.nd >= MAXSEG) {
    S.overflow = true;
    return;
  }
  for (int k = S.nd; k > idx; --k) S.d[k] = S.d[k - 1];
  S.d[idx].op = op;
  S.d[idx].len = (uint8_t)len;
  ++S.nd;
}
template <class T>
ESF_FN void erase(State<T>& S, int idx, int cnt) {
  for (int k = idx; k + cnt < S.nd; ++k) S.d[k] = S.d[k + cnt];
  S.nd -= cnt;
}

template <class T>
ESF_FN int common_prefix(const T* x, int nx, const T* y, int ny) {
  const int n = imin(nx, ny);
  int i = 0;
  while (i < n && x[i] == y[i]) ++i;
  return i;
}
template <class T>
ESF_FN int common_suffix(const T* x, int nx, const T* y, int ny) {
  const int n = imin(nx, ny);
  int i = 0;
  while (i < n && x[nx - 1 - i] == y[ny - 1 - i]) ++i;
  return i;
}
template <class T>
ESF_FN bool same(const T* x, const T* y, int n) {
  for (int i = 0; i < n; ++i)
    if (x[i] != y[i]) return false;
  return true;
}
template <class T>
ESF_FN int find_in(const T* hay, int nh, const T* needle, int nn) {
  if (nn == 0) return 0;
  for (int i = 0; i + nn <= nh; ++i)
    if (same(hay + i, needle, nn)) return i;
  return -1;
}
template <class T>
ESF_FN int overlap_len(const T* x, int nx, const T* y, int ny) {
  for (int l = imin(nx, ny); l >= 1; --l)
    if (same(x + nx - l, y, l)) return l;
  return 0;
}
// start positions (in a and b) of segment k of the sub-script that starts at index s0 / origin (a0, b0)
template <class T>
ESF_FN void seg_pos(const State<T>& S, int s0, int k, int a0, int b0, int* pa, int* pb) {
  for (int i = s0; i < k; ++i) {
    if (S.d[i].op != INS) a0 += S.d[i].len;
    if (S.d[i].op != DEL) b0 += S.d[i].len;
  }
  *pa = a0;
  *pb = b0;
}

// Reorder and merge like edit sections, factor out common affixes, slide single edits -- on the
// tail [s0, nd) of the segment array (cf. merge_pass in editscript.cpp).
template <class T>
ESF_FN void merge_pass(State<T>& S, int s0, int a0, int b0) {
  bool again = true;
  while (again && !S.overflow) {
    push(S, EQ, 0);
    if (S.overflow) return;
    int i = s0, ndel = 0, nins = 0, dl = 0, il = 0;
    int ra = a0, rb = b0, pa = a0, pb = b0;
    while (i < S.nd && !S.overflow) {
      const Seg cur = S.d[i];
      if (cur.op == INS) {
        ++nins;
        il += cur.len;
        pb += cur.len;
        ++i;
      } else if (cur.op == DEL) {
        ++ndel;
        dl += cur.len;
        pa += cur.len;
        ++i;
      } else {
        if (ndel + nins > 1) {
          if (ndel && nins) {
            int cp = common_prefix(S.b + rb, il, S.a + ra, dl);
            if (cp) {
              const int before = i - ndel - nins;
              if (before > s0 && S.d[before - 1].op == EQ) {
                S.d[before - 1].len = (uint8_t)(S.d[before - 1].len + cp);
              } else {
                insert_at(S, s0, EQ, cp);
                ++i;
              }
              ra += cp;
              rb += cp;
              il -= cp;
              dl -= cp;
            }
            cp = common_suffix(S.b + rb, il, S.a + ra, dl);
            if (cp) {
              S.d[i].len = (uint8_t)(S.d[i].len + cp);
              il -= cp;
              dl -= cp;
            }
          }
          i -= ndel + nins;
          erase(S, i, ndel + nins);
          if (dl) insert_at(S, i++, DEL, dl);
          if (il) insert_at(S, i++, INS, il);
          if (S.overflow) return;
          pa = ra + dl + S.d[i].len;
          pb = rb + il + S.d[i].len;
          ++i;
        } else if (i > s0 && S.d[i - 1].op == EQ) {
          S.d[i - 1].len = (uint8_t)(S.d[i - 1].len + cur.len);
          pa += cur.len;
          pb += cur.len;
          erase(S, i, 1);
        } else {
          pa += cur.len;
          pb += cur.len;
          ++i;
        }
        ndel = nins = 0;
        dl = il = 0;
        ra = pa;
        rb = pb;
      }
    }
    if (S.overflow) return;
    if (S.nd > s0 && S.d[S.nd - 1].len == 0) --S.nd;
    again = false;
    for (int k = s0 + 1; k + 1 < S.nd; ++k) {
      if (S.d[k - 1].op != EQ || S.d[k + 1].op != EQ) continue;
      int ka, kb;
      seg_pos(S, s0, k, a0, b0, &ka, &kb);
      const bool del = S.d[k].op == DEL;
      const T* base = del ? S.a : S.b;
      const int pos = del ? ka : kb;
      const int len = S.d[k].len, lp = S.d[k - 1].len, ln = S.d[k + 1].len;
      if (len >= lp && same(base + pos + len - lp, base + pos - lp, lp)) {
        S.d[k + 1].len = (uint8_t)(S.d[k + 1].len + lp);
        erase(S, k - 1, 1);
        again = true;
      } else if (len >= ln && same(base + pos, base + pos + len, ln)) {
        S.d[k - 1].len = (uint8_t)(S.d[k - 1].len + ln);
        erase(S, k + 1, 1);
        again = true;
      }
    }
  }
}

// Myers O(ND) middle snake on a[alo,ahi) x b[blo,bhi): true + the split point (relative), or false
// when the strings share nothing (cf. bisect in editscript.cpp).
template <class T>
ESF_FN bool bisect_split(const State<T>& S, int alo, int ahi, int blo, int bhi, int* sx, int* sy) {
  const int n = ahi - alo, m = bhi - blo;
  const int maxd = (n + m + 1) / 2, off = maxd, vlen = 2 * maxd;
  int8_t vf[MAXLEN * 2 + 4], vr[MAXLEN * 2 + 4];
  for (int i = 0; i < vlen + 2; ++i) vf[i] = vr[i] = -1;
  vf[off + 1] = 0;
  vr[off + 1] = 0;
  const int delta = n - m;
  const bool odd = (delta % 2) != 0;
  int fs = 0, fe = 0, rs = 0, re = 0;
  const T* a = S.a + alo;
  const T* b = S.b + blo;
  for (int dd = 0; dd < maxd; ++dd) {
    for (int k = -dd + fs; k <= dd - fe; k += 2) {
      const int ko = off + k;
      int x = (k == -dd || (k != dd && vf[ko - 1] < vf[ko + 1])) ? vf[ko + 1] : vf[ko - 1] + 1;
      int y = x - k;
      while (x < n && y < m && a[x] == b[y]) {
        ++x;
        ++y;
      }
      vf[ko] = (int8_t)x;
      if (x > n) {
        fe += 2;
      } else if (y > m) {
        fs += 2;
      } else if (odd) {
        const int ro = off + delta - k;
        if (ro >= 0 && ro < vlen && vr[ro] != -1 && x >= n - vr[ro]) {
          *sx = x;
          *sy = y;
          return true;
        }
      }
    }
    for (int k = -dd + rs; k <= dd - re; k += 2) {
      const int ko = off + k;
      int x = (k == -dd || (k != dd && vr[ko - 1] < vr[ko + 1])) ? vr[ko + 1] : vr[ko - 1] + 1;
      int y = x - k;
      while (x < n && y < m && a[n - x - 1] == b[m - y - 1]) {
        ++x;
        ++y;
      }
      vr[ko] = (int8_t)x;
      if (x > n) {
        re += 2;
      } else if (y > m) {
        rs += 2;
      } else if (!odd) {
        const int fo = off + delta - k;
        if (fo >= 0 && fo < vlen && vf[fo] != -1) {
          const int x1 = vf[fo], y1 = off + x1 - fo;
          if (x1 >= n - x) {
            *sx = x1;
            *sy = y1;
            return true;
          }
        }
      }
    }
  }
  return false;
}

// diff(a[0,na), b[0,nb)) with the nested clean-up passes of the recursive formulation.
template <class T>
ESF_FN void diff_main(State<T>& S, int na, int nb) {
  Frame st[MAXFRAME];
  int sp = 0;
  st[sp++] = Frame{0, (uint8_t)na, 0, (uint8_t)nb, 0, 0, 0, 0};
  while (sp > 0 && !S.overflow) {
    const Frame f = st[--sp];
    if (f.stage == 1) {
      if (f.suf) push(S, EQ, f.suf);
      merge_pass(S, f.start, f.alo, f.blo);
      continue;
    }
    const int alo = f.alo, ahi = f.ahi, blo = f.blo, bhi = f.bhi;
    const int p = common_prefix(S.a + alo, ahi - alo, S.b + blo, bhi - blo);
    const int s = common_suffix(S.a + alo + p, ahi - alo - p, S.b + blo + p, bhi - blo - p);
    const int start = S.nd;
    if (p) push(S, EQ, p);
    const int ml = alo + p, mh = ahi - s, nl = blo + p, nh = bhi - s;
    const int n = mh - ml, m = nh - nl;
    bool deferred = false;
    if (n == 0 && m == 0) {
    } else if (n == 0) {
      push(S, INS, m);
    } else if (m == 0) {
      push(S, DEL, n);
    } else {
      const bool a_longer = n > m;
      const int at = a_longer ? find_in(S.a + ml, n, S.b + nl, m) : find_in(S.b + nl, m, S.a + ml, n);
      if (at >= 0) {
        const int8_t op = a_longer ? DEL : INS;
        const int lng = a_longer ? n : m, sht = a_longer ? m : n;
        push(S, op, at);  // may be empty, like the restated algorithm
        push(S, EQ, sht);
        push(S, op, lng - at - sht);
      } else if (imin(n, m) == 1) {
        push(S, DEL, n);
        push(S, INS, m);
      } else {
        int x = 0, y = 0;
        if (bisect_split(S, ml, mh, nl, nh, &x, &y)) {
          if (sp + 3 > MAXFRAME) {
            S.overflow = true;
            return;
          }
          st[sp++] = Frame{(uint8_t)alo, 0, (uint8_t)blo, 0, (uint8_t)start, (uint8_t)s, 1, 0};
          st[sp++] = Frame{(uint8_t)(ml + x), (uint8_t)mh, (uint8_t)(nl + y), (uint8_t)nh, 0, 0, 0, 0};  // right half
          st[sp++] = Frame{(uint8_t)ml, (uint8_t)(ml + x), (uint8_t)nl, (uint8_t)(nl + y), 0, 0, 0, 0};  // left half first
          deferred = true;
        } else {
          push(S, DEL, n);
          push(S, INS, m);
        }
      }
    }
    if (!deferred) {
      if (s) push(S, EQ, s);
      merge_pass(S, start, alo, blo);
    }
  }
}

// character classes of the boundary scores: ASCII bytes here (host + device); editscript.cpp supplies the
// Unicode version for scalar values
struct AsciiClass {
  static ESF_FN bool alnum(uint32_t c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
  static ESF_FN bool space(uint32_t c) { return c == ' ' || (c >= 9 && c <= 13); }
};

// Boundary quality between two strings (6 = edge ... 0 = inside a word), ASCII restriction of
// boundary_score in editscript.cpp.
template <class CC, class T>
ESF_FN int boundary_score(const T* one, int n1, const T* two, int n2) {
  if (n1 == 0 || n2 == 0) return 6;
  const uint32_t c1 = one[n1 - 1], c2 = two[0];
  const bool na1 = !CC::alnum(c1), na2 = !CC::alnum(c2);
  const bool ws1 = na1 && CC::space(c1), ws2 = na2 && CC::space(c2);
  const bool lb1 = ws1 && (c1 == '\n' || c1 == '\r'), lb2 = ws2 && (c2 == '\n' || c2 == '\r');
  const bool tail_blank = (n1 >= 2 && one[n1 - 1] == '\n' && one[n1 - 2] == '\n') ||
                          (n1 >= 3 && one[n1 - 1] == '\n' && one[n1 - 2] == '\r' && one[n1 - 3] == '\n');
  bool head_blank = false;
  if (n2 >= 2 && two[0] == '\n' && two[1] == '\n') head_blank = true;
  if (n2 >= 3 && two[0] == '\n' && two[1] == '\r' && two[2] == '\n') head_blank = true;
  if (n2 >= 3 && two[0] == '\r' && two[1] == '\n' && two[2] == '\n') head_blank = true;
  if (n2 >= 4 && two[0] == '\r' && two[1] == '\n' && two[2] == '\r' && two[3] == '\n') head_blank = true;
  if ((lb1 && tail_blank) || (lb2 && head_blank)) return 5;
  if (lb1 || lb2) return 4;
  if (na1 && !ws1 && ws2) return 3;
  if (ws1 || ws2) return 2;
  if (na1 || na2) return 1;
  return 0;
}

template <class CC, class T>
ESF_FN void lossless_shift(State<T>& S) {
  for (int k = 1; k + 1 < S.nd; ++k) {
    if (S.d[k - 1].op != EQ || S.d[k + 1].op != EQ) continue;
    int ka, kb;
    seg_pos(S, 0, k, 0, 0, &ka, &kb);
    const bool del = S.d[k].op == DEL;
    const T* base = del ? S.a : S.b;
    int pos = del ? ka : kb;
    const int m = S.d[k].len;
    int l1 = S.d[k - 1].len, l2 = S.d[k + 1].len;
    const int cs = common_suffix(base + pos - l1, l1, base + pos, m);
    pos -= cs;
    l1 -= cs;
    l2 += cs;
    int best_l1 = l1, best_l2 = l2;
    int best = boundary_score<CC>(base + pos - l1, l1, base + pos, m) + boundary_score<CC>(base + pos, m, base + pos + m, l2);
    while (m > 0 && l2 > 0 && base[pos] == base[pos + m]) {
      ++pos;
      ++l1;
      --l2;
      const int sc = boundary_score<CC>(base + pos - l1, l1, base + pos, m) + boundary_score<CC>(base + pos, m, base + pos + m, l2);
      if (sc >= best) {
        best = sc;
        best_l1 = l1;
        best_l2 = l2;
      }
    }
    if (S.d[k - 1].len != best_l1) {
      int kk = k;
      if (best_l1) {
        S.d[kk - 1].len = (uint8_t)best_l1;
      } else {
        erase(S, kk - 1, 1);
        --kk;
      }
      if (best_l2) {
        S.d[kk + 1].len = (uint8_t)best_l2;
      } else {
        erase(S, kk + 1, 1);
        --kk;
      }
      k = kk;
    }
  }
}

template <class CC, class T>
ESF_FN void semantic_pass(State<T>& S) {
  bool changed = false;
  uint8_t eqs[MAXSEG];
  int ne = 0;
  bool have = false;
  int lasteq = 0, i = 0;
  int ins1 = 0, del1 = 0, ins2 = 0, del2 = 0;
  while (i < S.nd && !S.overflow) {
    if (S.d[i].op == EQ) {
      if (ne >= MAXSEG) {
        S.overflow = true;
        return;
      }
      eqs[ne++] = (uint8_t)i;
      ins1 = ins2;
      del1 = del2;
      ins2 = del2 = 0;
      lasteq = S.d[i].len;
      have = true;
    } else {
      if (S.d[i].op == INS) ins2 += S.d[i].len; else del2 += S.d[i].len;
      if (have && lasteq <= imax(ins1, del1) && lasteq <= imax(ins2, del2)) {
        const int at = eqs[ne - 1];
        insert_at(S, at, DEL, lasteq);
        if (S.overflow) return;
        S.d[at + 1].op = INS;
        --ne;
        if (ne > 0) --ne;
        i = ne == 0 ? -1 : (int)eqs[ne - 1];
        ins1 = del1 = ins2 = del2 = 0;
        have = false;
        changed = true;
      }
    }
    ++i;
  }
  if (S.overflow) return;
  if (changed) merge_pass(S, 0, 0, 0);
  if (S.overflow) return;
  lossless_shift<CC>(S);
  for (int k = 1; k < S.nd && !S.overflow; ++k) {
    if (S.d[k - 1].op == DEL && S.d[k].op == INS) {
      int ka, kb;
      seg_pos(S, 0, k - 1, 0, 0, &ka, &kb);
      const int dl = S.d[k - 1].len, il = S.d[k].len;
      const T* del = S.a + ka;
      const T* ins = S.b + kb;
      const int o1 = overlap_len(del, dl, ins, il), o2 = overlap_len(ins, il, del, dl);
      if (o1 >= o2) {
        if (o1 * 2 >= dl || o1 * 2 >= il) {
          insert_at(S, k, EQ, o1);
          if (S.overflow) return;
          S.d[k - 1].len = (uint8_t)(dl - o1);
          S.d[k + 1].len = (uint8_t)(il - o1);
          ++k;
        }
      } else if (o2 * 2 >= dl || o2 * 2 >= il) {
        insert_at(S, k, EQ, o2);
        if (S.overflow) return;
        S.d[k - 1].op = INS;
        S.d[k - 1].len = (uint8_t)(il - o2);
        S.d[k + 1].op = DEL;
        S.d[k + 1].len = (uint8_t)(dl - o2);
        ++k;
      }
      ++k;
    }
  }
}

// The edit script of a -> b as views into the two strings.  Returns the number of instructions, or
// -1 when the pair is outside this implementation's limits (caller falls back to the host).
template <class CC, class T>
ESF_FN int shortest_edit_script_t(const T* a, int na, const T* b, int nb, View* out /* [MAXSEG] */) {
  if (na > MAXLEN || nb > MAXLEN) return -1;
  State<T> S;
  S.a = a;
  S.b = b;
  S.nd = 0;
  S.overflow = false;
  diff_main(S, na, nb);
  if (S.overflow) return -1;
  semantic_pass<CC>(S);
  if (S.overflow) return -1;
  merge_pass(S, 0, 0, 0);
  if (S.overflow) return -1;
  int nv = 0, pa = 0, pb = 0;
  for (int i = 0; i < S.nd; ++i) {
    const Seg s = S.d[i];
    if (s.len > 0) {
      out[nv].op = s.op;
      out[nv].len = s.len;
      out[nv].pos = (uint8_t)(s.op == INS ? pb : pa);
      out[nv].pad = 0;
      ++nv;
    }
    if (s.op != INS) pa += s.len;
    if (s.op != DEL) pb += s.len;
  }
  return nv;
}
ESF_FN int shortest_edit_script(const uint8_t* a, int na, const uint8_t* b, int nb, View* out /* [MAXSEG] */) {
  return shortest_edit_script_t<AsciiClass, uint8_t>(a, na, b, nb, out);
}

// ---- confusable patterns as flat tables (built by Engine::ensure_confusable_table) --------------------
struct PatTable {
  const ConfPat* pats;
  const ConfInstr* instrs;
  const ConfOpt* opts;
  const uint16_t* text;  // option texts back to back (UTF-16 code units)
  uint32_t n_pats;
};

template <class T>
ESF_FN bool same_text(const T* x, const uint16_t* t, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i)
    if ((uint32_t)x[i] != (uint32_t)t[i]) return false;
  return true;
}
template <class T>
ESF_FN bool view_sfx(const T* p, uint32_t len, const uint16_t* t, uint32_t n) { return len >= n && same_text(p + len - n, t, n); }
template <class T>
ESF_FN bool view_pfx(const T* p, uint32_t len, const uint16_t* t, uint32_t n) { return len >= n && same_text(p, t, n); }
template <class T>
ESF_FN bool view_eq(const T* p, uint32_t len, const uint16_t* t, uint32_t n) { return len == n && same_text(p, t, n); }

// Confusable::found_in (src/confusables.rs:47-128) over the flat tables; `ref` is the script of a -> b
template <class Ch>
ESF_FN bool found_in(const PatTable& T, const ConfPat& c, const Ch* a, const Ch* b, const View* ref, int nref) {
  const int l = c.n_instr;
  int matches = 0;
  for (int i = 0; i < nref; ++i) {
    if (matches >= l) continue;
    const ConfInstr ins = T.instrs[c.first_instr + matches];
    bool found = false;
    if (ins.op == ref[i].op) {
      for (uint32_t o = 0; o < ins.n_opts && !found; ++o) {
        const ConfOpt opt = T.opts[ins.first_opt + o];
        const uint16_t* t = T.text + opt.text_off;
        const Ch* rp = (ref[i].op == INS ? b : a) + ref[i].pos;
        const uint32_t rl = ref[i].len;
        bool ok;
        if (ins.op != 0)
          ok = view_sfx(rp, rl, t, opt.text_len);
        else if (matches == 0 && matches == l - 1)
          ok = view_eq(rp, rl, t, opt.text_len);
        else if (matches == 0)
          ok = view_sfx(rp, rl, t, opt.text_len);
        else if (matches == l - 1)
          ok = view_pfx(rp, rl, t, opt.text_len);
        else
          ok = view_eq(rp, rl, t, opt.text_len);
        found = ok;
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

}  // namespace esf
}  // namespace anl
