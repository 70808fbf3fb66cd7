"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs.  Nothing under analiticcl_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile the oracle if needed (g++; seconds)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


class Threshold(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ratio", C.c_float), ("value", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [
        ("max_anagram_distance", Threshold),
        ("max_edit_distance", Threshold),
        ("max_matches", C.c_uint64),
        ("score_threshold", C.c_double),
        ("cutoff_threshold", C.c_double),
        ("stop_at_exact_match", C.c_int32),
        ("freq_weight", C.c_float),
        ("max_ngram", C.c_int32),
        ("unicodeoffsets", C.c_int32),
        ("max_seq", C.c_int32),
        ("lm_weight", C.c_float),
        ("variantmodel_weight", C.c_float),
        ("contextrules_weight", C.c_float),
    ]


class Result(C.Structure):
    _fields_ = [("vocab_id", C.c_uint64), ("dist_score", C.c_double), ("freq_score", C.c_double), ("via", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("queries", "modulo_tests", "deletions", "anagram_hits", "dl_pairs", "dl_cells", "survivors")]


def _threshold(v):
    """int -> Absolute, float -> Ratio, (float, int) -> RatioWithLimit (src/types.rs:75-83)."""
    if isinstance(v, tuple):
        return Threshold(1, float(v[0]), int(v[1]))
    if isinstance(v, bool):
        raise ValueError("bad threshold")
    if isinstance(v, int):
        return Threshold(2, 0.0, v)
    return Threshold(0, float(v), 0)


def make_params(max_anagram_distance=3, max_edit_distance=3, max_matches=20, score_threshold=0.25,
                cutoff_threshold=2.0, stop_at_exact_match=False, freq_weight=0.0, max_ngram=3,
                unicodeoffsets=False, max_seq=250, lm_weight=1.0, variantmodel_weight=3.0, contextrules_weight=1.0):
    """Defaults = SearchParameters::default() (src/types.rs:170-192)."""
    return Params(_threshold(max_anagram_distance), _threshold(max_edit_distance), max_matches, score_threshold,
                  cutoff_threshold, int(stop_at_exact_match), freq_weight, max_ngram, int(unicodeoffsets),
                  max_seq, lm_weight, variantmodel_weight, contextrules_weight)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, cp, u64, i64, i32, u32 = C.c_void_p, C.c_char_p, C.c_uint64, C.c_int64, C.c_int32, C.c_uint32
    P = C.POINTER
    sig = {
        "orc_new": (vp, [cp, u64, P(C.c_double)]),
        "orc_free": (None, [vp]),
        "orc_alphabet_len": (u32, [vp]),
        "orc_read_vocabulary": (i32, [vp, cp, i32, i32, i32, i32]),
        "orc_add_to_vocabulary": (u64, [vp, cp, i32, u32, i32, i32, i32]),
        "orc_set_have_freq": (None, [vp, i32]),
        "orc_add_confusable": (i32, [vp, cp, C.c_double]),
        "orc_set_confusables_before_pruning": (None, [vp]),
        "orc_build": (None, [vp]),
        "orc_has": (i32, [vp, cp]),
        "orc_vocab_size": (u64, [vp]),
        "orc_index_size": (u64, [vp]),
        "orc_instance_count": (u64, [vp]),
        "orc_sortedindex_count": (u64, [vp, u32]),
        "orc_max_key_bits": (u32, [vp]),
        "orc_vocab_text": (cp, [vp, u64]),
        "orc_vocab_freq": (u32, [vp, u64]),
        "orc_vocab_lexindex": (u32, [vp, u64]),
        "orc_vocab_lookup": (i64, [vp, cp]),
        "orc_free_str": (None, [vp]),
        "orc_anahash": (vp, [vp, cp]),
        "orc_normalize": (i64, [vp, cp, P(C.c_uint8), i64]),
        "orc_ana_insert": (vp, [cp, cp]),
        "orc_ana_contains": (i32, [cp, cp]),
        "orc_ana_delete": (vp, [cp, cp]),
        "orc_alphabet_upper_bound": (None, [cp, u32, P(u32), P(u32)]),
        "orc_deletions": (vp, [cp, u32, i32, i32, i32, i32, i32]),
        "orc_damerau_levenshtein": (i32, [P(C.c_uint8), u64, P(C.c_uint8), u64, u32]),
        "orc_lcs": (u32, [P(C.c_uint8), u64, P(C.c_uint8), u64]),
        "orc_prefix": (u32, [P(C.c_uint8), u64, P(C.c_uint8), u64]),
        "orc_suffix": (u32, [P(C.c_uint8), u64, P(C.c_uint8), u64]),
        "orc_threshold": (u32, [i32, C.c_float, u32, u64]),
        "orc_edit_script": (vp, [cp, cp]),
        "orc_confusable_found_in": (i32, [cp, cp, cp]),
        "orc_nearest": (vp, [vp, cp, u32, i32]),
        "orc_add_variant": (i32, [vp, u64, cp, C.c_double, i32, u32, i32, i32, i32]),
        "orc_read_variants": (i32, [vp, cp, i32, i32, i32]),
        "orc_vocab_type": (u32, [vp, u64]),
        "orc_find_variants": (i64, [vp, cp, P(Params), P(Result), i64]),
        "orc_find_variants_batch": (i32, [vp, cp, P(u64), u64, P(Params), i32, P(u64), P(P(Result)), P(Stats)]),
        "orc_free_results": (None, [P(Result)]),
        "orc_max_threads": (i32, []),
        "orc_find_all_segments": (i64, [vp, cp, u64, P(Params), P(u64), P(u64), P(u32), P(C.c_uint8), P(u64), i64,
                                        P(Result), i64]),
        "orc_find_all_matches": (i64, [vp, cp, u64, P(Params), P(C.c_uint8), P(u64), P(Result), P(u64), P(u64), P(u32),
                                       P(i32), P(u64), i64, P(Result), i64, P(u64), P(C.c_uint16), P(C.c_uint8), i64]),
        "orc_learn_variants": (u64, [vp, cp, P(u64), u64, P(Params), i32, i32]),
        "orc_learn_apply": (u64, [vp, cp, P(u64), u64, P(u64), P(C.c_double)]),
        "orc_vocab_links": (i64, [vp, u64, i32, P(u64), P(C.c_double), i64]),
        "orc_have_lm": (i32, [vp]),
        "orc_ngram_count": (u64, [vp]),
        "orc_add_contextrule": (i32, [vp, cp, C.c_float, cp, cp]),
        "orc_read_contextrules": (i32, [vp, cp]),
        "orc_rule_error": (cp, []),
        "orc_contextrule_count": (u64, [vp]),
        "orc_tag_count": (u64, [vp]),
        "orc_tag_name": (cp, [vp, u64]),
        "orc_set_bruteforce_limit": (None, [u64]),
        "orc_lm_score_tokens": (None, [vp, P(i64), u64, P(C.c_float), P(C.c_double)]),
        "orc_find_boundaries": (i64, [cp, u64, P(u64), P(u64), P(i32), i64]),
        "orc_find_match_ngrams": (i64, [cp, u64, u32, P(u64), P(u64), i64]),
        "orc_is_alphabetic": (i32, [u32]),
        "orc_is_lowercase": (i32, [u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _take_str(ptr):
    if not ptr:
        return None
    s = C.string_at(ptr).decode("utf-8")
    lib().orc_free_str(ptr)
    return s


def _u8(seq):
    arr = (C.c_uint8 * max(1, len(seq)))(*seq)
    return arr, len(seq)


TEST_ALPHABET_TSV = "\n".join(
    [f"{c}\t{c.upper()}" for c in "abcdefghijklmnopqrstuvwxyz"] + [".\t,"]) + "\n"
"""The 27-class alphabet of the reference's tests (src/test.rs:3-31)."""

VT = {"NONE": 0, "INDEXED": 1, "LM": 2, "TRANSPARENT": 5}
NO_VIA = (1 << 64) - 1
FH = {"sum": 0, "max": 1, "min": 2, "replace": 3}


class OracleModel:
    """Mirror of the reference VariantModel restricted to the variant-lookup path."""

    def __init__(self, alphabet_tsv=None, alphabet_file=None, weights=(0.5, 0.125, 0.125, 0.125, 0.125)):
        if alphabet_tsv is None:
            with open(alphabet_file, "rb") as f:
                raw = f.read()
        else:
            raw = alphabet_tsv.encode("utf-8") if isinstance(alphabet_tsv, str) else alphabet_tsv
        w = (C.c_double * 5)(*weights)
        self.h = lib().orc_new(raw, len(raw), w)
        self.nlex = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_free(self.h)
            self.h = None

    # -- model construction ---------------------------------------------------------------------
    def read_vocabulary(self, filename, text_column=0, freq_column=1, freq_handling="max", vocab_type="INDEXED"):
        rc = lib().orc_read_vocabulary(self.h, filename.encode(), text_column, -1 if freq_column is None else freq_column,
                                       FH[freq_handling], VT[vocab_type])
        if rc != 0:
            raise RuntimeError(f"oracle read_vocabulary({filename}) failed: {rc}")
        self.nlex += 1

    read_lexicon = read_vocabulary

    def add_to_vocabulary(self, text, frequency=None, freq_handling="max", vocab_type="INDEXED", index=0):
        return lib().orc_add_to_vocabulary(self.h, text.encode(), frequency is not None, frequency or 0,
                                           FH[freq_handling], VT[vocab_type], index)

    def add_variant(self, ref_id, variant, score, frequency=None, freq_handling="max", vocab_type="INDEXED", index=0):
        rc = lib().orc_add_variant(self.h, ref_id, variant.encode(), score, frequency is not None, frequency or 0,
                                   FH[freq_handling], VT[vocab_type] if isinstance(vocab_type, str) else vocab_type, index)
        if rc < 0:
            raise ValueError("add_variant: unknown reference id")
        return bool(rc)

    def read_variants(self, filename, transparent=False, freq_handling="max", vocab_type="INDEXED"):
        rc = lib().orc_read_variants(self.h, filename.encode(), FH[freq_handling], VT[vocab_type], int(transparent))
        if rc != 0:
            raise RuntimeError(f"oracle read_variants({filename}) failed: {rc}")
        self.nlex += 1

    def learn_variants(self, inputs, params, strict=True, auto_build=True):
        """src/lib.rs:1062-1139 -> number of variants added."""
        raw = [t.encode("utf-8") for t in inputs]
        offs = (C.c_uint64 * (len(raw) + 1))()
        for i, r in enumerate(raw):
            offs[i + 1] = offs[i] + len(r)
        return lib().orc_learn_variants(self.h, b"".join(raw), offs, len(raw), C.byref(params), int(strict), int(auto_build))

    def learn_apply(self, items):
        """The bookkeeping half of learn_variants on explicit (input text, result vocab id, dist_score) triples."""
        raw = [t.encode("utf-8") for t, _, _ in items]
        offs = (C.c_uint64 * (len(raw) + 1))()
        for i, r in enumerate(raw):
            offs[i + 1] = offs[i] + len(r)
        ids = (C.c_uint64 * max(1, len(items)))(*[int(v) for _, v, _ in items])
        sc = (C.c_double * max(1, len(items)))(*[float(d) for _, _, d in items])
        return lib().orc_learn_apply(self.h, b"".join(raw), offs, len(raw), ids, sc)

    def vocab_links(self, vid):
        """-> ([(target id, score)...] VariantOf, [variant id...] ReferenceFor)"""
        ids, sc = (C.c_uint64 * 256)(), (C.c_double * 256)()
        n = lib().orc_vocab_links(self.h, vid, 0, ids, sc, 256)
        of = [(ids[i], sc[i]) for i in range(n)]
        n = lib().orc_vocab_links(self.h, vid, 1, ids, sc, 256)
        return of, [ids[i] for i in range(n)]

    def read_lm(self, filename):
        """bindings/python/src/lib.rs:659-667: read_vocabulary with VocabType::LM."""
        self.read_vocabulary(filename, vocab_type="LM")

    def add_contextrule(self, pattern, score, tag=(), tagoffset=()):
        rc = lib().orc_add_contextrule(self.h, pattern.encode(), score, "\n".join(tag).encode(), "\n".join(tagoffset).encode())
        if rc != 0:
            raise RuntimeError("Error parsing context rule: " + lib().orc_rule_error().decode())

    def read_contextrules(self, filename):
        rc = lib().orc_read_contextrules(self.h, filename.encode())
        if rc != 0:
            raise RuntimeError(f"oracle read_contextrules({filename}) failed: {rc} {lib().orc_rule_error().decode()}")

    def have_lm(self):
        return bool(lib().orc_have_lm(self.h))

    def ngram_count(self):
        return lib().orc_ngram_count(self.h)

    def tags(self):
        return [lib().orc_tag_name(self.h, i).decode() for i in range(lib().orc_tag_count(self.h))]

    def lm_score_tokens(self, tokens):
        """tokens: vocabulary ids, None = out of vocabulary -> (logprob f32, perplexity f64)."""
        arr = (C.c_int64 * len(tokens))(*[-1 if t is None else t for t in tokens])
        lp, pp = C.c_float(), C.c_double()
        lib().orc_lm_score_tokens(self.h, arr, len(tokens), C.byref(lp), C.byref(pp))
        return lp.value, pp.value

    def add_to_confusables(self, script, weight):
        if lib().orc_add_confusable(self.h, script.encode(), weight) != 0:
            raise ValueError("bad confusable pattern " + script)

    def set_confusables_before_pruning(self):
        lib().orc_set_confusables_before_pruning(self.h)

    def build(self):
        lib().orc_build(self.h)

    # -- introspection ------------------------------------------------------------------------------
    def has(self, text):
        return bool(lib().orc_has(self.h, text.encode()))

    def vocab_text(self, vid):
        return lib().orc_vocab_text(self.h, vid).decode("utf-8")

    def vocab_lexindex(self, vid):
        return lib().orc_vocab_lexindex(self.h, vid)

    def vocab_size(self):
        return lib().orc_vocab_size(self.h)

    def vocab_freq(self, vid):
        return lib().orc_vocab_freq(self.h, vid)

    def vocab_type(self, vid):
        return lib().orc_vocab_type(self.h, vid)

    def vocab_lookup(self, text):
        return lib().orc_vocab_lookup(self.h, text.encode())

    def index_size(self):
        return lib().orc_index_size(self.h)

    def instance_count(self):
        return lib().orc_instance_count(self.h)

    def sortedindex_count(self, cc):
        return lib().orc_sortedindex_count(self.h, cc)

    def max_key_bits(self):
        return lib().orc_max_key_bits(self.h)

    def anahash(self, text):
        return int(_take_str(lib().orc_anahash(self.h, text.encode())))

    def normalize(self, text):
        buf = (C.c_uint8 * 1024)()
        n = lib().orc_normalize(self.h, text.encode(), buf, 1024)
        return list(buf[:n])

    def nearest(self, text, k, stop_at_exact=False):
        s = _take_str(lib().orc_nearest(self.h, text.encode(), k, int(stop_at_exact)))
        return [int(x) for x in s.split("\n") if x]

    # -- queries -----------------------------------------------------------------------------------
    def find_variants(self, text, params, with_via=False):
        cap = 4096
        while True:
            buf = (Result * cap)()
            n = lib().orc_find_variants(self.h, text.encode(), C.byref(params), buf, cap)
            if n <= cap:
                if with_via:
                    return [(buf[i].vocab_id, buf[i].dist_score, buf[i].freq_score, None if buf[i].via == NO_VIA else buf[i].via)
                            for i in range(n)]
                return [(buf[i].vocab_id, buf[i].dist_score, buf[i].freq_score) for i in range(n)]
            cap = n

    def find_variants_batch(self, queries, params, threads=0, want_stats=False, with_via=False):
        """queries: list[str] -> list[list[(vocab_id, dist, freq)]] (+ Stats)."""
        enc = [q.encode("utf-8") for q in queries]
        blob = b"".join(enc)
        n = len(enc)
        offs = (C.c_uint64 * (n + 1))()
        pos = 0
        for i, e in enumerate(enc):
            offs[i] = pos
            pos += len(e)
        offs[n] = pos
        out_offs = (C.c_uint64 * (n + 1))()
        res = C.POINTER(Result)()
        st = Stats()
        lib().orc_find_variants_batch(self.h, blob, offs, n, C.byref(params), threads, out_offs, C.byref(res),
                                      C.byref(st) if want_stats else None)
        out = []
        for i in range(n):
            if with_via:
                out.append([(res[j].vocab_id, res[j].dist_score, res[j].freq_score, None if res[j].via == NO_VIA else res[j].via)
                            for j in range(out_offs[i], out_offs[i + 1])])
            else:
                out.append([(res[j].vocab_id, res[j].dist_score, res[j].freq_score)
                            for j in range(out_offs[i], out_offs[i + 1])])
        lib().orc_free_results(res)
        return (out, st) if want_stats else out

    def find_variants_batch_raw(self, blob, offs, n, params, threads=0):
        """Timed entry for bench.py: no Python-side result decoding.  Returns (total results, Stats)."""
        out_offs = (C.c_uint64 * (n + 1))()
        res = C.POINTER(Result)()
        st = Stats()
        lib().orc_find_variants_batch(self.h, blob, offs, n, C.byref(params), threads, out_offs, C.byref(res), C.byref(st))
        total = out_offs[n]
        lib().orc_free_results(res)
        return total, st

    def find_all_segments(self, text, params):
        raw = text.encode("utf-8")
        seg_cap, res_cap = 4096, 1 << 16
        while True:
            sb = (C.c_uint64 * seg_cap)()
            se = (C.c_uint64 * seg_cap)()
            sn = (C.c_uint32 * seg_cap)()
            sl = (C.c_uint8 * seg_cap)()
            ro = (C.c_uint64 * (seg_cap + 1))()
            rs = (Result * res_cap)()
            n = lib().orc_find_all_segments(self.h, raw, len(raw), C.byref(params), sb, se, sn, sl, ro, seg_cap, rs, res_cap)
            if n < seg_cap and ro[n] <= res_cap:
                break
            seg_cap = max(seg_cap, n + 1)
            res_cap *= 4
        out = []
        for i in range(n):
            out.append({
                "begin": sb[i], "end": se[i], "n": sn[i], "looked_up": bool(sl[i]),
                "text": raw[sb[i]:se[i]].decode("utf-8"),
                "variants": [(rs[j].vocab_id, rs[j].dist_score, rs[j].freq_score) for j in range(ro[i], ro[i + 1])],
            })
        return out


    def find_all_matches(self, text, params, segments=None):
        """find_all_matches with the sequence consolidation (src/lib.rs:1790-1957, 2088-2495) incl. the language model
        and context rules of this model.  `segments` (optional): the variant lists per producer segment instead of
        lookups (see consolidate)."""
        return _find_all_matches(self.h, text, params, segments)


def consolidate(text, params, segments):
    """The consolidation alone: `segments` = one dict per producer segment (find_all_segments order) with
    "looked_up" and "variants" [(vocab_id, dist_score, freq_score)]; no model (hence no LM, no context rules), no lookups."""
    return _find_all_matches(None, text, params, segments)


def _find_all_matches(h, text, params, segments):
    raw = text.encode("utf-8")
    looked = offs = prov = None
    if segments is not None:
        n = len(segments)
        looked = (C.c_uint8 * max(1, n))(*[1 if s["looked_up"] else 0 for s in segments])
        offs = (C.c_uint64 * (n + 1))()
        flat = []
        for i, s in enumerate(segments):
            flat += list(s["variants"]) if s["looked_up"] else []
            offs[i + 1] = len(flat)
        prov = (Result * max(1, len(flat)))()
        for j, v in enumerate(flat):
            prov[j] = Result(int(v[0]), float(v[1]), float(v[2]), NO_VIA)
    seg_cap, res_cap = 4096, 1 << 16
    while True:
        sb, se = (C.c_uint64 * seg_cap)(), (C.c_uint64 * seg_cap)()
        sn, ss = (C.c_uint32 * seg_cap)(), (C.c_int32 * seg_cap)()
        ro = (C.c_uint64 * (seg_cap + 1))()
        rs = (Result * res_cap)()
        to = (C.c_uint64 * (seg_cap + 1))()
        tg, sq = (C.c_uint16 * res_cap)(), (C.c_uint8 * res_cap)()
        n = lib().orc_find_all_matches(h, raw, len(raw), C.byref(params), looked, offs, prov, sb, se, sn, ss, ro, seg_cap,
                                       rs, res_cap, to, tg, sq, res_cap)
        if n < seg_cap and ro[n] <= res_cap and to[n] <= res_cap:
            break
        seg_cap = max(seg_cap, n + 1)
        res_cap *= 4
    return [{"begin": sb[i], "end": se[i], "n": sn[i], "selected": ss[i], "text": raw[sb[i]:se[i]].decode("utf-8"),
             "variants": [(rs[j].vocab_id, rs[j].dist_score, rs[j].freq_score) for j in range(ro[i], ro[i + 1])],
             "tag": [tg[j] for j in range(to[i], to[i + 1])], "seqnr": [sq[j] for j in range(to[i], to[i + 1])]}
            for i in range(n)]


# -- free functions (primitives) --------------------------------------------------------------------
def ana_insert(a, b):
    return int(_take_str(lib().orc_ana_insert(str(a).encode(), str(b).encode())))


def ana_contains(a, b):
    return bool(lib().orc_ana_contains(str(a).encode(), str(b).encode()))


def ana_delete(a, b):
    s = _take_str(lib().orc_ana_delete(str(a).encode(), str(b).encode()))
    return None if s is None else int(s)


def alphabet_upper_bound(v, alphabet_size):
    a, b = C.c_uint32(), C.c_uint32()
    lib().orc_alphabet_upper_bound(str(v).encode(), alphabet_size, C.byref(a), C.byref(b))
    return a.value, b.value


def deletions(v, alphabet_size, mode, maxdepth=-1, breadthfirst=False, allow_duplicates=True, allow_empty_leaves=True):
    """-> list of (value, depth, charindex).  mode: 'parents' | 'singlebeam' | 'recursive'."""
    m = {"parents": 0, "singlebeam": 1, "recursive": 2}[mode]
    s = _take_str(lib().orc_deletions(str(v).encode(), alphabet_size, m, maxdepth, int(breadthfirst),
                                      int(allow_duplicates), int(allow_empty_leaves)))
    return [tuple(int(x) for x in line.split(":")) for line in s.split("\n") if line]


def damerau_levenshtein(s, t, maxd):
    a, na = _u8(s)
    b, nb = _u8(t)
    r = lib().orc_damerau_levenshtein(a, na, b, nb, maxd)
    return None if r < 0 else r


def lcs(s, t):
    a, na = _u8(s)
    b, nb = _u8(t)
    return lib().orc_lcs(a, na, b, nb)


def prefix(s, t):
    a, na = _u8(s)
    b, nb = _u8(t)
    return lib().orc_prefix(a, na, b, nb)


def suffix(s, t):
    a, na = _u8(s)
    b, nb = _u8(t)
    return lib().orc_suffix(a, na, b, nb)


def threshold(v, length):
    t = _threshold(v)
    return lib().orc_threshold(t.kind, t.ratio, t.value, length)


def edit_script(src, dst):
    return _take_str(lib().orc_edit_script(src.encode(), dst.encode()))


def confusable_found_in(pattern, src, dst):
    return bool(lib().orc_confusable_found_in(pattern.encode(), src.encode(), dst.encode()))


def find_boundaries(text):
    raw = text.encode("utf-8")
    cap = len(raw) + 2
    b = (C.c_uint64 * cap)()
    e = (C.c_uint64 * cap)()
    s = (C.c_int32 * cap)()
    n = lib().orc_find_boundaries(raw, len(raw), b, e, s, cap)
    return [(b[i], e[i], raw[b[i]:e[i]].decode("utf-8"), s[i]) for i in range(n)]


def find_match_ngrams(text, order):
    raw = text.encode("utf-8")
    cap = len(raw) + 2
    b = (C.c_uint64 * cap)()
    e = (C.c_uint64 * cap)()
    n = lib().orc_find_match_ngrams(raw, len(raw), order, b, e, cap)
    return [raw[b[i]:e[i]].decode("utf-8") for i in range(n)]
