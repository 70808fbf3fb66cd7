"""CPU-only: sequence consolidation of find_all_matches (most_likely_sequence, src/lib.rs:2088-2495, without LM and
context rules).  The oracle's restatement is checked against the expectations of the reference's tests 0702/0703
(tests/main.rs:1143-1253; their LM entries only feed the language model, which is out of scope, and are left out),
then the product's host code (csrc/search.cpp, anl_match_set_consolidate) is compared with the oracle on the same
lattices: real variant lists (looked up by the oracle on the CPU) and random ones with many cost ties."""
import ctypes as C

import numpy as np
import pytest

import workloads
from oracle import orc

TEST_PARAMS = dict(max_anagram_distance=2, max_edit_distance=2, max_matches=10, score_threshold=0.0,
                   cutoff_threshold=0.0, freq_weight=0.0, max_ngram=2)  # get_test_searchparams(), src/test.rs:48-68


@pytest.fixture(scope="module")
def L():
    from analiticcl_b200 import build, _capi
    build.build()
    return _capi.lib()


LM_ENTRIES = [("<bos> I", 2), ("I think", 2), ("I sink", 1), ("you are", 2), ("right <eos>", 2)]  # T:1154-1193


def small_model(words, lm=()):
    m = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in words:
        m.add_to_vocabulary(w, 2)
    for w, f in lm:
        m.add_to_vocabulary(w, f, vocab_type="LM")
    m.build()
    return m


def rendered(m, matches):
    return [(s["text"], m.vocab_text(s["variants"][s["selected"]][0]) if s["selected"] >= 0 else None) for s in matches]


# ---- the oracle against the reference's expectations -------------------------------------------------------
def test_oracle_reference_0702_0703():
    m = small_model(["I", "think", "sink", "you", "are", "right", "are right"])
    p = orc.make_params(**TEST_PARAMS)
    r = m.find_all_matches("I tink you are rihgt", p)
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right")]  # T:1196-1208
    assert (r[1]["begin"], r[1]["end"]) == (2, 6)
    r = m.find_all_matches("I tink you are\nrihgt", p)
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are\nrihgt", "are right")]  # T:1243-1252


def test_oracle_reference_0705_lm_disabled():
    """tests/main.rs:1364-1429, the whole test: the model holds the LM entries but `lm_weight = 0.0` switches the
    language model off (src/lib.rs:2336, 2392-2396), so the sequence is decided by the variant-model cost alone --
    exactly the configuration restated here.  (LM-typed entries are not indexed and never show up as variants.)"""
    m = small_model(["I", "think", "sink", "you", "are", "right", "are right"], LM_ENTRIES)
    r = m.find_all_matches("I tink you are rihgt", orc.make_params(**TEST_PARAMS, lm_weight=0.0))
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right")]  # T:1420-1428


def test_oracle_reference_0704_two_batches():
    """tests/main.rs:1274-1361 without the language model's say: a double line break is a hard boundary, the two
    sentences are consolidated independently."""
    m = small_model(["I", "think", "sink", "you", "are", "right", "am", "sure", "are right"])
    r = m.find_all_matches("I tink you are rihgt\n\nI am sur", orc.make_params(**TEST_PARAMS))
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right"),
                              ("I", "I"), ("am", "am"), ("sur", "sure")]  # T:1346-1360


def test_oracle_tutorial_golden(eng_oracle):
    """tutorial.ipynb:476-481: find_all_matches("We would like sep arate beds") with default parameters
    (max_ngram = 3) -- the consolidated sequence holds the bigram "sep arate" at index 3, code points 14:23, and
    its ranked variant list (tests/golden/tutorial.json, generated from the notebook's recorded output)."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial.json"), encoding="utf-8"))
    g = gold["find_all_matches"]["We would like sep arate beds"]
    r = eng_oracle.find_all_matches("We would like sep arate beds", orc.make_params())
    assert [s["text"] for s in r] == ["We", "would", "like", "sep arate", "beds"]
    m = r[g["index"]]
    assert (m["text"], m["begin"], m["end"], m["selected"]) == ("sep arate", 14, 23, 0)
    assert [(eng_oracle.vocab_text(v), d) for v, d, _ in m["variants"]] == [(v["text"], v["dist_score"]) for v in g["match"]["variants"]]


def test_oracle_two_batches_and_oov():
    """Hard boundaries split the text into independent lattices (T:1255-1340 shape); a token without variants is
    copied from the input (selected = None, src/lib.rs:2210-2247) and never blocks the path."""
    m = small_model(["I", "think", "sink", "you", "are", "right", "am", "sure", "are right"])
    p = orc.make_params(**TEST_PARAMS)
    r = m.find_all_matches("I tink you are rihgt. I am sure", p)
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right"),
                              ("I", "I"), ("am", "am"), ("sure", "sure")]
    r = m.find_all_matches("I tink zzzzzzzzzz are rihgt", p)
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("zzzzzzzzzz", None), ("are rihgt", "are right")]
    # unigram-only search: no lattice, every token with its best variant selected (src/lib.rs:1929-1932)
    r = m.find_all_matches("I tink you are rihgt", orc.make_params(**{**TEST_PARAMS, "max_ngram": 1}))
    assert [(s["text"], s["selected"]) for s in r] == [(t, 0) for t in ["I", "tink", "you", "are", "rihgt"]]


# ---- the product's host code against the oracle ------------------------------------------------------------
def product_consolidate(L, text, sp, segments, unicodeoffsets=False):
    from analiticcl_b200 import _capi
    raw = text.encode("utf-8")
    n = len(segments)
    looked = (C.c_uint8 * max(1, n))(*[1 if s["looked_up"] else 0 for s in segments])
    offs = (C.c_uint64 * (n + 1))()
    flat = []
    for i, s in enumerate(segments):
        flat += list(s["variants"]) if s["looked_up"] else []
        offs[i + 1] = len(flat)
    vs = (_capi.Variant * max(1, len(flat)))()
    for j, (vid, d, f) in enumerate(flat):
        vs[j].vocab_id, vs[j].dist_score, vs[j].freq_score, vs[j].via = int(vid), float(d), float(f), (1 << 64) - 1
    ms, out = C.c_void_p(), C.c_void_p()
    assert L.anl_debug_match_set_build(raw, len(raw), sp.data.max_ngram, int(unicodeoffsets), looked, offs, vs, n,
                                       C.byref(ms)) == 0, L.anl_last_error()
    try:
        assert L.anl_match_set_len(ms) == n
        assert L.anl_match_set_consolidate(ms, raw, len(raw), C.byref(sp.data), C.byref(out)) == 0, L.anl_last_error()
    finally:
        L.anl_match_set_free(ms)  # the consolidated set must not point into the input set
    got = []
    m = _capi.Match()
    for i in range(L.anl_match_set_len(out)):
        assert L.anl_match_set_get(out, i, C.byref(m)) == 0
        got.append({"begin": int(m.begin), "end": int(m.end), "n": int(m.n), "selected": int(m.selected),
                    "variants": [(m.variants[j].vocab_id, m.variants[j].dist_score, m.variants[j].freq_score)
                                 for j in range(m.n_variants)] if m.variants else []})
    L.anl_match_set_free(out)
    return got


def strip(matches):
    return [{k: s[k] for k in ("begin", "end", "n", "selected", "variants")} for s in matches]


def to_orc(sp):
    d = sp.data
    return orc.make_params(int(d.max_anagram_distance.value), int(d.max_edit_distance.value), d.max_matches, d.score_threshold,
                           d.cutoff_threshold, False, d.freq_weight, d.max_ngram, False)


@pytest.mark.parametrize("max_ngram,freq_weight", [(1, 0.0), (2, 0.0), (3, 0.0), (3, 0.25)])
def test_product_matches_oracle_on_real_lattices(L, eng_oracle, max_ngram, freq_weight):
    import analiticcl_b200 as A
    text = workloads.cfg3_text(700, 3001) + " It's a well-known co-operative re_entry; über naïve façade!?  Done" + \
        " qqqqqqqqqqqq xxxxxxxxxxxx. We would like sep arate beds to gether with out dis agree ment. A\nb"
    sp = A.SearchParameters(max_ngram=max_ngram, max_anagram_distance=2, max_edit_distance=2, freq_weight=freq_weight)
    op = to_orc(sp)
    segments = eng_oracle.find_all_segments(text, op)
    exp = eng_oracle.find_all_matches(text, op)
    assert strip(orc.consolidate(text, op, segments)) == strip(exp)  # the oracle's hook path = its model path
    got = product_consolidate(L, text, sp, segments)
    if max_ngram == 1:
        # the reference sets selected = Some(0) on every unigram, also on one with an empty variant list
        # (src/lib.rs:1930); the C ABI reports "nothing selected" (-1) for those (include/analiticcl_b200.h)
        exp = [dict(s, selected=0 if s["variants"] else -1) for s in exp]
    assert got == strip(exp)
    if max_ngram > 1:
        assert any(s["n"] > 1 for s in exp) and any(s["selected"] < 0 for s in exp) and len(exp) < len(segments)
        # a path: consecutive matches of a batch touch boundaries, none overlap
        assert all(a["end"] <= b["begin"] for a, b in zip(exp, exp[1:]))
    else:
        assert len(exp) == len(segments)


def test_product_matches_oracle_on_random_lattices(L):
    """Random variant lists over every segment (also the ones the producer would have skipped): scores on a coarse
    grid so that equal-cost paths are common and the documented tie rule is exercised; code-point offsets."""
    import analiticcl_b200 as A
    from analiticcl_b200 import _capi
    rng = np.random.default_rng(11)
    words = ["aa", "b", "ccc", "dé", "e-f", "g'h", "ij", "k_l", "mmmm", "ñ"]
    seps = [" ", " ", " ", ", ", ". ", "\n", "-", "  ", "; "]
    for case in range(61):
        # (the last case is long: > 128 batches, so the consolidation runs on several host threads)
        ntok = int(rng.integers(1, 40)) if case < 60 else 6000
        toks = [words[int(i)] for i in rng.integers(0, len(words), size=ntok)]
        text = "".join(t + seps[int(rng.integers(0, len(seps)))] for t in toks)
        if case % 3 == 0:
            text = " " + text.rstrip()
        max_ngram = int(rng.integers(2, 5))
        fw = float(rng.choice([0.0, 0.5]))
        sp = A.SearchParameters(max_ngram=max_ngram, freq_weight=fw)
        op = to_orc(sp)
        raw = text.encode("utf-8")
        cap = max_ngram * (len(raw) + 2)
        b, e = (C.c_uint64 * cap)(), (C.c_uint64 * cap)()
        o, bt = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
        nseg = L.anl_debug_segment_text(raw, len(raw), max_ngram, b, e, o, bt, cap)
        segments = []
        for k in range(nseg):
            looked = o[k] == 1 or rng.random() < 0.7
            nv = int(rng.integers(0, 4)) if looked else 0
            vs = sorted(((int(rng.integers(3, 1000)), float(rng.integers(0, 5)) / 4.0, float(rng.integers(0, 3)) / 2.0)
                         for _ in range(nv)), key=lambda v: -v[1])
            segments.append({"looked_up": bool(looked), "variants": vs})
        exp = orc.consolidate(text, op, segments)
        got = product_consolidate(L, text, sp, segments)
        assert got == strip(exp), (case, text[:200])
        if case == 60:
            assert bt[nseg - 1] >= 256
        # code-point offsets: same matches, offsets remapped (src/lib.rs:1949-1956)
        spu = A.SearchParameters(max_ngram=max_ngram, freq_weight=fw, unicodeoffsets=True)
        gotu = product_consolidate(L, text, spu, segments, unicodeoffsets=True)
        assert [(len(raw[:s["begin"]].decode()), len(raw[:s["end"]].decode())) for s in got] == \
            [(s["begin"], s["end"]) for s in gotu]
        assert [(s["n"], s["selected"], s["variants"]) for s in got] == [(s["n"], s["selected"], s["variants"]) for s in gotu]


def test_consolidate_rejects_foreign_match_set(L):
    import analiticcl_b200 as A
    sp = A.SearchParameters(max_ngram=2)
    ms, out = C.c_void_p(), C.c_void_p()
    raw = b"one two three"
    looked, offs = (C.c_uint8 * 5)(1, 1, 1, 1, 1), (C.c_uint64 * 6)()
    assert L.anl_debug_match_set_build(raw, len(raw), 2, 0, looked, offs, None, 5, C.byref(ms)) == 0, L.anl_last_error()
    assert L.anl_match_set_consolidate(ms, b"one two", 7, C.byref(sp.data), C.byref(out)) != 0
    assert b"segment count" in L.anl_last_error()
    assert L.anl_match_set_consolidate(ms, raw, len(raw), C.byref(sp.data), C.byref(out)) == 0
    assert L.anl_match_set_len(out) == 3  # three out-of-vocabulary unigrams (cost 2 each) beat nothing else
    L.anl_match_set_free(out)
    L.anl_match_set_free(ms)


def test_python_mirror_consolidates_like_the_reference(L, monkeypatch):
    """The Python mirror's find_all_matches (bindings/python/src/lib.rs:752-805) with the GPU lookups replaced by the
    oracle's (CPU) variant lists: max_ngram > 1 returns the most likely sequence, consolidate_matches=False the
    producer's view; the selected variant comes first."""
    import analiticcl_b200 as A
    words = ["I", "think", "sink", "you", "are", "right", "are right"]
    o = small_model(words, LM_ENTRIES)
    m = A.VariantModel(None, A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in words:
        m.add_to_vocabulary(w, 2, A.VocabParams())
    for w, f in LM_ENTRIES:  # tests/main.rs:1364-1429 (lm_weight = 0: the entries are carried, the LM is off)
        m.add_to_vocabulary(w, f, A.VocabParams(vocabtype="LM"))
    try:
        m.build()  # host side of build() (index arrays, language model); the upload needs a GPU
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)

    def fake_find_all_matches(h, raw, n, params_ref, out_ref):
        sp = params_ref._obj
        segments = o.find_all_segments(raw[:n].decode("utf-8"), orc.make_params(**{**TEST_PARAMS, "max_ngram": sp.max_ngram}))
        k = len(segments)
        looked = (C.c_uint8 * max(1, k))(*[1 if s["looked_up"] else 0 for s in segments])
        offs = (C.c_uint64 * (k + 1))()
        flat = []
        for i, s in enumerate(segments):
            flat += s["variants"] if s["looked_up"] else []
            offs[i + 1] = len(flat)
        from analiticcl_b200 import _capi
        vs = (_capi.Variant * max(1, len(flat)))()
        for j, (vid, d, f) in enumerate(flat):
            vs[j].vocab_id, vs[j].dist_score, vs[j].freq_score, vs[j].via = int(vid), float(d), float(f), (1 << 64) - 1
        return L.anl_debug_match_set_build(raw, n, sp.max_ngram, sp.unicodeoffsets, looked, offs, vs, k, out_ref)

    real = L.anl_find_all_matches
    monkeypatch.setattr(L, "anl_find_all_matches", fake_find_all_matches, raising=False)
    try:
        sp = A.SearchParameters(**TEST_PARAMS, lm_weight=0.0, context_weight=0.5)
        r = m.find_all_matches("I tink you are rihgt", sp)
        assert [(x["input"], x["variants"][0]["text"]) for x in r] == \
            [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right")]
        assert r[1]["offset"] == {"begin": 2, "end": 6} and [v["text"] for v in r[1]["variants"]] == ["think", "sink"]
        every = m.find_all_matches("I tink you are rihgt", A.SearchParameters(**TEST_PARAMS, consolidate_matches=False))
        assert [x["input"] for x in every][:5] == ["I", "tink", "you", "are", "rihgt"]
        assert {"are rihgt"} <= {x["input"] for x in every[5:]} and len(every) > len(r)
        r = m.find_all_matches("I tink zzzzzzzzzz are rihgt", sp)  # out-of-vocabulary token: copied, no variants
        assert [(x["input"], len(x["variants"])) for x in r][2] == ("zzzzzzzzzz", 0)
        one = m.find_all_matches("I tink you are rihgt", A.SearchParameters(**{**TEST_PARAMS, "max_ngram": 1}))
        assert [x["input"] for x in one] == ["I", "tink", "you", "are", "rihgt"]
    finally:
        L.anl_find_all_matches = real
