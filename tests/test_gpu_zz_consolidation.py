"""GPU: find_all_matches with the sequence consolidation (most_likely_sequence, src/lib.rs:2088-2495, variant-model
scores only) end to end -- GPU lookups through anl_find_all_matches, host post-pass anl_match_set_consolidate --
against the oracle, the tutorial's recorded output and the reference's test 0702.  (The consolidation itself is
covered without a GPU in test_host_consolidation.py; this file checks the composition with real lookups.)"""
import ctypes as C
import json
import os
import struct

import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial.json"), encoding="utf-8"))


def bits(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


@pytest.fixture(scope="module")
def A():
    import analiticcl_b200
    return analiticcl_b200


@pytest.fixture(scope="module")
def eng(A):
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build()
    return m


def test_tutorial_sep_arate(A, eng):
    """tutorial.ipynb:476-481, unchanged call: default parameters (max_ngram = 3), the bigram is match 3."""
    g = GOLD["find_all_matches"]["We would like sep arate beds"]
    got = eng.find_all_matches("We would like sep arate beds", A.SearchParameters(unicodeoffsets=True))
    assert [m["input"] for m in got] == ["We", "would", "like", "sep arate", "beds"]
    m = got[g["index"]]
    assert (m["input"], m["offset"]) == (g["match"]["input"], g["match"]["offset"])
    assert [(v["text"], bits(v["dist_score"]), bits(v["score"])) for v in m["variants"]] == \
        [(v["text"], bits(v["dist_score"]), bits(v["score"])) for v in g["match"]["variants"]]
    # the producer's view is still there on request
    every = eng.find_all_matches("We would like sep arate beds", A.SearchParameters(unicodeoffsets=True, consolidate_matches=False))
    assert len(every) > len(got) and {"sep", "arate", "sep arate"} <= {m["input"] for m in every}


def test_reference_0702(A):
    """tests/main.rs:1143-1208 without its LM entries (the language model is out of scope)."""
    m = A.VariantModel(None, A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in ["I", "think", "sink", "you", "are", "right", "are right"]:
        m.add_to_vocabulary(w, 2, A.VocabParams())
    m.build()
    sp = A.SearchParameters(max_anagram_distance=2, max_edit_distance=2, max_matches=10, score_threshold=0.0,
                            cutoff_threshold=0.0, freq_weight=0.0, max_ngram=2)
    for text, last in (("I tink you are rihgt", "are rihgt"), ("I tink you are\nrihgt", "are\nrihgt")):
        r = m.find_all_matches(text, sp)
        assert [(x["input"], x["variants"][0]["text"]) for x in r] == \
            [("I", "I"), ("tink", "think"), ("you", "you"), (last, "are right")]
    assert r[1]["offset"] == {"begin": 2, "end": 6}


@pytest.mark.parametrize("max_ngram,freq_weight", [(2, 0.0), (3, 0.0), (3, 0.25)])
def test_consolidated_matches_equal_oracle(A, eng, eng_oracle, max_ngram, freq_weight, monkeypatch):
    from analiticcl_b200 import _capi
    monkeypatch.setenv("ANL_SEARCH_WINDOW", "300")  # several lookup windows
    text = workloads.cfg3_text(1500, 3002) + " It's a well-known co-operative re_entry; über naïve façade!?  Done" + \
        " qqqqqqqqqqqq xxxxxxxxxxxx. We would like sep arate beds to gether with out dis agree ment. A\nb"
    sp = A.SearchParameters(max_ngram=max_ngram, max_anagram_distance=2, max_edit_distance=2, freq_weight=freq_weight)
    op = orc.make_params(2, 2, 20, 0.25, 2.0, False, freq_weight, max_ngram, False)
    exp = eng_oracle.find_all_matches(text, op)
    L = _capi.lib()
    raw = text.encode("utf-8")
    ms, best = C.c_void_p(), C.c_void_p()
    assert L.anl_find_all_matches(eng._h, raw, len(raw), C.byref(sp.data), C.byref(ms)) == 0, L.anl_last_error()
    assert L.anl_match_set_consolidate(ms, raw, len(raw), C.byref(sp.data), C.byref(best)) == 0, L.anl_last_error()
    L.anl_match_set_free(ms)
    n = L.anl_match_set_len(best)
    m = _capi.Match()
    got = []
    for i in range(n):
        assert L.anl_match_set_get(best, i, C.byref(m)) == 0
        got.append((int(m.begin), int(m.end), int(m.n), int(m.selected),
                    [(m.variants[j].vocab_id, bits(m.variants[j].dist_score), bits(m.variants[j].freq_score))
                     for j in range(m.n_variants)] if m.variants else []))
    L.anl_match_set_free(best)
    want = [(s["begin"], s["end"], s["n"], s["selected"], [(v, bits(d), bits(f)) for v, d, f in s["variants"]]) for s in exp]
    bad = [i for i, (g, e) in enumerate(zip(got, want)) if g != e]
    assert len(got) == len(want) and not bad, (len(got), len(want), bad[:3], got[bad[0]] if bad else None)
    assert any(s["n"] > 1 for s in exp) and any(s["selected"] < 0 for s in exp)


def test_reference_0705_lm_disabled(A):
    """tests/main.rs:1364-1429 as it stands: the model carries the test's LM entries, `lm_weight = 0.0` switches the
    language model off, the sequence is decided by the variant-model cost alone."""
    m = A.VariantModel(None, A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in ["I", "think", "sink", "you", "are", "right", "are right"]:
        m.add_to_vocabulary(w, 2, A.VocabParams())
    for w, f in [("<bos> I", 2), ("I think", 2), ("I sink", 1), ("you are", 2), ("right <eos>", 2)]:
        m.add_to_vocabulary(w, f, A.VocabParams(vocabtype="LM"))
    m.build()
    sp = A.SearchParameters(max_anagram_distance=2, max_edit_distance=2, max_matches=10, score_threshold=0.0,
                            cutoff_threshold=0.0, freq_weight=0.0, max_ngram=2, lm_weight=0.0, context_weight=0.5)
    r = m.find_all_matches("I tink you are rihgt", sp)
    assert [(x["input"], x["variants"][0]["text"]) for x in r] == \
        [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right")]


# ---- language model and context rules: the reference's own tests through the Python mirror with GPU lookups ----------
WORDS = ["I", "think", "sink", "you", "are", "right"]
LM_ENTRIES = [("<bos> I", 2), ("I think", 2), ("I sink", 1), ("you are", 2), ("right <eos>", 2)]
TEST_SP = dict(max_anagram_distance=2, max_edit_distance=2, max_matches=10, score_threshold=0.0, cutoff_threshold=0.0,
               freq_weight=0.0, max_ngram=2)  # get_test_searchparams(), src/test.rs:48-68


def small(A, words, lm=()):
    m = A.VariantModel(None, A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in words:
        m.add_to_vocabulary(w, 2, A.VocabParams())
    for w, f in lm:
        m.add_to_vocabulary(w, f, A.VocabParams(vocabtype="LM"))
    m.build()
    return m


def test_reference_0702_0704_with_language_model(A):
    """tests/main.rs:1143-1361 as they stand (LM entries loaded, lm_weight = 1)."""
    m = small(A, WORDS + ["are right"], LM_ENTRIES)
    assert m.have_lm()
    sp = A.SearchParameters(**TEST_SP)
    for text, last in (("I tink you are rihgt", "are rihgt"), ("I tink you are\nrihgt", "are\nrihgt")):
        r = m.find_all_matches(text, sp)
        assert [(x["input"], x["variants"][0]["text"]) for x in r] == [("I", "I"), ("tink", "think"), ("you", "you"), (last, "are right")]
    m = small(A, WORDS + ["am", "sure", "are right"], LM_ENTRIES + [("I am", 2), ("sure <eos>", 2)])
    r = m.find_all_matches("I tink you are rihgt\n\nI am sur", sp)
    assert [(x["input"], x["variants"][0]["text"]) for x in r] == \
        [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right"), ("I", "I"), ("am", "am"), ("sur", "sure")]


def test_reference_0902_0905_context_rules(A):
    """tests/main.rs:1575-1728: bonus + tag, penalty, single-word rules sharing a tag, two tags on one rule."""
    sp = A.SearchParameters(**{**TEST_SP, "max_ngram": 1}, lm_weight=0.0)
    text = "I tink you are rihgt"
    m = small(A, WORDS)
    m.add_contextrule("I; think", 1.1, ["testtag"], [])
    r = m.find_all_matches(text, sp)
    assert [(x["input"], x["variants"][0]["text"]) for x in r] == [("I", "I"), ("tink", "think"), ("you", "you"), ("are", "are"), ("rihgt", "right")]
    assert (r[0]["tag"], r[0]["seqnr"], r[1]["tag"], r[1]["seqnr"]) == (["testtag"], [0], ["testtag"], [1])
    assert all("tag" not in x for x in r[2:])
    m = small(A, WORDS)
    m.add_contextrule("I; think", 0.9, [], [])
    r = m.find_all_matches(text, sp)
    assert [x["variants"][0]["text"] for x in r] == ["I", "sink", "you", "are", "right"]
    m = small(A, WORDS)
    for w in ("think", "are", "right"):
        m.add_contextrule(w, 1.0, ["testtag"], [])
    r = m.find_all_matches(text, sp)
    assert [(x.get("tag"), x.get("seqnr")) for x in r] == [(None, None), (["testtag"], [0]), (None, None), (["testtag"], [0]), (["testtag"], [0])]
    m = small(A, WORDS)
    m.add_contextrule("I; think", 1.1, ["testtag", "testtag2"], [])
    r = m.find_all_matches(text, sp)
    assert (r[0]["tag"], r[0]["seqnr"], r[1]["tag"], r[1]["seqnr"]) == (["testtag", "testtag2"], [0, 0], ["testtag", "testtag2"], [1, 1])
    assert [x["variants"][0]["text"] for x in r] == ["I", "think", "you", "are", "right"]


def test_language_model_and_rules_equal_oracle_on_running_text(A, eng, eng_oracle, tmp_path):
    """A bigram model counted from the text itself + two rules, on the eng lexicon: GPU lookups + host sequence stage
    against the oracle end to end (the eng fixtures are shared: the LM is loaded into fresh models)."""
    text = workloads.cfg3_text(400, 77) + " We would like sep arate beds to gether."
    toks = [t for t in text.replace(".", " ").split() if t]
    counts = {}
    for a, b in zip(toks, toks[1:]):
        counts[a + " " + b] = counts.get(a + " " + b, 0) + 1
    lmf = tmp_path / "lm.tsv"
    lmf.write_text("".join(f"{k}\t{v}\n" for k, v in sorted(counts.items())[:300]), encoding="utf-8")
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    for mm in (o, m):
        mm.read_lexicon(workloads.lexicon_path("eng"))
        mm.read_lm(str(lmf))
        mm.build()
        mm.add_contextrule("would; like", 1.1, ["modal"], [])
        mm.add_contextrule("^; ?", 0.95, [], [])
    assert m.have_lm() and o.have_lm()
    for kw in (dict(max_ngram=2), dict(max_ngram=3, max_seq=20, lm_weight=2.0), dict(max_ngram=1, contextrules_weight=0.5)):
        sp = A.SearchParameters(max_anagram_distance=2, max_edit_distance=2, **kw)
        op = orc.make_params(max_anagram_distance=2, max_edit_distance=2, **kw)
        exp = o.find_all_matches(text, op)
        got = m.find_all_matches(text, sp)
        assert [(x["input"], x["offset"]["begin"], x["offset"]["end"]) for x in got] == [(s["text"], s["begin"], s["end"]) for s in exp]
        for g, s in zip(got, exp):
            ev = [o.vocab_text(v[0]) for v in s["variants"]]
            if s["selected"] > 0:
                ev.insert(0, ev.pop(s["selected"]))
            assert [v["text"] for v in g["variants"]] == ev
            assert g.get("tag", []) == [o.tags()[t] for t in s["tag"]] and g.get("seqnr", []) == s["seqnr"]
