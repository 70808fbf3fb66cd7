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
