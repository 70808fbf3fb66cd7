"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Compared per query: the list of (vocab_id, dist_score, freq_score) -- same candidates, same order,
identical f64 bits.  Run on the B200 box: `pytest -m gpu`.
"""
import json
import os
import struct

import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu


def bits(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


def assert_same(got, exp, queries, tag=""):
    assert len(got) == len(exp)
    bad = []
    for i, (g, e) in enumerate(zip(got, exp)):
        gg = [(v, bits(d), bits(f)) for v, d, f in g]
        ee = [(v, bits(d), bits(f)) for v, d, f in e]
        if gg != ee:
            bad.append((i, queries[i], g[:4], e[:4], len(g), len(e)))
    assert not bad, f"{tag}: {len(bad)} / {len(got)} queries differ, first: {bad[:3]}"


def to_orc_params(sp):
    d = sp.data

    def thr(t):
        return {0: float(t.ratio), 1: (float(t.ratio), int(t.value)), 2: int(t.value)}[t.kind]
    return orc.make_params(thr(d.max_anagram_distance), thr(d.max_edit_distance), d.max_matches, d.score_threshold,
                           d.cutoff_threshold, d.stop_criterion == 1, d.freq_weight, d.max_ngram, bool(d.unicodeoffsets))


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


@pytest.fixture(scope="module")
def nld_pair(A):
    path = workloads.nld_freq_lexicon()
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(path)
    m.build()
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o.read_lexicon(path)
    o.build()
    return m, o


GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial.json"), encoding="utf-8"))


def test_index_goldens(eng):
    assert eng.instance_count() == 119773 and eng.index_size() == 108802
    hist = [27, 248, 942, 2593, 5623, 10163, 14617, 16911, 16391, 13930, 10650, 7194, 4434, 2459, 1384, 667, 339, 128,
            62, 20, 9, 8, 2, 1]
    assert [eng.anagram_count_of_length(i + 1) for i in range(24)] == hist
    assert eng.max_key_bits() == 102
    assert "separate" in eng and "seperate" not in eng


@pytest.mark.parametrize("q", ["separate", "seperate"])
def test_tutorial_find_variants(A, eng, q):
    got = eng.find_variants(q, A.SearchParameters())
    gold = GOLD["find_variants"][q]
    assert [g["text"] for g in got] == [g["text"] for g in gold]
    for g, e in zip(got, gold):
        assert bits(g["dist_score"]) == bits(e["dist_score"]) and bits(g["score"]) == bits(e["score"])
        assert g["freq_score"] == e["freq_score"] and len(g["lexicons"]) == 1


def test_tutorial_find_all_matches(A, eng):
    got = eng.find_all_matches("We would like seperate beds", A.SearchParameters(unicodeoffsets=True, max_ngram=1))
    gold = GOLD["find_all_matches"]["We would like seperate beds"]
    assert [(g["input"], g["offset"]) for g in got] == [(e["input"], e["offset"]) for e in gold]
    for g, e in zip(got, gold):
        assert [(v["text"], bits(v["dist_score"])) for v in g["variants"]] == \
               [(v["text"], bits(v["dist_score"])) for v in e["variants"]]
    # the 1-ulp ranking cases of "sep arate" (tutorial.ipynb:476)
    e = GOLD["find_all_matches"]["We would like sep arate beds"]["match"]
    g = eng.find_variants("sep arate", A.SearchParameters())
    assert [(v["text"], bits(v["dist_score"])) for v in g] == [(v["text"], bits(v["dist_score"])) for v in e["variants"]]


CASES = [
    ("cfg1 k2", dict(max_anagram_distance=2, max_edit_distance=2), 3000, 1001),
    ("default k3", dict(), 1500, 77),
    ("k4", dict(max_anagram_distance=4, max_edit_distance=4), 300, 4001),
    ("ratio", dict(max_anagram_distance=0.3, max_edit_distance=(0.4, 3)), 500, 5),
    ("k3 edit2 unlimited", dict(max_edit_distance=2, max_matches=0, score_threshold=0.0, cutoff_threshold=0.0), 500, 6),
    ("crop ties", dict(max_matches=3, score_threshold=0.0, cutoff_threshold=0.0), 800, 7),
    ("crop 1", dict(max_matches=1), 500, 8),
    ("stop at exact", dict(stop_at_exact_match=True), 600, 9),
    ("freq weight", dict(freq_weight=0.3), 500, 10),
    ("k1", dict(max_anagram_distance=1, max_edit_distance=1), 500, 11),
    ("anagram>edit", dict(max_anagram_distance=3, max_edit_distance=1), 500, 12),
]


@pytest.mark.parametrize("tag,kw,n,seed", CASES, ids=[c[0] for c in CASES])
def test_eng_parity(A, eng, eng_oracle, tag, kw, n, seed):
    words = workloads.read_words("eng")
    qs = workloads.misspellings(words, n, seed, min_len=2, max_len=18, edit_probs=((0, 0.15), (1, 0.45), (2, 0.3), (3, 0.1)))
    sp = A.SearchParameters(**kw)
    got = eng.find_variants_raw(qs, sp)
    exp = eng_oracle.find_variants_batch(qs, to_orc_params(sp), threads=0)
    assert_same(got, exp, qs, tag)


def test_edge_inputs(A, eng, eng_oracle):
    qs = ["a", "I", "zz", "x" * 30, "Frankfurt", "FRANK", "naïve", "ÆØÅ", "'s", "don't", "12345", "hello world",
          "the quick brown fox", "œuvre", "ab" * 60, "q" * 240, "Ω≈ç√", " ", "-", "résumé", "e" * 9, "st" * 5,
          "pneumonoultramicroscopicsilicovolcanoconiosis", "Stael", "Tesla", "steal"]
    sp = A.SearchParameters()
    got = eng.find_variants_raw(qs, sp)
    exp = eng_oracle.find_variants_batch(qs, to_orc_params(sp))
    assert_same(got, exp, qs, "edge")


def test_empty_query_flag(A, eng):
    out = eng.find_variants_par(["", "frog", ""], A.SearchParameters())
    assert [o["variants"] == [] for o in out] == [True, False, True]


def test_capacity_overflow_rerun(A, eng_oracle, monkeypatch):
    """Fixed-capacity hit / result buffers overflow -> flagged queries are re-run with exact sizes."""
    monkeypatch.setenv("ANL_HIT_CAP", "16")
    monkeypatch.setenv("ANL_OUT_CAP", "4")
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build()
    qs = workloads.misspellings(workloads.read_words("eng"), 400, 99, min_len=3, max_len=10)
    sp = A.SearchParameters()
    got = m.find_variants_raw(qs, sp)
    exp = eng_oracle.find_variants_batch(qs, to_orc_params(sp))
    assert_same(got, exp, qs, "overflow")


def test_pool_overflow_and_reruns_with_confusables(A, nld_pair, monkeypatch):
    """A result pool that is too small (the score stage runs again with the exact size) together with hit-list
    overflows (re-run, patched into the pool, exported again) on a model with device-side confusables."""
    monkeypatch.setenv("ANL_POOL_PER_QUERY", "1")
    monkeypatch.setenv("ANL_HIT_CAP", "64")
    monkeypatch.setenv("ANL_CHUNK", "4096")
    _, o = nld_pair
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.nld_freq_lexicon())
    o2 = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o2.read_lexicon(workloads.nld_freq_lexicon())
    for pat, w in workloads.CFG2_CONFUSABLES:
        m.add_to_confusables(pat, w)
        o2.add_to_confusables(pat, w)
    m.build()
    o2.build()
    qs = workloads.ocr_noise(workloads.read_words("nld"), 9000, 4711)
    sp = A.SearchParameters(freq_weight=0.25)
    assert_same(m.find_variants_raw(qs, sp), o2.find_variants_batch(qs, to_orc_params(sp)), qs, "pool overflow + reruns + confusables")


@pytest.mark.parametrize("inflight", ["1", "4"])
def test_chunk_pipeline_with_reruns(A, eng_oracle, monkeypatch, inflight):
    """The batch call cuts the queries into chunks with several batches in flight and fetches them in two phases;
    chunks whose hit lists overflow get an asynchronous re-run that is collected rounds later.  Small chunks and a
    small capacity make some chunks overflow and others not: the results must still be the oracle's, in order."""
    monkeypatch.setenv("ANL_CHUNK", "1024")
    monkeypatch.setenv("ANL_INFLIGHT", inflight)
    monkeypatch.setenv("ANL_HIT_CAP", "96")
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build()
    words = workloads.read_words("eng")
    # long words (few candidates: no overflow) and short ones (many candidates) in alternating stretches
    long_q = workloads.misspellings(words, 6000, 7, min_len=12, max_len=24)
    short_q = workloads.misspellings(words, 6000, 8, min_len=3, max_len=6)
    qs = []
    for k in range(0, 6000, 1500):
        qs += long_q[k:k + 1500] + short_q[k:k + 1500]
    qs = qs[:11000] + [""] + qs[11000:]
    sp = A.SearchParameters()
    got = m.find_variants_raw(qs, sp)
    exp = eng_oracle.find_variants_batch(qs, to_orc_params(sp), threads=0)
    assert_same(got, exp, qs, f"pipeline inflight={inflight}")


@pytest.mark.parametrize("per_query", ["3", "100000"], ids=["queue overflow -> fused rerun", "roomy queue"])
def test_split_probe_queue(A, eng, eng_oracle, monkeypatch, per_query):
    """Split probe path (Bloom stage -> global queue of staged nodes -> exact stage): a queue that is too small is
    detected after the run and answered by the fused probe kernel; either way the results are the oracle's."""
    monkeypatch.setenv("ANL_QUEUE_PER_QUERY", per_query)
    qs = workloads.misspellings(workloads.read_words("eng"), 700, 4321)
    sp = A.SearchParameters()
    assert_same(eng.find_variants_raw(qs, sp), eng_oracle.find_variants_batch(qs, to_orc_params(sp)), qs, "split queue")
    monkeypatch.setenv("ANL_SPLIT", "0")  # (read once per process: only effective if this is the first batch)


def test_nld_frequency_parity(A, nld_pair):
    m, o = nld_pair
    qs = workloads.ocr_noise(workloads.read_words("nld"), 1500, 2003)
    for kw in (dict(freq_weight=0.25), dict(), dict(freq_weight=0.25, max_matches=5)):
        sp = A.SearchParameters(**kw)
        assert_same(m.find_variants_raw(qs, sp), o.find_variants_batch(qs, to_orc_params(sp)), qs, f"nld {kw}")


@pytest.mark.parametrize("early", [False, True], ids=["late", "early"])
def test_nld_confusables_parity(A, early):
    path = workloads.nld_freq_lexicon()
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(path)
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o.read_lexicon(path)
    for pat, w in workloads.CFG2_CONFUSABLES:
        m.add_to_confusables(pat, w)
        o.add_to_confusables(pat, w)
    if early:
        m.set_confusables_before_pruning()
        o.set_confusables_before_pruning()
    m.build()
    o.build()
    qs = workloads.ocr_noise(workloads.read_words("nld"), 800, 2004)
    sp = A.SearchParameters(freq_weight=0.25)
    assert_same(m.find_variants_raw(qs, sp), o.find_variants_batch(qs, to_orc_params(sp)), qs, "confusables")


RICH_CONFUSABLES = [  # identities as context, anchors, alternatives, multi-character and non-ASCII options
    ("-[f]+[s]", 1.1), ("-[y]+[i]", 1.1), ("-[c]+[e]", 0.95), ("-[l]+[i]", 1.05), ("-[u]+[n]", 0.95),
    ("=[s]-[c]+[e]", 1.2), ("^-[b]+[h]", 1.15), ("-[rn]+[m]", 1.3), ("-[m]+[rn]", 1.25), ("+[e|n]$", 0.9),
    ("-[ij]+[y]", 1.07), ("-[a|e|i|o|u]+[a|e|i|o|u]=[n|r|s|t]", 0.97), ("-[é|e]+[ë|a]", 1.4), ("=[ge]-[l]", 0.8),
    ("^=[ver|be|ge|on]+[s|t]", 1.12), ("-[i]=[l]+[i]", 1.5), ("-[ë]+[e]", 1.21),
]


@pytest.mark.parametrize("early", [False, True], ids=["late", "early"])
@pytest.mark.parametrize("lex", ["nld", "eng"])
def test_device_confusable_stage_parity(A, lex, early):
    """The confusable kernel (edit script per thread) + finish kernel against the oracle: rich pattern set,
    OCR noise and misspellings, non-ASCII and over-long queries (host post-pass) mixed into one batch."""
    path = workloads.nld_freq_lexicon() if lex == "nld" else workloads.lexicon_path("eng")
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(path)
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o.read_lexicon(path)
    for pat, w in RICH_CONFUSABLES:
        m.add_to_confusables(pat, w)
        o.add_to_confusables(pat, w)
    if early:
        m.set_confusables_before_pruning()
        o.set_confusables_before_pruning()
    m.build()
    o.build()
    words = workloads.read_words(lex)
    qs = workloads.ocr_noise(words, 1500, 77) + workloads.misspellings(words, 700, 78)
    qs += ["x" * 70, "aan" * 30, "één", "zeeën", "reëel", "cafe", "café", "s", "ij", "geld", "verstaan", "bestaan"]
    qs += [w for w in words[::4001]][:60]
    sp = A.SearchParameters(freq_weight=0.25 if lex == "nld" else 0.0, max_matches=7)
    assert_same(m.find_variants_raw(qs, sp), o.find_variants_batch(qs, to_orc_params(sp)), qs, f"device confusables {lex}")


def small(A, words, confusables=(), weights=None):
    m = A.VariantModel(None, weights or A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in words:
        m.add_to_vocabulary(w, None, A.VocabParams())
    for p, wt in confusables:
        m.add_to_confusables(p, wt)
    m.build()
    return m


TEST_PARAMS = dict(max_anagram_distance=2, max_edit_distance=2, max_matches=10, score_threshold=0.0,
                   cutoff_threshold=0.0, freq_weight=0.0, max_ngram=2)


def test_reference_model_kats(A):
    """tests/main.rs:858-911, 935-1020, 1120-1140 through the product API."""
    m = small(A, ["rites", "tiers", "tires", "tries", "tyres", "rides", "brides", "dire"])
    assert all(w in m for w in ["rites", "tiers", "dire"]) and "unknown" not in m
    r = m.find_variants("rite", A.SearchParameters(**TEST_PARAMS))
    assert [(v["text"], v["dist_score"]) for v in r] == [("rites", 0.75), ("dire", 0.4375)]
    m = small(A, ["huis", "huls"])
    r = m.find_variants("huys", A.SearchParameters(**TEST_PARAMS))
    assert [v["text"] for v in r] == ["huis", "huls"] and r[0]["dist_score"] == r[1]["dist_score"]
    for q in ("huys", "Huys"):
        m = small(A, ["huis", "huls"], [("-[y]+[i]", 1.1)])
        r = m.find_variants(q, A.SearchParameters(**TEST_PARAMS))
        assert [v["text"] for v in r] == ["huis", "huls"] and r[0]["dist_score"] > r[1]["dist_score"]
    m = small(A, ["huis", "huls"], [("-[y]+[p]", 1.1)])
    r = m.find_variants("Huys", A.SearchParameters(**TEST_PARAMS))
    assert len(r) == 2 and r[0]["dist_score"] == r[1]["dist_score"]
    m = small(A, ["I", "think", "sink", "you", "are", "right"])
    r = m.find_all_matches("I tink you are rihgt", A.SearchParameters(**{**TEST_PARAMS, "max_ngram": 1}))
    assert [x["input"] for x in r] == ["I", "tink", "you", "are", "rihgt"]
    assert r[1]["variants"][0]["text"] == "think" and r[4]["variants"][0]["text"] == "right"
    m = small(A, ["I", "think", "you", "are", "right"])
    r = m.find_all_matches("I thиnk you are righт", A.SearchParameters(**{**TEST_PARAMS, "max_ngram": 1}, unicodeoffsets=True))
    assert (r[1]["input"], r[1]["offset"]) == ("thиnk", {"begin": 2, "end": 7}) and r[1]["variants"][0]["text"] == "think"
    r = m.find_all_matches("I thиnk you are rihgt", A.SearchParameters(**{**TEST_PARAMS, "max_ngram": 1}))
    assert r[1]["offset"] == {"begin": 2, "end": 8} and r[4]["variants"][0]["text"] == "right"


def test_python_binding_test(A):
    """bindings/python/tests/tests.py:12-26, unchanged apart from the import."""
    m = A.VariantModel(workloads.ALPHABET, A.Weights(), debug=False)
    m.read_lexicon(workloads.AMPHIBIANS)
    m.read_lexicon(workloads.REPTILES)
    m.build()
    results = m.find_all_matches("Salamander lizard frog snake toad", A.SearchParameters(max_edit_distance=3, max_ngram=1))
    assert len(results) == 5
    exp = [("Salamander", workloads.AMPHIBIANS, "salamander"), ("lizard", workloads.REPTILES, "lizard"),
           ("frog", workloads.AMPHIBIANS, "frog"), ("snake", workloads.REPTILES, "snake"),
           ("toad", workloads.AMPHIBIANS, "toad")]
    for r, (orig, lexicon, term) in zip(results, exp):
        assert r["input"] == orig and r["variants"][0]["text"] == term and r["variants"][0]["lexicons"] == [lexicon]


def test_python_binding_test_under_its_own_module_name(tmp_path, monkeypatch):
    """The same test with the reference's import line (`from analiticcl import VariantModel, Weights,
    SearchParameters`, bindings/python/tests/tests.py:3) and its relative paths, run from a directory laid out like
    bindings/python/ (tests/*.tsv next to it, ../../examples/simple.alphabet.tsv above)."""
    import shutil
    from analiticcl import VariantModel, Weights, SearchParameters
    py = tmp_path / "bindings" / "python"
    (py / "tests").mkdir(parents=True)
    (tmp_path / "examples").mkdir()
    shutil.copy(workloads.AMPHIBIANS, py / "tests" / "amphibians.tsv")
    shutil.copy(workloads.REPTILES, py / "tests" / "reptiles.tsv")
    shutil.copy(workloads.ALPHABET, tmp_path / "examples" / "simple.alphabet.tsv")
    monkeypatch.chdir(py)
    model = VariantModel("../../examples/simple.alphabet.tsv", Weights(), debug=False)
    model.read_lexicon("tests/amphibians.tsv")
    model.read_lexicon("tests/reptiles.tsv")
    model.build()
    results = model.find_all_matches("Salamander lizard frog snake toad", SearchParameters(max_edit_distance=3, max_ngram=1))
    assert len(results) == 5
    for r, (orig, lexicon, term) in zip(results, [("Salamander", "tests/amphibians.tsv", "salamander"),
                                                  ("lizard", "tests/reptiles.tsv", "lizard"), ("frog", "tests/amphibians.tsv", "frog"),
                                                  ("snake", "tests/reptiles.tsv", "snake"), ("toad", "tests/amphibians.tsv", "toad")]):
        assert r["input"] == orig and len(r["variants"]) > 0
        assert r["variants"][0]["text"] == term and r["variants"][0]["lexicons"] == [lexicon]


def test_weights_variants(A):
    """Features with weight <= 0 are skipped (src/lib.rs:1352-1377); non-default weights stay bit-exact."""
    words = workloads.read_words("eng")[:30000]
    qs = workloads.misspellings(words, 400, 31, min_len=3, max_len=12)
    for wkw in (dict(lcs=0.0, case=0.0), dict(ld=1.0, lcs=0.3, prefix=0.0, suffix=0.2, case=0.05)):
        m = A.VariantModel(workloads.ALPHABET, A.Weights(**wkw))
        ww = A.Weights(**wkw)
        o = orc.OracleModel(alphabet_file=workloads.ALPHABET, weights=(ww.ld, ww.lcs, ww.prefix, ww.suffix, ww.case))
        for w in words:
            m.add_to_vocabulary(w, None, A.VocabParams())
            o.add_to_vocabulary(w)
        m.build()
        o.build()
        sp = A.SearchParameters()
        assert_same(m.find_variants_raw(qs, sp), o.find_variants_batch(qs, to_orc_params(sp)), qs, f"weights {wkw}")


@pytest.mark.parametrize("max_ngram", [1, 2, 3])
def test_find_all_matches_segments_parity(A, eng, eng_oracle, max_ngram, monkeypatch):
    """find_all_matches batch producer (src/lib.rs:1790-1903): same segments (offsets, order), same
    redundant-match pruning, same variant lists as the reference algorithm -- over several windows."""
    import ctypes as C
    from analiticcl_b200 import _capi
    monkeypatch.setenv("ANL_SEARCH_WINDOW", "300")  # force several windows
    text = workloads.cfg3_text(2500, 3001) + " It's a well-known co-operative re_entry; über naïve façade!?  Done"
    sp = A.SearchParameters(max_ngram=max_ngram, max_anagram_distance=2, max_edit_distance=2)
    raw = text.encode("utf-8")
    ms = C.c_void_p()
    L = _capi.lib()
    assert L.anl_find_all_matches(eng._h, raw, len(raw), C.byref(sp.data), C.byref(ms)) == 0, L.anl_last_error()
    exp = eng_oracle.find_all_segments(text, to_orc_params(sp))
    n = L.anl_match_set_len(ms)
    assert n == len(exp)
    m = _capi.Match()
    bad = []
    for i in range(n):
        assert L.anl_match_set_get(ms, i, C.byref(m)) == 0
        e = exp[i]
        got_vars = [(m.variants[j].vocab_id, bits(m.variants[j].dist_score), bits(m.variants[j].freq_score))
                    for j in range(m.n_variants)] if m.variants else None
        exp_vars = [(v, bits(d), bits(f)) for v, d, f in e["variants"]] if e["looked_up"] else None
        if (m.begin, m.end, m.n) != (e["begin"], e["end"], e["n"]) or got_vars != exp_vars:
            bad.append((i, e["text"], (m.begin, m.end, m.n), (e["begin"], e["end"], e["n"])))
    L.anl_match_set_free(ms)
    assert not bad, bad[:5]
    if max_ngram > 1:
        assert any(not e["looked_up"] for e in exp) and any(e["looked_up"] and e["n"] > 1 for e in exp)


def _all_matches_raw(A, model, text, sp):
    import ctypes as C
    from analiticcl_b200 import _capi
    L = _capi.lib()
    raw = text.encode("utf-8")
    ms = C.c_void_p()
    assert L.anl_find_all_matches(model._h, raw, len(raw), C.byref(sp.data), C.byref(ms)) == 0, L.anl_last_error()
    out = []
    m = _capi.Match()
    for i in range(L.anl_match_set_len(ms)):
        assert L.anl_match_set_get(ms, i, C.byref(m)) == 0
        vs = tuple((m.variants[j].vocab_id, bits(m.variants[j].dist_score), bits(m.variants[j].freq_score))
                   for j in range(m.n_variants)) if m.variants else None
        out.append((int(m.begin), int(m.end), int(m.n), int(m.selected), vs))
    a, b = C.c_uint64(), C.c_uint64()
    L.anl_match_set_lookup_counts(ms, C.byref(a), C.byref(b))
    L.anl_match_set_free(ms)
    return out, a.value, b.value


@pytest.mark.gpu
def test_find_all_matches_large_text_equals_piecewise(A, eng):
    """Size-independent property for the multi-threaded producer (parallel boundary scan stitched across
    byte ranges, hash-partitioned de-duplication, parallel assembly): the matches of a long text equal the
    matches of its sentence groups looked up one small call at a time (single-threaded paths), shifted."""
    sents = workloads.cfg3_text(60000, 77).split(". ")
    extra = "It's a well-known co-operative re_entry über naïve façade"
    pieces, cur = [], []
    for i, s in enumerate(sents):
        cur.append(s if i % 97 else s + " " + extra)
        if len(cur) == 150:
            pieces.append(". ".join(cur) + ". ")
            cur = []
    if cur:
        pieces.append(". ".join(cur) + ". ")
    text = "".join(pieces)
    assert len(text.encode("utf-8")) > 400_000 and all(len(p.encode("utf-8")) < 60_000 for p in pieces)
    sp = A.SearchParameters(max_ngram=3, max_anagram_distance=2, max_edit_distance=2)
    whole, lookups, distinct = _all_matches_raw(A, eng, text, sp)
    assert 0 < distinct < lookups  # running text repeats itself: the producer de-duplicates
    exp, shift = [], 0
    for p in pieces:
        part, _, _ = _all_matches_raw(A, eng, p, sp)
        exp += [(b + shift, e + shift, n, sel, vs) for b, e, n, sel, vs in part]
        shift += len(p.encode("utf-8"))
    assert len(whole) == len(exp)
    bad = [i for i, (g, e) in enumerate(zip(whole, exp)) if g != e]
    assert not bad, (len(bad), whole[bad[0]], exp[bad[0]])


# ---- variant lists (SURVEY 8 "next" row f-4; src/lib.rs:460-514, 766-897, 1677-1727) ------------------------
def _variant_list(words, path, seed, with_freq):
    """A synthetic error list: ~3000 references from the lexicon, each with 1-3 misspelt variants and scores."""
    import numpy as np
    rng = np.random.default_rng(seed)
    refs = [words[int(i)] for i in rng.choice(len(words), size=3000, replace=False) if len(words[int(i)]) >= 5]
    lines = []
    for k, r in enumerate(refs):
        vs = workloads.misspellings([r], int(rng.integers(1, 4)), seed + k, min_len=1, max_len=99)
        cols = [r] + (["%d" % int(rng.integers(1, 1000))] if with_freq else [])
        for v in vs:
            cols += [v, "%.2f" % float(rng.uniform(0.3, 1.0))] + (["%d" % int(rng.integers(1, 50))] if with_freq else [])
        lines.append("\t".join(cols))
    lines.insert(5, lines[0])  # a repeated line: the (reference, variant) links are stored again (reference quirk)
    path.write_text("\n".join(lines) + "\n")


def assert_same_via(got, exp, queries, tag):
    bad = []
    for i, (g, e) in enumerate(zip(got, exp)):
        if [(v, bits(d), bits(f), via) for v, d, f, via in g] != [(v, bits(d), bits(f), via) for v, d, f, via in e]:
            bad.append((i, queries[i], g[:4], e[:4], len(g), len(e)))
    assert len(got) == len(exp) and not bad, f"{tag}: {len(bad)} / {len(got)} queries differ, first: {bad[:3]}"


@pytest.mark.parametrize("transparent,with_freq,confusables", [(False, False, None), (True, False, None), (True, True, "late"),
                                                             (False, True, "early")])
def test_variant_lists(A, tmp_path, transparent, with_freq, confusables):
    words = workloads.read_words("eng")
    f = tmp_path / "variants.tsv"
    _variant_list(words, f, 17, with_freq)
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    lex = workloads.lexicon_path("eng")
    if with_freq:
        lex = workloads.nld_freq_lexicon()  # a lexicon with a frequency column (have_freq = true)
        words = workloads.read_words("nld")
        _variant_list(words, f, 17, with_freq)
    m.read_lexicon(lex)
    o.read_lexicon(lex)
    m.read_variants(str(f), transparent=transparent)
    o.read_variants(str(f), transparent=transparent)
    if confusables:
        for pat, wt in workloads.CFG2_CONFUSABLES:
            m.add_to_confusables(pat, wt)
            o.add_to_confusables(pat, wt)
        if confusables == "early":
            m.set_confusables_before_pruning()
            o.set_confusables_before_pruning()
    m.build()
    o.build()
    # queries: the variants themselves, further misspellings of them, and ordinary misspellings
    listed = [c for ln in f.read_text().splitlines() for c in ln.split("\t")[(2 if with_freq else 1)::(3 if with_freq else 2)]]
    qs = listed[:1500] + workloads.misspellings(listed, 1000, 5, min_len=1, max_len=99) + workloads.misspellings(words, 1000, 6)
    for kw in (dict(), dict(max_matches=3, freq_weight=0.25), dict(max_matches=0, cutoff_threshold=1.5)):
        sp = A.SearchParameters(**kw)
        got = m.find_variants_raw(qs, sp, with_via=True)
        exp = o.find_variants_batch(qs, to_orc_params(sp), threads=0, with_via=True)
        assert_same_via(got, exp, qs, f"variants transparent={transparent} freq={with_freq} conf={confusables} {kw}")
    assert any(via is not None for r in got for *_, via in r)
