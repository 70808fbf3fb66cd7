"""Pins the CPU oracle (oracle/oracle.cpp) against the reference's own known-answer tests and
documentation goldens.  Citations: /root/reference/tests/main.rs (T:line), tutorial.ipynb, README.md.
CPU only."""
import json
import os

import pytest

import workloads
from oracle import orc

ALPHA_SIZE = 27  # get_test_alphabet() returns (alphabet, 27)  (src/test.rs:33-46)


@pytest.fixture(scope="module")
def tm():
    return orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)


def ah(tm, s):
    return tm.anahash(s)


# ---- 01xx hashing (T:30-153) -----------------------------------------------------------------
def test_hash_basic(tm):
    assert ah(tm, "a") == 2 and ah(tm, "b") == 3 and ah(tm, "c") == 5
    assert ah(tm, "ab") == 6 == ah(tm, "ba")
    assert ah(tm, "abc") == 30
    assert ah(tm, "abcabcabc") == 30 ** 3


def test_hash_alphabet_equivalence(tm):
    assert ah(tm, "abc") == ah(tm, "ABC") == ah(tm, "bAc")
    assert ah(tm, "a.b") == ah(tm, "a,b")


def test_hash_big(tm):
    v = ah(tm, "xyz" * 24)
    assert v == (89 * 97 * 101) ** 24 and v > 2 ** 64


def test_hash_anagram(tm):
    assert ah(tm, "stressed") == ah(tm, "desserts")
    assert ah(tm, "dormitory") == ah(tm, "dirtyroom")
    assert ah(tm, "presents") == ah(tm, "serpents")


def test_hash_insert_contains_delete(tm):
    ab, b, c, abc, ac, x = (ah(tm, s) for s in ("ab", "b", "c", "abc", "ac", "x"))
    assert orc.ana_insert(ab, c) == abc == orc.ana_insert(c, ab)
    assert orc.ana_contains(abc, c) and orc.ana_contains(abc, ab) and orc.ana_contains(abc, abc)
    assert not orc.ana_contains(c, abc) and not orc.ana_contains(ab, c) and not orc.ana_contains(ab, abc)
    assert orc.ana_delete(abc, c) == ab and orc.ana_delete(abc, b) == ac
    assert orc.ana_delete(c, abc) is None and orc.ana_delete(abc, x) is None
    assert orc.ana_insert(0, c) == c  # insert into zero yields the value (src/anahash.rs:146)


def test_hash_upper_bound(tm):
    assert orc.alphabet_upper_bound(ah(tm, "abc"), ALPHA_SIZE) == (2, 3)
    assert orc.alphabet_upper_bound(ah(tm, "ab"), ALPHA_SIZE) == (1, 2)
    assert orc.alphabet_upper_bound(ah(tm, "x"), ALPHA_SIZE) == (23, 1)


# ---- 02xx iterators, exact yield order (T:156-556) ---------------------------------------------
PR = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101, 103]


def test_iterator_parents(tm):
    d = orc.deletions(ah(tm, "house"), ALPHA_SIZE, "parents")
    assert [PR[c] for _, _, c in d] == [ah(tm, x) for x in "usohe"]
    assert [v for v, _, _ in d] == [ah(tm, x) for x in ("hose", "houe", "huse", "ouse", "hous")]


def test_iterator_parents_dup(tm):
    d = orc.deletions(ah(tm, "pass"), ALPHA_SIZE, "parents")
    assert [PR[c] for _, _, c in d] == [ah(tm, x) for x in "spa"]
    assert [v for v, _, _ in d] == [ah(tm, x) for x in ("pas", "ass", "pss")]


def test_iterator_singlebeam(tm):
    d = orc.deletions(ah(tm, "house"), ALPHA_SIZE, "singlebeam")
    assert [PR[c] for _, _, c in d] == [ah(tm, x) for x in "usohe"]
    assert [v for v, _, _ in d] == [ah(tm, "hose"), ah(tm, "hoe"), ah(tm, "he"), ah(tm, "e"), 1]
    assert [dep for _, dep, _ in d] == [1, 2, 3, 4, 5]


def _vals(tm, words):
    return [1 if w == "" else ah(tm, w) for w in words]


def test_iterator_recursive_dfs(tm):
    d = orc.deletions(ah(tm, "abcd"), ALPHA_SIZE, "recursive")
    exp = ["abc", "ab", "a", "", "b", "", "ac", "a", "", "c", "", "bc", "b", "", "c", "", "abd", "ab", "a"]
    assert [v for v, _, _ in d][:len(exp)] == _vals(tm, exp)


def test_iterator_recursive_no_empty_leaves(tm):
    d = orc.deletions(ah(tm, "abcd"), ALPHA_SIZE, "recursive", allow_empty_leaves=False)
    exp = ["abc", "ab", "a", "b", "ac", "a", "c", "bc", "b", "c", "abd", "ab", "a"]
    assert [v for v, _, _ in d][:len(exp)] == _vals(tm, exp)


def test_iterator_recursive_no_duplicates(tm):
    d = orc.deletions(ah(tm, "abcd"), ALPHA_SIZE, "recursive", allow_empty_leaves=False, allow_duplicates=False)
    exp = ["abc", "ab", "a", "b", "ac", "c", "bc", "abd"]
    assert [v for v, _, _ in d][:len(exp)] == _vals(tm, exp)


def test_iterator_recursive_bfs(tm):
    d = orc.deletions(ah(tm, "abcd"), ALPHA_SIZE, "recursive", breadthfirst=True)
    exp = ["abc", "abd", "acd", "bcd", "ab", "ac", "bc", "ab", "ad", "bd", "ac", "ad", "cd", "bc", "bd", "cd",
           "a", "b", "a", "c"]
    dep = [1] * 4 + [2] * 12 + [3] * 4
    assert [v for v, _, _ in d][:len(exp)] == _vals(tm, exp)
    assert [x for _, x, _ in d][:len(dep)] == dep


@pytest.mark.parametrize("maxd,exp,dep", [
    (-1, ["abc", "abd", "acd", "bcd", "ab", "ac", "bc", "ad", "bd", "cd", "a", "b", "c", "d"], [1] * 4 + [2] * 6 + [3] * 4),
    (3, ["abc", "abd", "acd", "bcd", "ab", "ac", "bc", "ad", "bd", "cd", "a", "b", "c", "d"], [1] * 4 + [2] * 6 + [3] * 4),
    (2, ["abc", "abd", "acd", "bcd", "ab", "ac", "bc", "ad", "bd", "cd"], [1] * 4 + [2] * 6),
])
def test_iterator_recursive_bfs_unique(tm, maxd, exp, dep):
    d = orc.deletions(ah(tm, "abcd"), ALPHA_SIZE, "recursive", maxdepth=maxd, breadthfirst=True,
                      allow_duplicates=False, allow_empty_leaves=False)
    assert [v for v, _, _ in d] == _vals(tm, exp)  # complete list: "all done!"
    assert [x for _, x, _ in d] == dep


# ---- 03xx normalisation and distances (T:559-807) ------------------------------------------------
def test_normalize(tm):
    assert tm.normalize("a") == [0] and tm.normalize("b") == [1]
    assert tm.normalize("aé") == [0, 28]  # unknown symbol = alphabet.len()+1 (src/anahash.rs:76)
    assert tm.anahash("é") == 107  # ... but hashes with prime index alphabet.len() (src/anahash.rs:42)


def n(tm, s):
    return tm.normalize(s)


def test_damerau_levenshtein(tm):
    dl = lambda a, b: orc.damerau_levenshtein(n(tm, a), n(tm, b), 99)
    assert dl("a", "a") == 0 and dl("a", "b") == 1 and dl("ab", "ac") == 1
    assert dl("a", "ab") == 1 and dl("ab", "a") == 1
    assert dl("ab", "ba") == 1  # transposition
    assert dl("abc", "xyz") == 3
    assert dl("hipotesis", "hypothesis") == 2
    assert dl("ca", "abc") == 2  # true (unrestricted) Damerau-Levenshtein, OSA would give 3
    assert orc.damerau_levenshtein(n(tm, "abc"), n(tm, "xyz"), 2) is None
    assert orc.damerau_levenshtein(n(tm, "abcdef"), n(tm, "ab"), 3) is None  # length pre-check
    assert orc.damerau_levenshtein([], n(tm, "ab"), 3) == 2


def test_lcs_prefix_suffix(tm):
    assert orc.lcs(n(tm, "test"), n(tm, "testable")) == 4
    assert orc.lcs(n(tm, "fasttest"), n(tm, "testable")) == 4
    assert orc.lcs(n(tm, "abcdefhij"), n(tm, "def")) == 3 == orc.lcs(n(tm, "def"), n(tm, "abcdefhij"))
    assert orc.prefix(n(tm, "test"), n(tm, "testable")) == 4 == orc.prefix(n(tm, "testable"), n(tm, "test"))
    assert orc.prefix(n(tm, "fasttest"), n(tm, "testable")) == 0 == orc.prefix(n(tm, "fasttest"), n(tm, "test"))
    assert orc.suffix(n(tm, "test"), n(tm, "testable")) == 0 == orc.suffix(n(tm, "testable"), n(tm, "test"))
    assert orc.suffix(n(tm, "fasttest"), n(tm, "testable")) == 0
    assert orc.suffix(n(tm, "fasttest"), n(tm, "test")) == 4


def test_thresholds():
    # src/lib.rs:982-1012
    assert orc.threshold(3, 8) == 3 and orc.threshold(3, 5) == 2 and orc.threshold(3, 1) == 0
    assert orc.threshold(0.5, 9) == 4 and orc.threshold(0.9, 40) == 12
    assert orc.threshold((0.5, 2), 9) == 2 and orc.threshold((0.2, 7), 9) == 1


# ---- 04xx model (T:810-911) ------------------------------------------------------------------------
TEST_PARAMS = dict(max_anagram_distance=2, max_edit_distance=2, max_matches=10, score_threshold=0.0,
                   cutoff_threshold=0.0, freq_weight=0.0, max_ngram=2)  # get_test_searchparams(), src/test.rs:48-68


def small_model(words, confusables=()):
    m = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in words:
        m.add_to_vocabulary(w)
    for pat, wt in confusables:
        m.add_to_confusables(pat, wt)
    m.build()
    return m


def test_model_build_and_anagram_order():
    lex = ["rites", "tiers", "tires", "tries", "tyres", "rides", "brides", "dire"]
    m = small_model(lex)
    for w in lex:
        assert m.has(w)
    assert not m.has("unknown")
    # instances of one anagram keep insertion order (T:836-855); ids start at 3 (src/vocab.rs:145-181)
    assert [m.vocab_lookup(w) for w in lex] == list(range(3, 11))
    r = m.find_variants("rite", orc.make_params(**TEST_PARAMS))  # T:858-869 only requires that this runs
    assert [(m.vocab_text(v), d) for v, d, _ in r] == [("rites", 0.75), ("dire", 0.4375)]


def test_score_tie_order():
    m = small_model(["huis", "huls"])
    r = m.find_variants("huys", orc.make_params(**TEST_PARAMS))
    assert [m.vocab_text(v) for v, _, _ in r] == ["huis", "huls"]  # deterministic tie order (T:872-911)
    assert r[0][1] == r[1][1] and r[0][2] == r[1][2]


# ---- 05xx confusables (T:914-1020) --------------------------------------------------------------------
def test_confusable_found_in():
    assert orc.edit_script("huys", "huis") == "=[hu]-[y]+[i]=[s]"
    assert orc.confusable_found_in("-[y]+[i]", "huys", "huis")
    assert not orc.confusable_found_in("-[y]+[i]", "huys", "huls")


@pytest.mark.parametrize("q", ["huys", "Huys"])
def test_confusable_rescoring(q):
    m = small_model(["huis", "huls"], [("-[y]+[i]", 1.1)])
    r = m.find_variants(q, orc.make_params(**TEST_PARAMS))
    assert [m.vocab_text(v) for v, _, _ in r] == ["huis", "huls"]
    assert r[0][1] > r[1][1]


def test_confusable_nomatch():
    m = small_model(["huis", "huls"], [("-[y]+[p]", 1.1)])
    r = m.find_variants("Huys", orc.make_params(**TEST_PARAMS))
    assert len(r) == 2 and r[0][1] == r[1][1]


# ---- 06xx boundaries and n-grams (T:1023-1117) ----------------------------------------------------------
def test_find_boundaries():
    b = orc.find_boundaries('Hallo allemaal, ik zeg: "Welkom in Aix-les-bains!".')
    assert len(b) == 9
    assert b[0][:3] == (5, 6, " ")
    assert [x[2] for x in b] == [" ", ", ", " ", ': "', " ", " ", "-", "-", '!".']
    HARD, NORMAL, WEAK = 3, 2, 1
    assert [x[3] for x in b] == [NORMAL, HARD, NORMAL, HARD, NORMAL, NORMAL, WEAK, WEAK, HARD]


def test_find_ngrams():
    assert orc.find_match_ngrams("dit is een mooie test", 1) == ["dit", "is", "een", "mooie", "test"]
    assert orc.find_match_ngrams("dit is een mooie test.", 1) == ["dit", "is", "een", "mooie", "test"]
    assert orc.find_match_ngrams("hello, world!", 1) == ["hello", "world"]
    assert orc.find_match_ngrams("dit is een mooie test.", 2) == ["dit is", "is een", "een mooie", "mooie test"]
    assert orc.find_match_ngrams("hello,world!", 2) == ["hello,world"]
    assert orc.find_match_ngrams("hello, world!", 2) == ["hello, world"]
    assert orc.find_match_ngrams("hello!", 2) == []


# ---- 07xx / 09xx find_all_matches, unigram path (T:1120-1140, 1432-1481, 1513-1572) ----------------------
def test_find_all_matches_unigram():
    m = small_model(["I", "think", "sink", "you", "are", "right"])
    p = orc.make_params(**{**TEST_PARAMS, "max_ngram": 1})
    segs = m.find_all_segments("I tink you are rihgt", p)
    assert [s["text"] for s in segs] == ["I", "tink", "you", "are", "rihgt"]
    best = [m.vocab_text(s["variants"][0][0]) for s in segs]
    assert best[1] == "think" and best[4] == "right"


def test_find_all_matches_utf8_offsets():
    m = small_model(["I", "think", "you", "are", "right"])
    p = orc.make_params(**{**TEST_PARAMS, "max_ngram": 1})
    segs = m.find_all_segments("I thиnk you are rihgt", p)
    assert segs[1]["text"] == "thиnk" and (segs[1]["begin"], segs[1]["end"]) == (2, 8)
    assert m.vocab_text(segs[1]["variants"][0][0]) == "think"
    assert m.vocab_text(segs[4]["variants"][0][0]) == "right"


def test_multiple_lexicons():
    m = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    m.read_lexicon(workloads.AMPHIBIANS)
    m.read_lexicon(workloads.REPTILES)
    m.build()
    p = orc.make_params(**{**TEST_PARAMS, "max_ngram": 1})
    segs = m.find_all_segments("Salamander lizard frog snake toad", p)
    assert [s["text"] for s in segs] == ["Salamander", "lizard", "frog", "snake", "toad"]
    best = [s["variants"][0][0] for s in segs]
    assert [m.vocab_text(v) for v in best] == ["salamander", "lizard", "frog", "snake", "toad"]
    assert [m.vocab_lexindex(v) for v in best] == [1, 2, 1, 2, 1]


# ---- large-lexicon goldens (README.md:107-109,222-241; tutorial.ipynb) -------------------------------------
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tutorial.json"), encoding="utf-8"))


def test_index_goldens(eng_oracle):
    m = eng_oracle
    assert m.instance_count() == 119773 and m.index_size() == 108802
    hist = [27, 248, 942, 2593, 5623, 10163, 14617, 16911, 16391, 13930, 10650, 7194, 4434, 2459, 1384, 667, 339, 128,
            62, 20, 9, 8, 2, 1]
    assert [m.sortedindex_count(i + 1) for i in range(24)] == hist
    for w in ("least", "slate", "Stael", "stale", "steal", "tales", "teals", "Tesla"):
        assert m.anahash(w) == 1227306
    for w in ("ales", "Elsa"):
        assert m.anahash(w) == 17286
    assert m.max_key_bits() == 102


def _check(m, got, gold_variants):
    assert [m.vocab_text(v) for v, _, _ in got] == [g["text"] for g in gold_variants]
    for (v, d, f), g in zip(got, gold_variants):
        assert d == g["dist_score"] and f == g["freq_score"]  # bit-exact f64 (repr round trip)


@pytest.mark.parametrize("q", ["separate", "seperate"])
def test_tutorial_find_variants(eng_oracle, q):
    _check(eng_oracle, eng_oracle.find_variants(q, orc.make_params()), GOLD["find_variants"][q])


def test_tutorial_find_all_matches(eng_oracle):
    # per-token variant lists of find_all_matches("We would like seperate beds") -- each is the
    # find_variants list of that unigram (selected variant is index 0 in all five).
    for match in GOLD["find_all_matches"]["We would like seperate beds"]:
        _check(eng_oracle, eng_oracle.find_variants(match["input"], orc.make_params()), match["variants"])
    g = GOLD["find_all_matches"]["We would like sep arate beds"]["match"]
    _check(eng_oracle, eng_oracle.find_variants(g["input"], orc.make_params()), g["variants"])  # 1-ulp cases
    segs = eng_oracle.find_all_segments("We would like seperate beds", orc.make_params())
    uni = [s for s in segs if s["n"] == 1]
    assert [(s["text"], s["begin"], s["end"]) for s in uni] == [
        (mm["input"], mm["offset"]["begin"], mm["offset"]["end"]) for mm in GOLD["find_all_matches"]["We would like seperate beds"]]


# ---- 0801 variant lists (T:1484-1510; src/lib.rs:460-514, 766-897, 1677-1727) ------------------------
def test_expand_variants_transparent():
    m = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    ref = m.add_to_vocabulary("afgescheid")
    assert m.add_variant(ref, "afghescheydt", 1.0, vocab_type="TRANSPARENT")
    m.build()
    # very strict parameters: the reference form cannot match, the transparent variant can (T:1497-1499)
    r = m.find_variants("afgheschaydt", orc.make_params(**dict(TEST_PARAMS, max_anagram_distance=2, max_edit_distance=2)),
                        with_via=True)
    assert len(r) == 1
    assert m.vocab_text(r[0][0]) == "afgescheid"
    assert m.vocab_text(r[0][3]) == "afghescheydt"  # via = the variant that matched


def test_expand_variants_semantics(tmp_path):
    """Non-transparent variants stay in the list next to their reference; the expanded score is dist * variant score;
    the frequency score is min(reference frequency, the variant's own) before normalisation; consecutive duplicates
    of one vocabulary id are dropped after ranking (src/lib.rs:1510-1533)."""
    f = tmp_path / "variants.tsv"
    f.write_text("separate\tseperate\t0.9\tseparete\t0.8\nhouse\thuose\t0.5\n")
    m = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    m.add_to_vocabulary("separated")
    m.read_variants(str(f), transparent=False)
    m.build()
    p = orc.make_params(**dict(TEST_PARAMS, max_anagram_distance=3, max_edit_distance=3))
    r = m.find_variants("seperate", p, with_via=True)
    got = [(m.vocab_text(v), round(d, 6), None if via is None else m.vocab_text(via)) for v, d, _, via in r]
    # "seperate" matches itself exactly (1.0) -> expands to separate at 1.0 * 0.9 via seperate; the direct match of
    # "separate" scores lower than 0.9, is ranked behind the expanded one and is NOT adjacent to it, so it stays
    names = [g[0] for g in got]
    assert got[0] == ("seperate", 1.0, None)
    assert ("separate", 0.9, "seperate") in got
    assert names.count("separate") >= 1 and "separete" in names
    # transparent: the variants themselves disappear, only references remain
    m2 = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    m2.add_to_vocabulary("separated")
    m2.read_variants(str(f), transparent=True)
    m2.build()
    r2 = m2.find_variants("seperate", p, with_via=True)
    names2 = [m2.vocab_text(v) for v, _, _, _ in r2]
    assert "seperate" not in names2 and "separete" not in names2 and "separate" in names2
    assert m2.vocab_type(m2.vocab_lookup("seperate")) & 4 and not (m2.vocab_type(m2.vocab_lookup("separate")) & 4)


def test_read_variants_with_frequencies(tmp_path):
    f = tmp_path / "variants_freq.tsv"
    f.write_text("separate\t100\tseperate\t0.9\t7\nhouse\t50\thuose\t0.5\t2\n")
    m = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    m.read_variants(str(f))
    assert m.vocab_freq(m.vocab_lookup("separate")) == 100 and m.vocab_freq(m.vocab_lookup("seperate")) == 7
    assert m.vocab_freq(m.vocab_lookup("huose")) == 2
