"""CPU-only: the oracle's restatement of the language-model and context-rule terms of most_likely_sequence
(src/lib.rs:2088-2674, src/search.rs:338-524) against the reference's own tests: 0702-0705 (tests/main.rs:1143-1429,
with their LM entries) and 0902-0905 (tests/main.rs:1575-1728: bonus, penalty, tags, sequence numbers, multiple tags)."""
import pytest

from oracle import orc

TEST_PARAMS = dict(max_anagram_distance=2, max_edit_distance=2, max_matches=10, score_threshold=0.0,
                   cutoff_threshold=0.0, freq_weight=0.0, max_ngram=2)  # get_test_searchparams(), src/test.rs:48-68
LM_ENTRIES = [("<bos> I", 2), ("I think", 2), ("I sink", 1), ("you are", 2), ("right <eos>", 2)]  # T:1154-1193
WORDS = ["I", "think", "sink", "you", "are", "right"]


def model(words, lm=()):
    m = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in words:
        m.add_to_vocabulary(w, 2)
    for w, f in lm:
        m.add_to_vocabulary(w, f, vocab_type="LM")
    m.build()
    return m


def rendered(m, matches):
    """(input text, match_to_str) per match: the selected variant's text, else the input (src/lib.rs:2756-2762)."""
    return [(s["text"], m.vocab_text(s["variants"][s["selected"]][0]) if s["selected"] >= 0 else s["text"]) for s in matches]


def test_reference_0702_0703_with_language_model():
    m = model(WORDS + ["are right"], LM_ENTRIES)
    assert m.have_lm() and m.ngram_count() == 5
    p = orc.make_params(**TEST_PARAMS)
    r = m.find_all_matches("I tink you are rihgt", p)
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right")]  # T:1196-1208
    assert (r[1]["begin"], r[1]["end"]) == (2, 6)
    r = m.find_all_matches("I tink you are\nrihgt", p)
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are\nrihgt", "are right")]  # T:1243-1252


def test_reference_0704_two_batches_with_language_model():
    m = model(WORDS + ["am", "sure", "are right"], LM_ENTRIES + [("I am", 2), ("sure <eos>", 2)])  # T:1255-1340
    r = m.find_all_matches("I tink you are rihgt\n\nI am sur", orc.make_params(**TEST_PARAMS))
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right"),
                              ("I", "I"), ("am", "am"), ("sur", "sure")]  # T:1346-1360


def test_reference_0705_lm_weight_zero():
    m = model(WORDS + ["are right"], LM_ENTRIES)
    r = m.find_all_matches("I tink you are rihgt", orc.make_params(**TEST_PARAMS, lm_weight=0.0))  # T:1417-1418
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are rihgt", "are right")]  # T:1420-1428


RULE_PARAMS = dict(TEST_PARAMS, max_ngram=1, lm_weight=0.0)  # T:1590-1592


def test_reference_0902_context_rule_bonus_and_tag():
    m = model(WORDS)
    m.add_contextrule("I; think", 1.1, ["testtag"], [])
    r = m.find_all_matches("I tink you are rihgt", orc.make_params(**RULE_PARAMS))
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are", "are"), ("rihgt", "right")]
    assert (r[0]["tag"], r[0]["seqnr"]) == ([0], [0]) and (r[1]["tag"], r[1]["seqnr"]) == ([0], [1])  # T:1596-1602
    assert all(s["tag"] == [] for s in r[2:])
    assert m.tags() == ["testtag"]


def test_reference_0903_context_rule_penalty():
    m = model(WORDS)
    m.add_contextrule("I; think", 0.9, [], [])
    r = m.find_all_matches("I tink you are rihgt", orc.make_params(**RULE_PARAMS))
    assert rendered(m, r) == [("I", "I"), ("tink", "sink"), ("you", "you"), ("are", "are"), ("rihgt", "right")]  # T:1633-1641


def test_reference_0904_single_word_rules_share_a_tag():
    m = model(WORDS)
    for w in ("think", "are", "right"):
        m.add_contextrule(w, 1.0, ["testtag"], [])
    r = m.find_all_matches("I tink you are rihgt", orc.make_params(**RULE_PARAMS))
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are", "are"), ("rihgt", "right")]
    assert [(s["tag"], s["seqnr"]) for s in r] == [([], []), ([0], [0]), ([], []), ([0], [0]), ([0], [0])]  # T:1671-1687


def test_reference_0905_multiple_tags():
    m = model(WORDS)
    m.add_contextrule("I; think", 1.1, ["testtag", "testtag2"], [])
    r = m.find_all_matches("I tink you are rihgt", orc.make_params(**RULE_PARAMS))
    assert rendered(m, r) == [("I", "I"), ("tink", "think"), ("you", "you"), ("are", "are"), ("rihgt", "right")]
    assert (r[0]["tag"], r[0]["seqnr"]) == ([0, 1], [0, 0]) and (r[1]["tag"], r[1]["seqnr"]) == ([0, 1], [1, 1])  # T:1711-1718
    assert m.tags() == ["testtag", "testtag2"]


def test_lm_score_tokens_restated_by_hand():
    """src/lib.rs:2643-2674: known bigram = ln(joint / prior) in f32 (prior = unigram count, 1 when unseen), unknown
    bigram or an out-of-vocabulary token = the smoothing constant; perplexity = -logprob / transitions."""
    import numpy as np
    m = model(WORDS, LM_ENTRIES + [("we", 4), ("we think", 2)])  # ("we" is LM-only: a unigram count; "I" is INDEXED: none)
    I, we, think, you = (m.vocab_lookup(w) for w in ("I", "we", "think", "you"))
    lp, pp = m.lm_score_tokens([0, I, think, None, you, 1])
    s = np.float32(-13.815510557964274)
    exp = np.float32(0)
    for term in (np.log(np.float32(2.0)),   # <bos> I: prior <bos> unseen -> 1 < joint 2 -> ln(joint)
                 np.log(np.float32(2.0)),   # I think: no unigram count for "I" -> prior 1 < joint 2 -> ln(joint)
                 s, s,                      # think ?, ? you
                 s):                        # you <eos>: unseen bigram
        exp = np.float32(exp + term)
    assert lp == pytest.approx(float(exp), rel=1e-6)
    assert pp == pytest.approx(-float(exp) / 5, rel=1e-6)
    lp, pp = m.lm_score_tokens([0, we, think, 1])
    exp = np.float32(np.float32(s + np.log(np.float32(2.0) / np.float32(4.0))) + s)  # <bos> we unseen; we think 2 / 4; think <eos> unseen
    assert lp == pytest.approx(float(exp), rel=1e-6)
    assert pp == pytest.approx(-float(exp) / 3, rel=1e-6)


def test_exhaustive_and_per_state_enumeration_agree():
    """The oracle ranks the paths of small lattices by exhaustive enumeration and of large ones from per-state lists of
    the best partial paths; both must give the same sequences, also when max_seq cuts the list (ties included)."""
    m = model(WORDS + ["are right", "tin", "thin", "ink", "your", "our"], LM_ENTRIES + [("you are", 3), ("thin <eos>", 1)])
    m.add_contextrule("I; think", 1.1, ["a"], [])
    m.add_contextrule("you | your; ?; right", 0.8, ["b"], ["1:1"])
    texts = ["I tink you are rihgt", "tink I yor are rihgt tin", "I tink", "rihgt", "you are you are you are"]
    for max_seq in (1, 2, 3, 7, 250):
        for text in texts:
            p = orc.make_params(**dict(TEST_PARAMS, max_ngram=3), max_seq=max_seq)
            orc.lib().orc_set_bruteforce_limit(200000)
            a = m.find_all_matches(text, p)
            orc.lib().orc_set_bruteforce_limit(0)
            b = m.find_all_matches(text, p)
            orc.lib().orc_set_bruteforce_limit(200000)
            assert a == b, (text, max_seq)
