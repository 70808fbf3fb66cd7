"""CPU-only: the product's host code for the language-model and context-rule terms of the sequence consolidation
(csrc/sequence.cpp, anl_model_consolidate) against the oracle: the reference's tests 0702-0705 / 0902-0905 through the
C ABI, lm_score_tokens bit for bit, and random lattices with random n-gram tables and rule sets (tags, offsets, negation,
disjunction, any, no-lexicon, @lexicon; max_seq cutting the path list; cost ties)."""
import ctypes as C

import numpy as np
import pytest

from oracle import orc

TEST_PARAMS = dict(max_anagram_distance=2, max_edit_distance=2, max_matches=10, score_threshold=0.0,
                   cutoff_threshold=0.0, freq_weight=0.0, max_ngram=2)  # get_test_searchparams(), src/test.rs:48-68
LM_ENTRIES = [("<bos> I", 2), ("I think", 2), ("I sink", 1), ("you are", 2), ("right <eos>", 2)]
WORDS = ["I", "think", "sink", "you", "are", "right"]


@pytest.fixture(scope="module")
def L():
    from analiticcl_b200 import build, _capi
    build.build()
    return _capi.lib()


def both(words, lm=(), rules=(), lexicons=None):
    """The same model twice: oracle and product (host side only: build() uploads to a GPU when there is one)."""
    import analiticcl_b200 as A
    o = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    m = A.VariantModel(None, A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in words:
        idx = (lexicons or {}).get(w, 0)
        assert o.add_to_vocabulary(w, 2, index=idx) == m.add_to_vocabulary(w, 2, A.VocabParams(index=idx))
    for w, f in lm:
        assert o.add_to_vocabulary(w, f, vocab_type="LM") == m.add_to_vocabulary(w, f, A.VocabParams(vocabtype="LM"))
    o.build()
    try:
        m.build()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    for pattern, score, tag, tagoffset in rules:
        o.add_contextrule(pattern, score, tag, tagoffset)
        m.add_contextrule(pattern, score, tag, tagoffset)
    return o, m


def product_sequence(L, m, text, sp, segments):
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
    assert L.anl_debug_match_set_build(raw, len(raw), sp.data.max_ngram, 0, looked, offs, vs, n, C.byref(ms)) == 0, L.anl_last_error()
    try:
        assert L.anl_model_consolidate(m._h, ms, raw, len(raw), C.byref(sp.data), C.byref(out)) == 0, L.anl_last_error()
    finally:
        L.anl_match_set_free(ms)
    got = []
    mm = _capi.Match()
    tg, sq = C.POINTER(C.c_uint16)(), C.POINTER(C.c_uint8)()
    for i in range(L.anl_match_set_len(out)):
        assert L.anl_match_set_get(out, i, C.byref(mm)) == 0
        nt = L.anl_match_set_tags(out, i, C.byref(tg), C.byref(sq))
        got.append({"begin": int(mm.begin), "end": int(mm.end), "n": int(mm.n), "selected": int(mm.selected),
                    "variants": [(mm.variants[j].vocab_id, mm.variants[j].dist_score, mm.variants[j].freq_score)
                                 for j in range(mm.n_variants)] if mm.variants else [],
                    "tag": [tg[k] for k in range(nt)], "seqnr": [sq[k] for k in range(nt)]})
    L.anl_match_set_free(out)
    return got


def strip(matches):
    return [{k: s[k] for k in ("begin", "end", "n", "selected", "variants", "tag", "seqnr")} for s in matches]


def sp_and_op(**kw):
    import analiticcl_b200 as A
    sp = A.SearchParameters(**kw)
    okw = {k: v for k, v in kw.items() if k != "consolidate_matches"}
    return sp, orc.make_params(**okw)


CASES = [
    # (words, lm, rules, text, extra parameters)  -- the reference's tests, tests/main.rs
    (WORDS + ["are right"], LM_ENTRIES, [], "I tink you are rihgt", {}),                                            # 0702
    (WORDS + ["are right"], LM_ENTRIES, [], "I tink you are\nrihgt", {}),                                           # 0703
    (WORDS + ["am", "sure", "are right"], LM_ENTRIES + [("I am", 2), ("sure <eos>", 2)], [], "I tink you are rihgt\n\nI am sur", {}),  # 0704
    (WORDS + ["are right"], LM_ENTRIES, [], "I tink you are rihgt", {"lm_weight": 0.0}),                            # 0705
    (WORDS, [], [("I; think", 1.1, ["testtag"], [])], "I tink you are rihgt", {"lm_weight": 0.0, "max_ngram": 1}),  # 0902
    (WORDS, [], [("I; think", 0.9, [], [])], "I tink you are rihgt", {"lm_weight": 0.0, "max_ngram": 1}),           # 0903
    (WORDS, [], [(w, 1.0, ["testtag"], []) for w in ("think", "are", "right")], "I tink you are rihgt", {"lm_weight": 0.0, "max_ngram": 1}),  # 0904
    (WORDS, [], [("I; think", 1.1, ["testtag", "testtag2"], [])], "I tink you are rihgt", {"lm_weight": 0.0, "max_ngram": 1}),  # 0905
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_reference_tests_through_the_c_abi(L, case):
    words, lm, rules, text, extra = CASES[case]
    o, m = both(words, lm, rules)
    assert m.have_lm() == o.have_lm() == bool(lm)
    assert m.tags() == o.tags()
    sp, op = sp_and_op(**{**TEST_PARAMS, **extra})
    segments = o.find_all_segments(text, op)
    exp = o.find_all_matches(text, op)
    assert strip(o.find_all_matches(text, op, segments)) == strip(exp)  # the oracle's hook path = its lookup path
    assert product_sequence(L, m, text, sp, segments) == strip(exp)
    rendered = [o.vocab_text(s["variants"][s["selected"]][0]) if s["selected"] >= 0 else s["text"] for s in exp]
    assert rendered[1] == ("sink" if case == 5 else "think")  # the penalty of test 0903 flips the choice


def test_lm_score_tokens_bit_for_bit(L):
    o, m = both(WORDS, LM_ENTRIES + [("we", 4), ("we think", 2), ("you", 7), ("a b c", 3), ("zzz you", 5)])
    assert L.anl_model_ngram_count(m._h) == o.ngram_count()
    rng = np.random.default_rng(3)
    n_vocab = o.vocab_size()
    for _ in range(300):
        toks = [0] + [None if rng.random() < 0.15 else int(rng.integers(0, n_vocab)) for _ in range(int(rng.integers(0, 9)))] + [1]
        arr = (C.c_int64 * len(toks))(*[-1 if t is None else t for t in toks])
        lp, pp = C.c_float(), C.c_double()
        L.anl_debug_lm_score_tokens(m._h, arr, len(toks), C.byref(lp), C.byref(pp))
        assert (lp.value, pp.value) == o.lm_score_tokens(toks)


def test_contextrule_errors_like_the_reference(L):
    o, m = both(WORDS)
    for bad in ("nosuchword", "I; @nolexicon", "I | nosuchword"):
        with pytest.raises(RuntimeError, match="Error parsing context rule"):
            m.add_contextrule(bad, 1.1)
        with pytest.raises(RuntimeError):
            o.add_contextrule(bad, 1.1)
    with pytest.raises(RuntimeError, match="tag offset should be an integer"):
        m.add_contextrule("I", 1.1, ["t"], ["x:1"])
    assert L.anl_model_contextrule_count(m._h) == 0


def test_read_contextrules_file(L, tmp_path):
    o, m = both(WORDS)
    f = tmp_path / "rules.tsv"
    f.write_text("# comment\n\nI; think\t1.1\tsubj ; verb\t0:1;1:1\nyou | I; ?; right\t0.8\tclause\n^; are\t1.05\n!think; !(are|right)\t0.95\tneg\t1:\n",
                 encoding="utf-8")
    o.read_contextrules(str(f))
    m.read_contextrules(str(f))
    assert L.anl_model_contextrule_count(m._h) == 4
    assert m.tags() == o.tags() == ["subj", "verb", "clause", "neg"]
    sp, op = sp_and_op(**{**TEST_PARAMS, "lm_weight": 0.0})
    for text in ("I tink you are rihgt", "zzzzzzzzzz are rihgt you sink", "think think I I"):
        segments = o.find_all_segments(text, op)
        assert product_sequence(L, m, text, sp, segments) == strip(o.find_all_matches(text, op, segments)), text
    bad = tmp_path / "bad.tsv"
    bad.write_text("I; think\n", encoding="utf-8")
    with pytest.raises(RuntimeError, match="at least two columns"):
        m.read_contextrules(str(bad))


def test_random_lattices_language_models_and_rules(L):
    """Random variant lists (coarse score grid: equal-cost paths are common), random bigram tables, random rules."""
    rng = np.random.default_rng(2024)
    words = ["aa", "b", "ccc", "dd", "e", "ff", "gg", "hh", "aa b", "ccc dd", "e ff gg", ","]
    lexicons = {w: int(rng.integers(0, 3)) for w in words}
    seps = [" ", " ", " ", ", ", ". ", "\n", "-", "  ", "; "]
    atoms = ["?", "^", "@lexA", "!aa", "b|ccc", "!(dd|e)", "aa", "ccc dd", "ff", ","]
    checked = tagged = 0
    for case in range(40):
        lm = []
        for _ in range(int(rng.integers(0, 12))):
            k = int(rng.integers(1, 4))
            toks = [str(rng.choice(["<bos>", "<eos>", "aa", "b", "ccc", "dd", "e", "ff", ",", "qq"])) for _ in range(k)]
            lm.append((" ".join(toks), int(rng.integers(1, 9))))
        rules = []
        for _ in range(int(rng.integers(0, 5))):
            plen = int(rng.integers(1, 4))
            pat = "; ".join(str(rng.choice(atoms)) for _ in range(plen))
            ntag = int(rng.integers(0, 3))
            tag = [f"t{int(rng.integers(0, 4))}" for _ in range(ntag)]
            off = [f"{int(rng.integers(0, plen))}:{int(rng.integers(1, 3))}" for _ in range(ntag)] if rng.random() < 0.5 else []
            rules.append((pat, float(rng.choice([0.8, 0.9, 1.0, 1.1, 1.25])), tag, off))
        import analiticcl_b200 as A
        o = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
        m = A.VariantModel(None, A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
        # three "lexicons" so that @lexA and the no-lexicon test have something to see
        o.lexicon_names = None
        for w in words:
            assert o.add_to_vocabulary(w, 2, index=lexicons[w]) == m.add_to_vocabulary(w, 2, A.VocabParams(index=lexicons[w]))
        for w, f in lm:
            assert o.add_to_vocabulary(w, f, vocab_type="LM") == m.add_to_vocabulary(w, f, A.VocabParams(vocabtype="LM"))
        o.build()
        try:
            m.build()
        except RuntimeError as e:
            assert "no CPU fallback" in str(e)
        for pat, score, tag, off in rules:
            if "@lexA" in pat:
                continue  # (no lexicon file was read: the name cannot resolve -- covered by the error test)
            o.add_contextrule(pat, score, tag, off)
            m.add_contextrule(pat, score, tag, off)
        ntok = int(rng.integers(1, 9))
        toks = [str(rng.choice(["aa", "b", "ccc", "dd", "e", "ff", "xq"])) for _ in range(ntok)]
        text = "".join(t + seps[int(rng.integers(0, len(seps)))] for t in toks)
        for max_seq in (1, 4, 250):
            kw = dict(max_ngram=int(rng.integers(1, 4)), freq_weight=float(rng.choice([0.0, 0.5])), max_seq=max_seq,
                      lm_weight=float(rng.choice([0.0, 1.0, 2.0])), variantmodel_weight=float(rng.choice([1.0, 3.0])),
                      contextrules_weight=float(rng.choice([0.0, 1.0])))
            sp, op = sp_and_op(**kw)
            raw = text.encode("utf-8")
            cap = kw["max_ngram"] * (len(raw) + 2)
            b, e = (C.c_uint64 * cap)(), (C.c_uint64 * cap)()
            od, bt = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
            nseg = L.anl_debug_segment_text(raw, len(raw), kw["max_ngram"], b, e, od, bt, cap)
            segments = []
            for k in range(nseg):
                looked = od[k] == 1 or rng.random() < 0.7
                nv = int(rng.integers(0, 4)) if looked else 0
                vs = sorted(((int(rng.integers(3, 3 + len(words))), float(rng.integers(0, 5)) / 4.0, float(rng.integers(0, 3)) / 2.0)
                             for _ in range(nv)), key=lambda v: -v[1])
                segments.append({"looked_up": bool(looked), "variants": vs})
            exp = strip(o.find_all_matches(text, op, segments))
            got = product_sequence(L, m, text, sp, segments)
            assert got == exp, (case, max_seq, text, lm, rules, kw)
            checked += 1
            tagged += any(s["tag"] for s in exp)
    assert checked == 120 and tagged > 5


def test_lexicon_patterns_and_language_model_from_files(L, tmp_path):
    """@lexicon resolves by file name or by path suffix (src/search.rs:452-461); read_lm = read_vocabulary with the LM
    type (bindings/python/src/lib.rs:659-667)."""
    import analiticcl_b200 as A
    la, lb, lmf = tmp_path / "first.tsv", tmp_path / "second.tsv", tmp_path / "lm.tsv"
    la.write_text("I\t5\nthink\t3\nyou\t4\n", encoding="utf-8")
    lb.write_text("sink\t3\nare\t6\nright\t2\nyou\t1\n", encoding="utf-8")
    lmf.write_text("<bos> I\t4\nI think\t1\nI sink\t3\nI\t4\nyou are\t2\nright <eos>\t2\n", encoding="utf-8")
    o = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    m = A.VariantModel(None, A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for f in (la, lb):
        o.read_lexicon(str(f))
        m.read_lexicon(str(f))
    o.read_lm(str(lmf))
    m.read_lm(str(lmf))
    o.build()
    try:
        m.build()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    assert m.have_lm() and o.have_lm() and L.anl_model_ngram_count(m._h) == o.ngram_count() == 5  # ("I" is already an INDEXED entry: it keeps that type and gives no unigram count)
    for mm in (o, m):
        mm.add_contextrule("@first.tsv; @second.tsv", 1.2, ["ab"], [])
        mm.add_contextrule("@" + str(lb) + "; ^", 0.7, [], [])
    with pytest.raises(RuntimeError, match="was not loaded"):
        m.add_contextrule("@third.tsv", 1.1)
    for kw in ({}, {"lm_weight": 0.0}, {"contextrules_weight": 0.0}, {"lm_weight": 3.0, "variantmodel_weight": 0.5}):
        sp, op = sp_and_op(**{**TEST_PARAMS, **kw})
        for text in ("I tink you are rihgt", "I tink you are zzzzzzzzzz rihgt", "sink you"):
            segments = o.find_all_segments(text, op)
            exp = strip(o.find_all_matches(text, op, segments))
            assert product_sequence(L, m, text, sp, segments) == exp, (kw, text)
    # the language model prefers "I sink" (count 3) over "I think" (count 1) when it outweighs the variant model
    sp, op = sp_and_op(**{**TEST_PARAMS, "lm_weight": 3.0, "variantmodel_weight": 0.5, "contextrules_weight": 0.0})
    r = o.find_all_matches("I tink", op)
    assert o.vocab_text(r[1]["variants"][r[1]["selected"]][0]) == "sink"
    sp, op = sp_and_op(**{**TEST_PARAMS, "lm_weight": 0.0, "contextrules_weight": 0.0})
    r = o.find_all_matches("I tink", op)
    assert o.vocab_text(r[1]["variants"][r[1]["selected"]][0]) == "think"


def test_pattern_parser_accepts_and_rejects_like_the_oracle(L):
    """Random pattern strings (valid and malformed: stray operators, unknown words, empty positions, nested negation):
    product and oracle must agree on accept / reject, and on what an accepted rule matches."""
    rng = np.random.default_rng(99)
    o, m = both(WORDS)
    atoms = ["I", "think", "sink", "you", "?", "^", "nosuch", "", " ", "!", "|", "!(", ")", "@x", "are", "!(I|you)", "!think", "I|are"]
    accepted = rejected = 0
    for _ in range(400):
        n = int(rng.integers(1, 4))
        pat = ";".join("".join(str(rng.choice(atoms)) for _ in range(int(rng.integers(1, 3)))) for _ in range(n))
        tag = ["t"] if rng.random() < 0.5 else []
        off = [str(rng.choice(["0:", "1:1", ":", "0:2", "x", "1"]))] if tag and rng.random() < 0.5 else []
        ok_o = ok_m = True
        try:
            o.add_contextrule(pat, 1.1, tag, off)
        except RuntimeError:
            ok_o = False
        try:
            m.add_contextrule(pat, 1.1, tag, off)
        except (RuntimeError, ValueError):
            ok_m = False
        assert ok_o == ok_m, (pat, tag, off)
        accepted += ok_o
        rejected += not ok_o
    assert accepted > 30 and rejected > 30
    assert L.anl_model_contextrule_count(m._h) == orc.lib().orc_contextrule_count(o.h)
    assert m.tags() == o.tags()
    sp, op = sp_and_op(**{**TEST_PARAMS, "lm_weight": 0.0, "max_ngram": 1})
    for text in ("I tink you are rihgt", "you sink I think", "zzzzzzzzzz I are"):
        segments = o.find_all_segments(text, op)
        assert product_sequence(L, m, text, sp, segments) == strip(o.find_all_matches(text, op, segments)), text
