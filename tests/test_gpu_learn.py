"""GPU: learn_variants (src/lib.rs:1062-1139) end to end -- batched GPU lookups, the host bookkeeping, the rebuild --
against the oracle: the returned count, the vocabulary afterwards (texts, frequencies, types, variant links) and the
lookups on the rebuilt model (learned links show up as `via`).  Strict mode (one batched find_variants call over all
inputs) and running-text mode (find_all_matches per input)."""
import ctypes as C

import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu


def pair(words):
    import analiticcl_b200 as A
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    for w, f in words:
        assert o.add_to_vocabulary(w, f) == m.add_to_vocabulary(w, f, A.VocabParams())
    o.build()
    m.build()
    return o, m


def state(x, n):
    if isinstance(x, orc.OracleModel):
        return [(x.vocab_text(i), x.vocab_freq(i), x.vocab_type(i), x.vocab_links(i)) for i in range(n)]
    out = []
    for i in range(n):
        info = x._vocab(i)
        out.append((C.string_at(info.text, info.text_len).decode("utf-8"), info.frequency, info.vocabtype, x.vocab_links(i)))
    return out


WORDS = [(w, f) for f, w in enumerate(["house", "mouse", "horse", "hose", "houses", "tree", "three", "there", "their", "separate",
                                        "desperate", "operate", "the", "then", "than", "that"], start=2)]


@pytest.mark.parametrize("strict", [True, False])
def test_learn_variants_equals_oracle(strict):
    import analiticcl_b200 as A
    o, m = pair(WORDS)
    if strict:
        inputs = ["huose", "hause", "mouse", "house", "seperate", "tre", "thre", "zzzzzzzz", "huose", "thn", "teh"]
    else:
        inputs = ["teh huose and teh mouse", "a seperate tre, thre of them", "zzzzzzzz", "then than that thn"]
    kw = dict(max_anagram_distance=2, max_edit_distance=2, max_ngram=2 if not strict else 3)
    sp, op = A.SearchParameters(**kw), orc.make_params(**kw)
    n_o = o.learn_variants(inputs, op, strict=strict, auto_build=True)
    n_m = m.learn_variants(inputs, sp, strict=strict, auto_build=True)
    assert n_m == n_o > 5
    n = o.vocab_size()
    assert n == m._vocab_size() and state(m, n) == state(o, n)
    # lookups on the rebuilt models: the learned links expand results (via) exactly as in the oracle
    queries = ["huose", "hause", "seperate", "thre", "teh", "hose", "treee"]
    got = m.find_variants_raw(queries, sp, with_via=True)
    for q, g in zip(queries, got):
        assert g == o.find_variants(q, op, with_via=True), q
    # a second round learns from the enlarged model (the CLI iterates, src/bin/analiticcl.rs:500-545)
    assert m.learn_variants(inputs, sp, strict=strict, auto_build=True) == o.learn_variants(inputs, op, strict=strict, auto_build=True)
    n = o.vocab_size()
    assert state(m, n) == state(o, n)


def test_learn_without_rebuild_invalidates_the_index():
    import analiticcl_b200 as A
    o, m = pair(WORDS)
    sp = A.SearchParameters(max_anagram_distance=2, max_edit_distance=2)
    assert m.learn_variants(["huose"], sp, strict=True, auto_build=False) > 0
    with pytest.raises(RuntimeError, match="not been built"):
        m.find_variants("huose", sp)
    m.build()
    assert m.find_variants("huose", sp)
