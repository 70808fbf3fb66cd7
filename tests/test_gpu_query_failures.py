"""Failures are per query, never per batch (the reference answers every query on its own, src/lib.rs:972-1027):
a query outside the limits of the GPU path comes back with an empty list and ANL_QUERY_UNSUPPORTED in its flags;
every other query of the batch -- and of a find_all_matches text -- is answered as usual."""
import ctypes as C

import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu

QUERY_EMPTY, QUERY_UNSUPPORTED = 1, 2


def run_with_flags(m, A, qs, sp):
    from analiticcl_b200 import _capi
    L = _capi.lib()
    blob, offs = _capi.pack(qs)
    rs = C.c_void_p()
    st = L.anl_find_variants_batch(m._h, blob, _capi.u64ptr(offs), len(qs), C.byref(sp.data), C.byref(rs))
    assert st == 0, L.anl_last_error()
    try:
        o = L.anl_result_set_offsets(rs)
        v = L.anl_result_set_variants(rs)
        lists = [[(v[j].vocab_id, v[j].dist_score, v[j].freq_score) for j in range(o[i], o[i + 1])] for i in range(len(qs))]
        flags = [L.anl_result_set_flags(rs, i) for i in range(len(qs))]
        return lists, flags
    finally:
        L.anl_result_set_free(rs)


def test_over_budget_queries_inside_a_10k_batch(eng_oracle):
    """max_anagram_distance = Ratio(0.5) thresholds to min(floor(len / 2), 12): beyond the 6 the device enumerates for
    queries of 14 symbols and more.  Those queries are flagged; the others equal the oracle."""
    import analiticcl_b200 as A
    from test_gpu_parity import assert_same, to_orc_params
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build()
    qs = workloads.misspellings(workloads.read_words("eng"), 10_000, 4242, min_len=3, max_len=13)
    qs = [q[:13] for q in qs]
    heavy = {17: "counterrevolutionaries", 5000: "x" * 300, 9999: "antidisestablishmentarianism", 9000: ""}
    for i, q in heavy.items():
        qs[i] = q
    sp = A.SearchParameters(max_anagram_distance=0.5, max_edit_distance=2)
    got, flags = run_with_flags(m, A, qs, sp)
    for i, q in heavy.items():
        assert got[i] == []
    assert flags[17] == QUERY_UNSUPPORTED and flags[9999] == QUERY_UNSUPPORTED
    assert flags[9000] == QUERY_EMPTY
    assert flags[5000] == 0  # 300 symbols: longer than any entry + distance -> empty by construction, not a limit
    ok = [i for i in range(len(qs)) if i not in heavy]
    assert all(flags[i] == 0 for i in ok)
    sample = ok[::7]
    exp = eng_oracle.find_variants_batch([qs[i] for i in sample], to_orc_params(sp))
    assert_same([got[i] for i in sample], exp, [qs[i] for i in sample], "batch with over-budget queries")


def test_over_budget_token_inside_a_text(eng_oracle):
    import analiticcl_b200 as A
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build()
    text = "I tink you are rihgt about counterrevolutionaries " + "z" * 300 + " and teh rest"
    sp = A.SearchParameters(max_anagram_distance=0.5, max_edit_distance=2, max_ngram=1)
    got = m.find_all_matches(text, sp)
    by_input = {g["input"]: [v["text"] for v in g["variants"]] for g in got}
    assert by_input["counterrevolutionaries"] == [] and by_input["z" * 300] == []
    assert by_input["tink"][:1] and "think" in by_input["tink"] and "right" in by_input["rihgt"] and "the" in by_input["teh"]
    assert [g["input"] for g in got] == text.split(" ")
