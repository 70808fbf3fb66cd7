"""CPU-only: the C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    from analiticcl_b200 import build, _capi
    build.build()
    return ctypes.CDLL(_capi.SO_PATH)


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "analiticcl_b200.h"), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(anl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(so):
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(so, n)]
    assert not missing, missing


def test_ctypes_stub_covers_header():
    from analiticcl_b200 import _capi
    assert sorted(_capi.SIGNATURES) == declared_symbols()


def test_host_side_without_gpu():
    """Model construction, normalisation and hashing are host code; build() must fail loudly
    (no CPU fallback) when there is no device."""
    import torch
    import analiticcl_b200 as A
    import workloads
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.AMPHIBIANS)
    assert m.anahash("least") == 1227306
    assert m.normalize("aFé") == [0, 32, 1]  # F is not in simple.alphabet.tsv: UNK = alphabet.len()+1
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.build()
        with pytest.raises(RuntimeError, match="not been built"):
            m.find_variants("frog", A.SearchParameters())


def test_search_parameters_surface():
    import analiticcl_b200 as A
    p = A.SearchParameters()
    assert (p.max_anagram_distance, p.max_edit_distance, p.max_matches) == (3, 3, 20)
    assert (p.score_threshold, p.cutoff_threshold, p.max_ngram, p.freq_weight) == (0.25, 2.0, 3, 0.0)
    p = A.SearchParameters(max_edit_distance=(0.5, 4), max_anagram_distance=0.25, stop_at_exact_match=True, bogus=1)
    assert p.max_edit_distance == (0.5, 4) and p.max_anagram_distance == 0.25 and p.stop_at_exact_match
    assert A.SearchParameters(max_edit_distance="0.5;4").max_edit_distance == (0.5, 4)
    w = A.Weights(ld=0.6)
    assert w.to_dict() == {"ld": 0.6, "lcs": 0.125, "prefix": 0.125, "suffix": 0.125, "case": 0.125}
    v = A.VocabParams(freq_column=2, vocabtype="TRANSPARENT", freqhandling="sum")
    assert v.freq_column == 2 and v.data.vocab_type == 5 and v.data.freq_handling == 0
    # the attributes the reference binding makes assignable (#[setter], bindings/python/src/lib.rs:262-446)
    p = A.SearchParameters()
    p.max_seq, p.lm_weight, p.variantmodel_weight, p.contextrules_weight, p.context_weight, p.freq_weight = 7, 2.5, 1.5, 0.5, 0.25, 0.75
    p.single_thread, p.stop_at_exact_match, p.consolidate_matches, p.unicodeoffsets = True, True, False, True
    p.max_anagram_distance, p.max_edit_distance, p.max_ngram, p.max_matches = 2, (0.5, 3), 2, 5
    assert p.to_dict() == {"max_anagram_distance": 2, "max_edit_distance": (0.5, 3), "max_matches": 5, "score_threshold": 0.25,
                           "cutoff_threshold": 2.0, "max_ngram": 2, "max_seq": 7, "single_thread": True, "freq_weight": 0.75,
                           "lm_weight": 2.5, "contextrules_weight": 0.5, "variantmodel_weight": 1.5, "consolidate_matches": False,
                           "unicodeoffsets": True}
    assert p.stop_at_exact_match and p.context_weight == 0.25
    with pytest.raises(ValueError):
        p.max_edit_distance = "nonsense"
    with pytest.raises(AttributeError):
        p.score_threshold = 0.1  # (no setter in the reference binding either)


def test_variant_list_loading_matches_oracle(tmp_path):
    """Host side of the variant lists (no GPU needed): read_variants / add_variant build the same vocabulary as the
    oracle -- ids in order of first mention, frequencies, the TRANSPARENT flag -- for both file layouts."""
    import ctypes as C
    import analiticcl_b200 as A
    from analiticcl_b200 import _capi
    from oracle import orc
    import workloads
    plain = tmp_path / "v.tsv"
    plain.write_text("separate\tseperate\t0.9\tseparete\t0.8\nhouse\thuose\t0.5\nseparate\tseperate\t0.7\nodd\todd\t1.0\n")
    freq = tmp_path / "vf.tsv"
    freq.write_text("separate\t100\tseperate\t0.9\t7\nhouse\t50\thuose\t0.5\t2\thuose\t0.4\t9\n")
    for path, transparent in ((plain, False), (plain, True), (freq, True)):
        m = A.VariantModel(workloads.ALPHABET, A.Weights())
        o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
        rid = m.add_to_vocabulary("separated", 3)
        assert rid == o.add_to_vocabulary("separated", 3)
        m.read_variants(str(path), transparent=transparent)
        o.read_variants(str(path), transparent=transparent)
        assert m.add_variant(rid, "seperated", 0.6) and o.add_variant(rid, "seperated", 0.6, index=0)
        assert not m.add_variant(rid, "separated", 1.0)  # a variant of itself is refused
        n = _capi.lib().anl_model_vocab_size(m._h)
        assert n == o.vocab_size()
        for vid in range(3, n):
            info = m._vocab(vid)
            assert C.string_at(info.text, info.text_len).decode() == o.vocab_text(vid)
            assert info.frequency == o.vocab_freq(vid), o.vocab_text(vid)
            assert info.vocabtype == o.vocab_type(vid), o.vocab_text(vid)


def test_no_exception_crosses_the_c_boundary(so):
    """A C++ exception inside an entry point (here: std::bad_alloc from an absurd variant count) comes back as a
    status code with a message, not as a terminate() in the caller's process."""
    from analiticcl_b200 import _capi
    L = _capi.lib()
    looked = (ctypes.c_uint8 * 1)(1)
    offs = (ctypes.c_uint64 * 2)(0, 1 << 57)
    ms = ctypes.c_void_p()
    rc = L.anl_debug_match_set_build(b"a", 1, 1, 0, looked, offs, None, 1, ctypes.byref(ms))
    assert rc != 0 and b"out of memory" in L.anl_last_error()


def test_build_limits_are_reported_for_the_first_offending_entry():
    """The host index build rejects what the 192-bit device keys cannot hold; the keys are
    computed on all cores and the entry reported is the first one in vocabulary order, as in a serial pass."""
    import analiticcl_b200 as A
    import workloads
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))  # > 100 k entries: several threads
    m.add_to_vocabulary("z" * 40, 1, A.VocabParams())
    m.add_to_vocabulary("y" * 300, 1, A.VocabParams())
    m.add_to_vocabulary("z" * 41, 1, A.VocabParams())
    with pytest.raises(RuntimeError, match="exceeds 192 bits: z{40}$"):
        m.build()
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.add_to_vocabulary("y" * 300, 1, A.VocabParams())
    m.add_to_vocabulary("z" * 40, 1, A.VocabParams())
    with pytest.raises(RuntimeError, match="longer than"):
        m.build()
