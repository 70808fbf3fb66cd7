"""GPU: a model whose index was read from a file (anl_model_load_index) answers exactly like the model that built
and saved it -- and like the oracle.  (File round trip and validation: tests/test_host_index_persistence.py.)"""
import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu


def test_loaded_index_equals_built_index(tmp_path, eng_oracle):
    import analiticcl_b200 as A

    def fresh():
        m = A.VariantModel(workloads.ALPHABET, A.Weights())
        m.read_lexicon(workloads.lexicon_path("eng"))
        return m
    built = fresh()
    built.build()
    path = str(tmp_path / "eng.idx")
    built.save_index(path)
    loaded = fresh()
    loaded.load_index(path)
    assert (loaded.index_size(), loaded.instance_count()) == (108802, 119773)
    qs = workloads.misspellings(workloads.read_words("eng"), 1500, 4242) + ["seperate", "a", "x" * 40]
    sp = A.SearchParameters()
    got = loaded.find_variants_raw(qs, sp)
    assert got == built.find_variants_raw(qs, sp)
    assert got == eng_oracle.find_variants_batch(qs, orc.make_params())
    other = A.VariantModel(workloads.ALPHABET, A.Weights())
    other.read_lexicon(workloads.lexicon_path("nld"))
    with pytest.raises(RuntimeError, match="different vocabulary"):
        other.load_index(path)
