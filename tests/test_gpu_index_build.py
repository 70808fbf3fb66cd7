"""The index build on the device (csrc/gpu_build.cu) against the host build: every array of the index is the same bit
for bit, the open-addressing table holds the same slots in a different (equally valid) probe order, and lookups on the
device-built index equal the oracle's."""
import ctypes as C

import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu

NAMES = ["ana_key", "ana_inst_off", "ana_charcount", "inst_vocab", "inst_freq", "inst_gid", "inst_rows", "table", "bloom",
         "post_ana", "post_cls", "active_classes", "slot digest", "reachable", "occupied", "scalars"]


def digest(m):
    from analiticcl_b200 import _capi
    out = (C.c_uint64 * 16)()
    _capi.lib().anl_debug_index_digest(m._h, out, 16)
    return list(out)


def model(A, lexicon, cls=None, confusables=()):
    m = (cls or A.VariantModel)(workloads.ALPHABET, A.Weights())
    m.read_lexicon(lexicon)
    for pat, w in confusables:
        m.add_to_confusables(pat, w)
    return m


@pytest.mark.parametrize("lex", ["eng", "nld freq"])
def test_device_build_equals_host_build(lex):
    import analiticcl_b200 as A
    path = workloads.lexicon_path("eng") if lex == "eng" else workloads.nld_freq_lexicon()
    host, dev = model(A, path), model(A, path)
    host.build(gpu_build=False)
    dev.build(gpu_build=True)
    dh, dd = digest(host), digest(dev)
    for i, name in enumerate(NAMES):
        if name == "table":
            continue
        assert dh[i] == dd[i], f"{lex}: {name} differs between the host build and the device build"
    assert dd[13] == 1 and dh[13] == 1
    assert host.index_stats() == dev.index_stats()


def test_device_built_index_lookups(eng_oracle, tmp_path):
    import analiticcl_b200 as A
    from test_gpu_parity import assert_same, to_orc_params
    m = model(A, workloads.lexicon_path("eng"))
    m.build(gpu_build=True)
    qs = workloads.misspellings(workloads.read_words("eng"), 4000, 606, min_len=2, max_len=18) + ["", "a", "separate"]
    for kw in (dict(), dict(max_anagram_distance=2, max_edit_distance=2), dict(max_matches=3, score_threshold=0.0)):
        sp = A.SearchParameters(**kw)
        assert_same(m.find_variants_raw(qs, sp), eng_oracle.find_variants_batch(qs, to_orc_params(sp), threads=0), qs, f"device build {kw}")
    assert "separate" in m and "seperate" not in m
    # a device-built index goes through the same persistence as a host-built one
    f = str(tmp_path / "eng.gpu.idx")
    m.save_index(f)
    again = model(A, workloads.lexicon_path("eng"))
    again.load_index(f)
    assert digest(again) == digest(m)
    sp = A.SearchParameters()
    assert again.find_variants_raw(qs[:500], sp) == m.find_variants_raw(qs[:500], sp)


def test_device_build_of_a_shard():
    import analiticcl_b200 as A
    from analiticcl_b200 import sharded
    import os
    path = workloads.nld_freq_lexicon()
    digs = {}
    for where in ("0", "1"):
        os.environ["ANL_GPU_BUILD"] = where
        try:
            m = model(A, path, cls=sharded.ShardedVariantModel)
            m.build(device=0, shard=1, n_shards=3)
            digs[where] = digest(m)
        finally:
            del os.environ["ANL_GPU_BUILD"]
    for i, name in enumerate(NAMES):
        if name != "table":
            assert digs["0"][i] == digs["1"][i], f"shard 1/3: {name} differs"
    assert digs["1"][13] == 1


def test_device_build_reports_the_host_builds_errors():
    import analiticcl_b200 as A
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.add_to_vocabulary("z" * 40, 1, A.VocabParams())  # 101^40 needs 267 bits
    m.add_to_vocabulary("frog", 1, A.VocabParams())
    with pytest.raises(RuntimeError, match="exceeds 192 bits: z"):
        m.build(gpu_build=True)
    with pytest.raises(RuntimeError, match="exceeds 192 bits: z"):
        m.build(gpu_build=False)
