"""GPU parity on BASELINE config 5's regime: the synthetic corpus lexicon (concatenated eng entries, Zipf
frequencies; here its 2 M-entry instance, the size bench.py already caches) is the only one whose index leaves the
L2-resident regime: the Bloom filter takes the dense branch (several keys per 64-bit word, percent-level false
positives that the exact stage must reject), the table is hundreds of MB and the postings are HBM-resident.
Checked bit-exact against the CPU oracle, unsharded and through the lexicon-sharded path (2 emulated shards)."""
import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu

ENTRIES = 2_000_000
N_QUERIES = 1500


@pytest.fixture(scope="module")
def cfg5():
    lexicon = workloads.cfg5_lexicon(ENTRIES)
    qs = workloads.cfg5_queries(N_QUERIES, 5002, ENTRIES)
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o.read_lexicon(lexicon)
    o.build()
    exp = {}
    for tag, kw in (("k3", dict()), ("k3 freq", dict(freq_weight=0.25, max_matches=10))):
        exp[tag] = (kw, o.find_variants_batch(qs, orc.make_params(**kw)))
    del o
    return lexicon, qs, exp


def test_cfg5_dense_bloom_parity(cfg5, monkeypatch):
    import analiticcl_b200 as A
    from test_gpu_parity import assert_same
    lexicon, qs, exp = cfg5
    # The 10 M-entry lexicon of the bench caps its filter at 128 MB (~6.6 keys per 64-bit word); the same density on
    # the 2 M-entry instance needs the cap lowered: 32 MB -> ~5.7 keys per word, percent-level false positives.
    monkeypatch.setenv("ANL_BLOOM_MAX_MB", "32")
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(lexicon)
    m.build()
    st = m.index_stats()
    assert st["instances"] > 1_990_000 and st["sd"] == 1
    # the dense-Bloom branch of the index build: several keys per 64-bit filter word, table far beyond the L2
    keys_per_word = st["table_keys"] / (st["bloom_bytes"] / 8)
    assert keys_per_word > 4.0, keys_per_word
    assert st["table_bytes"] >= 512 << 20 and st["bloom_bytes"] <= 32 << 20
    for tag, (kw, want) in exp.items():
        got = m.find_variants_raw(qs, A.SearchParameters(**kw))
        assert_same(got, want, qs, "cfg5 " + tag)
    # (entries are concatenations of two words: the neighbourhood is sparse, most queries find just their source)
    assert sum(len(r) for r in exp["k3"][1]) > 0.8 * N_QUERIES  # the comparison is not vacuous


def test_cfg5_sharded_two_emulated_shards(cfg5):
    import analiticcl_b200 as A
    from analiticcl_b200 import sharded
    from test_gpu_parity import assert_same
    lexicon, qs, exp = cfg5
    ms = []
    for s in range(2):
        m = sharded.ShardedVariantModel(workloads.ALPHABET, A.Weights())
        m.read_lexicon(lexicon)
        m.build(device=0, shard=s, n_shards=2)
        ms.append(m)
    kw, want = exp["k3 freq"]
    sp = A.SearchParameters(**kw)
    batches, exports = zip(*[m.score(qs, sp, 0) for m in ms])
    got = sharded.merge_exports_locally(ms, list(batches), list(exports), len(qs))
    assert_same(got, want, qs, "cfg5 sharded x2")
