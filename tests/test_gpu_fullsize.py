"""GPU, BASELINE.json's full size (cfg 2: 1 M OCR-noise queries on nld with frequencies and late confusables):
size-independent properties of the whole batch, through the C ABI, plus an oracle spot check."""
import ctypes as C

import numpy as np
import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu

N = 1_000_000


def _batch(L, model, queries, sp):
    from analiticcl_b200 import _capi
    blob, offs = _capi.pack(queries)
    rs = C.c_void_p()
    assert L.anl_find_variants_batch(model._h, blob, _capi.u64ptr(offs), len(queries), C.byref(sp.data), C.byref(rs)) == 0, \
        L.anl_last_error()
    n = len(queries)
    o = np.ctypeslib.as_array(L.anl_result_set_offsets(rs), shape=(n + 1,)).copy()
    total = int(o[n])
    raw = np.ctypeslib.as_array(C.cast(L.anl_result_set_variants(rs), C.POINTER(C.c_uint64)), shape=(total, 4)).copy()
    L.anl_result_set_free(rs)
    return o, raw  # raw columns: vocab_id, dist_score bits, freq_score bits, via


def test_cfg2_full_batch_properties():
    import analiticcl_b200 as A
    from analiticcl_b200 import _capi
    L = _capi.lib()
    path = workloads.nld_freq_lexicon()
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(path)
    for pat, w in workloads.CFG2_CONFUSABLES:
        m.add_to_confusables(pat, w)
    m.build()
    sp = A.SearchParameters(max_anagram_distance=3, max_edit_distance=3, freq_weight=0.25)
    qs = workloads.cfg2_queries(N, 2003)
    offs, raw = _batch(L, m, qs, sp)
    counts = np.diff(offs.astype(np.int64))
    assert len(counts) == N and counts.min() >= 0
    dist = raw[:, 1].copy().view(np.float64)
    freq = raw[:, 2].copy().view(np.float64)
    assert np.all(raw[:, 3] == np.uint64(0xFFFFFFFFFFFFFFFF))            # via = None
    assert np.all(raw[:, 0] < np.uint64(L.anl_model_vocab_size(m._h)))                 # no flag bits leak into vocab ids
    assert np.all((freq >= 0.0) & (freq <= 1.0)) and np.all(np.isfinite(dist)) and np.all(dist > 0.0)
    # ranking: inside every list the combined score (src/types.rs:335-341) never increases
    fw = np.float64(np.float32(0.25))
    score = (dist + fw * freq) / (1.0 + fw)
    inner = np.ones(len(score), dtype=bool)
    inner[offs[:-1][counts > 0].astype(np.int64)] = False               # first element of each list
    assert np.all(score[1:][inner[1:]] <= score[:-1][inner[1:]])
    # batch-composition independence: the same queries in reverse order, in one call, give the same lists
    offs_r, raw_r = _batch(L, m, qs[::-1], sp)
    counts_r = np.diff(offs_r.astype(np.int64))
    assert np.array_equal(counts_r[::-1], counts)
    idx = np.flatnonzero(counts > 0)
    pick = idx[:: max(1, len(idx) // 200_000)]                          # compare 200 k lists element by element
    for i in pick[:200_000]:
        a = raw[offs[i]:offs[i + 1]]
        j = N - 1 - i
        b = raw_r[offs_r[j]:offs_r[j + 1]]
        assert np.array_equal(a, b), i
    # checksum of checksums over everything: an order-sensitive sum per list, compared list by list
    def list_sums(o, r):
        c = np.diff(o.astype(np.int64))
        pos = np.arange(len(r), dtype=np.uint64) - np.repeat(o[:-1].astype(np.uint64), c)
        key = ((r[:, 0] * np.uint64(0x9E3779B97F4A7C15)) ^ r[:, 1] ^ (r[:, 2] << np.uint64(1))) * (pos * np.uint64(2) + np.uint64(1))
        sums = np.zeros(len(c), dtype=np.uint64)
        nz = c > 0
        sums[nz] = np.add.reduceat(key, o[:-1].astype(np.int64)[nz])
        return sums
    assert np.array_equal(list_sums(offs, raw), list_sums(offs_r, raw_r)[::-1])
    # against the oracle
    o = orc.OracleModel(alphabet_file=workloads.ALPHABET)
    o.read_lexicon(path)
    for pat, w in workloads.CFG2_CONFUSABLES:
        o.add_to_confusables(pat, w)
    o.build()
    # 24 000 of the 1 M queries (four slices across the batch; ~4 s of oracle time on 16 cores): bit-exact lists
    op = orc.make_params(max_anagram_distance=3, max_edit_distance=3, freq_weight=0.25)
    for lo in (0, 333_000, 666_000, N - 6000):
        exp = o.find_variants_batch(qs[lo:lo + 6000], op, threads=0)
        for k, e in enumerate(exp):
            i = lo + k
            got = [(int(v), float(d), float(f)) for v, d, f in
                   zip(raw[offs[i]:offs[i + 1], 0], dist[offs[i]:offs[i + 1]], freq[offs[i]:offs[i + 1]])]
            assert got == [(int(v), float(d), float(f)) for v, d, f in e], (i, qs[i])
