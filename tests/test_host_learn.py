"""CPU-only: the bookkeeping half of learn_variants (src/lib.rs:1106-1130; csrc/host_model.cpp learn_apply) against the
oracle's restatement: frequencies of known inputs (one count per consecutive run), new inputs as TRANSPARENT-only entries,
variant links in both directions, the returned count.  The lookups that feed it are covered on the GPU
(test_gpu_learn.py).  The reference holds no test for learn mode: PARITY UNPINNED beyond this restatement."""
import ctypes as C

import numpy as np

from oracle import orc


def models(words):
    import analiticcl_b200 as A
    o = orc.OracleModel(alphabet_tsv=orc.TEST_ALPHABET_TSV)
    m = A.VariantModel(None, A.Weights(), alphabet_tsv=orc.TEST_ALPHABET_TSV)
    for w in words:
        assert o.add_to_vocabulary(w, 2) == m.add_to_vocabulary(w, 2, A.VocabParams())
    return o, m


def product_apply(m, items):
    from analiticcl_b200 import _capi
    L = _capi.lib()
    blob, offs = _capi.pack([t for t, _, _ in items])
    ids = (C.c_uint64 * max(1, len(items)))(*[int(v) for _, v, _ in items])
    sc = (C.c_double * max(1, len(items)))(*[float(d) for _, _, d in items])
    count = C.c_uint64()
    assert L.anl_debug_learn_apply(m._h, blob, _capi.u64ptr(offs), len(items), ids, sc, C.byref(count)) == 0, L.anl_last_error()
    return count.value


def state(o_or_m, n):
    if isinstance(o_or_m, orc.OracleModel):
        return [(o_or_m.vocab_text(i), o_or_m.vocab_freq(i), o_or_m.vocab_type(i), o_or_m.vocab_links(i)) for i in range(n)]
    out = []
    for i in range(n):
        info = o_or_m._vocab(i)
        out.append((C.string_at(info.text, info.text_len).decode("utf-8"), info.frequency, info.vocabtype, o_or_m.vocab_links(i)))
    return out


def test_learn_apply_by_hand():
    o, m = models(["huis", "huys", "boom"])
    huis, huys, boom = 3, 4, 5
    items = [("huys", huis, 0.75), ("huys", huys, 1.0),  # a known input: one more occurrence; its exact match is not linked
             ("hvis", huis, 0.5), ("hvis", huys, 0.5),   # a new input: added once, linked to both
             ("boom", boom, 1.0), ("hvis", huis, 0.5)]   # a repeated pair is no new reference, but counts again (:1124-1127)
    assert o.learn_apply(items) == product_apply(m, items) == 4
    n = o.vocab_size()
    assert n == 7 and state(o, n) == state(m, n)
    st = state(m, n)
    assert st[huys][1] == 3 and st[boom][1] == 3 and st[huis][1] == 2
    assert st[6][0] == "hvis" and st[6][2] == 4 and st[6][1] == 2  # TRANSPARENT alone (not INDEXED); added with 1, one more run later
    assert st[huis][3][1] == [huys, 6] and st[6][3][0] == [(huis, 0.5), (huys, 0.5), (huis, 0.5)]


def test_learn_apply_random_equals_oracle():
    rng = np.random.default_rng(5)
    words = [f"w{i}" for i in range(30)]
    for _ in range(20):
        o, m = models(words)
        items = []
        for _ in range(int(rng.integers(1, 80))):
            t = str(rng.choice(words + [f"n{i}" for i in range(10)]))
            for _ in range(int(rng.integers(1, 4))):
                items.append((t, int(rng.integers(3, 33)), float(rng.integers(1, 5)) / 4.0))
        assert o.learn_apply(items) == product_apply(m, items)
        n = o.vocab_size()
        assert state(o, n) == state(m, n)
