"""GPU: the batch producer of find_all_matches as kernels (csrc/gpu_segment.cu; src/search.rs:190-336,
src/lib.rs:1822-1903) against the oracle's segmentation and the host producer -- boundaries with strengths, hard-delimited
batches and every 1..max_ngram-gram, element for element; then find_all_matches end to end with either producer."""
import ctypes as C
import random

import pytest

import workloads
from oracle import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from analiticcl_b200 import build, _capi
    build.build()
    return _capi.lib()


def produce(L, device, text, max_ngram):
    raw = text.encode("utf-8")
    cap = max_ngram * (len(raw) + 2)
    b, e = (C.c_uint64 * cap)(), (C.c_uint64 * cap)()
    o, bt = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
    bcap = len(raw) + 2
    bb, be, bs = (C.c_uint64 * bcap)(), (C.c_uint64 * bcap)(), (C.c_int32 * bcap)()
    nb = C.c_uint64(0)
    n = L.anl_debug_segment_text_device(device, raw, len(raw), max_ngram, b, e, o, bt, cap, bb, be, bs, bcap, C.byref(nb))
    assert 0 <= n <= cap, L.anl_last_error()
    assert nb.value <= bcap
    return list(zip(b[:n], e[:n], o[:n], bt[:n])), list(zip(bb[:nb.value], be[:nb.value], bs[:nb.value]))


KATS = ['Hallo allemaal, ik zeg: "Welkom in Aix-les-bains!".', "dit is een mooie test", "dit is een mooie test.", "hello, world!",
        "hello,world!", "hello!", " ", "a", "..", ". a", "a .", "It's a well-known co-operative re_entry; über naïve façade!?  Done",
        "één twee  drie\n\nvier", "trailing space ", " leading", "  two leading", "x.y.z", "日本語 テキスト mixed with latin",
        "𝔘𝔫𝔦𝔠𝔬𝔡𝔢 four-byte 😀 letters and emoji 😀😀 end", "a b", "a  b", "-a-", "'", "naïve"]


def random_text(rng, n_tokens):
    letters = "abcdefghijklmnopqrstuvwxyzéüñßøÀ日本語𝔘"
    seps = [" ", " ", " ", " ", "-", "'", "_", ", ", ". ", "  ", "\n", "!", " — ", "…", "1", " 42 ", "😀"]
    out = []
    if rng.random() < 0.5:
        out.append(rng.choice(seps))
    for _ in range(n_tokens):
        out.append("".join(rng.choice(letters) for _ in range(rng.randint(1, 9))))
        out.append(rng.choice(seps))
    if rng.random() < 0.5:
        out.pop()
    return "".join(out)


def test_device_producer_matches_host_and_oracle(L):
    rng = random.Random(77)
    texts = KATS + [random_text(rng, rng.randint(1, 60)) for _ in range(150)]
    for t in texts:
        for max_ngram in (1, 2, 3, 5):
            host, hb = produce(L, -1, t, max_ngram)
            dev, db = produce(L, 0, t, max_ngram)
            assert dev == host, (t, max_ngram)
            assert db == hb, t
        assert [(b, e, s) for b, e, s in db] == [(b, e, s) for b, e, _, s in orc.find_boundaries(t)], t


def test_device_producer_large_text(L):
    rng = random.Random(5)
    big = workloads.cfg3_text(300000, 11) + random_text(rng, 20000) + " ça va? — oui…  " * 2000 + "".join(KATS) * 50
    assert len(big.encode("utf-8")) > 2_000_000
    for max_ngram in (1, 3):
        host, hb = produce(L, -1, big, max_ngram)
        dev, db = produce(L, 0, big, max_ngram)
        assert len(dev) == len(host) and dev == host
        assert db == hb


def test_find_all_matches_same_with_either_producer(monkeypatch):
    import analiticcl_b200 as A
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path("eng"))
    m.build()
    text = workloads.cfg3_text(30000, 3)
    p = A.SearchParameters(max_ngram=3, max_anagram_distance=3, max_edit_distance=3, consolidate_matches=False)
    monkeypatch.setenv("ANL_SEGMENT", "host")
    a = m.find_all_matches(text, p)
    monkeypatch.setenv("ANL_SEGMENT", "device")
    b = m.find_all_matches(text, p)
    assert len(a) == len(b) > 30000
    assert a == b
    # and through the consolidation, which reuses the producer's segmentation
    p2 = A.SearchParameters(max_ngram=3, max_anagram_distance=3, max_edit_distance=3)
    monkeypatch.setenv("ANL_SEGMENT", "host")
    c = m.find_all_matches(text, p2)
    monkeypatch.setenv("ANL_SEGMENT", "device")
    d = m.find_all_matches(text, p2)
    assert c == d and 0 < len(c) <= len(a)
