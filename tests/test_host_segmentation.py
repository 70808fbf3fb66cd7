"""CPU-only: the host-side batch producer of find_all_matches (csrc/search.cpp, multi-threaded boundary scan and
n-gram generation) through the C-ABI test hooks, against the oracle's segmentation and against itself piecewise."""
import ctypes as C

import numpy as np
import pytest

import workloads
from oracle import orc


@pytest.fixture(scope="module")
def L():
    from analiticcl_b200 import build, _capi
    build.build()
    return _capi.lib()


def boundaries(L, text):
    raw = text.encode("utf-8")
    cap = len(raw) + 2
    b, e, s = (C.c_uint64 * cap)(), (C.c_uint64 * cap)(), (C.c_int32 * cap)()
    n = L.anl_debug_find_boundaries(raw, len(raw), b, e, s, cap)
    assert 0 <= n <= cap
    return [(b[i], e[i], s[i]) for i in range(n)]


def segments(L, text, max_ngram):
    raw = text.encode("utf-8")
    cap = max_ngram * (len(raw) + 2)
    b, e = (C.c_uint64 * cap)(), (C.c_uint64 * cap)()
    o, bt = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
    n = L.anl_debug_segment_text(raw, len(raw), max_ngram, b, e, o, bt, cap)
    assert 0 <= n <= cap
    return [(b[i], e[i], o[i], bt[i]) for i in range(n)], raw


KATS = ['Hallo allemaal, ik zeg: "Welkom in Aix-les-bains!".', "dit is een mooie test", "dit is een mooie test.", "hello, world!",
        "hello,world!", "hello!", "", " ", "a", "..", "It's a well-known co-operative re_entry; über naïve façade!?  Done",
        "één twee  drie\n\nvier", "trailing space ", " leading", "x.y.z", "日本語 テキスト mixed with latin"]


def test_boundaries_match_oracle(L):
    for t in KATS:
        assert boundaries(L, t) == [(b, e, s) for b, e, _, s in orc.find_boundaries(t)], t
    # long text: the scan runs on several threads and stitches runs across range borders
    big = workloads.cfg3_text(40000, 5) + " ça va? — oui…  " * 2000 + "".join(KATS) * 50
    assert len(big.encode("utf-8")) > 300_000
    assert boundaries(L, big) == [(b, e, s) for b, e, _, s in orc.find_boundaries(big)]


def test_ngrams_match_oracle_without_hard_boundaries(L):
    """Inside one hard-delimited batch the segments of order n are exactly find_match_ngrams(text, n)."""
    for t in ["dit is een mooie test", "hello,world", "I tink you are rihgt", "it's a well-known co-operative", "a b c d e f g h"]:
        segs, raw = segments(L, t, 3)
        for order in (1, 2, 3):
            got = [raw[b:e].decode("utf-8") for b, e, o, _ in segs if o == order]
            assert got == orc.find_match_ngrams(t, order), (t, order)


def test_large_text_equals_piecewise(L):
    sents = workloads.cfg3_text(50000, 99).split(". ")
    pieces, cur = [], []
    for i, s in enumerate(sents):
        cur.append(s if i % 53 else s + " über-naïve façade's")
        if len(cur) == 120:
            pieces.append(". ".join(cur) + ". ")
            cur = []
    if cur:
        pieces.append(". ".join(cur) + ". ")
    text = "".join(pieces)
    assert len(text.encode("utf-8")) > 350_000
    whole, _ = segments(L, text, 3)
    exp, shift, batch0 = [], 0, 0
    for p in pieces:
        part, raw = segments(L, p, 3)
        exp += [(b + shift, e + shift, o, bt + batch0) for b, e, o, bt in part]
        shift += len(raw)
        batch0 += (max(bt for _, _, _, bt in part) + 1) if part else 0
    assert len(whole) == len(exp)
    assert whole == exp
    # orders ascend inside a batch (the producer relies on a batch's unigrams coming first)
    arr = np.array([(bt, o) for _, _, o, bt in whole])
    same_batch = arr[1:, 0] == arr[:-1, 0]
    assert np.all(arr[1:, 1][same_batch] >= arr[:-1, 1][same_batch])


def test_alphabetic_is_the_derived_property(L):
    """char::is_alphabetic (src/search.rs:198,204) is the derived property Alphabetic, which includes Other_Alphabetic:
    dependent vowel signs and similar marks are part of a word, they do not split it."""
    for cp, alpha in ((0x093F, True), (0x0902, True), (0x0E31, True), (0x0345, True), (0x0301, False), (0x0030, False),
                      (0x2160, True), (0x00AA, True), (0x3042, True), (0x2019, False)):
        assert bool(orc.lib().orc_is_alphabetic(cp)) == alpha, hex(cp)
    assert bool(orc.lib().orc_is_lowercase(0x10780)) and bool(orc.lib().orc_is_lowercase(0x00E9))
    assert not orc.lib().orc_is_lowercase(0x00C9) and not orc.lib().orc_is_lowercase(0x4E00)
    t = "हिंदी भाषा, ภาษาไทย"
    raw = t.encode("utf-8")
    got = boundaries(L, t)
    assert got == [(b, e, s) for b, e, _, s in orc.find_boundaries(t)]
    assert [raw[b:e].decode() for b, e, _ in got] == [" ", ", ", ""]


def test_concurrent_callers_on_large_texts():
    """Several host threads inside the producer at once (ctypes releases the GIL; the reference's find_all_matches is
    &self and thread-safe): while one caller holds the worker pool the others run their parts on short-lived
    threads, whose thread_locals are gone when the parts' results are read -- the per-part scratch must belong to
    the caller.  Runs in a child process with 8 host threads so a crash cannot take the test session down."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, threading, ctypes as C\n"
        "sys.path.insert(0, %r)\n"
        "import workloads\n"
        "from analiticcl_b200 import _capi\n"
        "L = _capi.lib()\n"
        "raw = workloads.cfg3_text(330000, 17).encode('utf-8')\n"
        "assert len(raw) > 2_000_000\n"
        "cap = 3 * (len(raw) // 4)\n"
        "res = [None] * 4\n"
        "def work(i):\n"
        "    b, e = (C.c_uint64 * cap)(), (C.c_uint64 * cap)()\n"
        "    o, bt = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()\n"
        "    out = []\n"
        "    for _ in range(3):\n"
        "        n = L.anl_debug_segment_text(raw, len(raw), 3, b, e, o, bt, cap)\n"
        "        out.append((n, sum(b[k] for k in range(0, min(n, cap), 997)), sum(e[k] for k in range(0, min(n, cap), 991))))\n"
        "    res[i] = out\n"
        "th = [threading.Thread(target=work, args=(i,)) for i in range(4)]\n"
        "[t.start() for t in th]; [t.join() for t in th]\n"
        "assert all(r is not None and len(set(r)) == 1 for r in res), res\n"
        "assert len({r[0] for r in res}) == 1, res\n"
        "print('ok', res[0][0][0])\n") % root
    env = dict(os.environ, ANL_HOST_THREADS="8")
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.returncode, p.stderr[-2000:])
    assert p.stdout.startswith("ok")
