"""CPU-only: persistence of the built index (anl_model_save_index / anl_model_load_index).  Without a GPU build()
and load_index() stop at the upload ("no CPU fallback"), but the host copy of the index is complete by then, so
the file round trip, the validation of the file and the equality of a loaded and a built index can be checked
here; tests/test_gpu_zz_consolidation.py checks that lookups on a loaded index equal those on a built one."""
import os

import pytest
import torch

import workloads

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="host-only variant of the persistence test")


def model(A, lexicon="eng", extra=()):
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.lexicon_path(lexicon))
    for w in extra:
        m.add_to_vocabulary(w, 1, A.VocabParams())
    return m


def host_build(m):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.build()


def stats(m):
    return (m.index_size(), m.instance_count(), m.max_key_bits(), [m.anagram_count_of_length(i) for i in range(1, 30)])


def test_index_file_round_trip(tmp_path):
    import analiticcl_b200 as A
    a = model(A)
    with pytest.raises(RuntimeError, match="not been built"):
        a.save_index(str(tmp_path / "x.idx"))
    host_build(a)
    f1, f2 = str(tmp_path / "eng.idx"), str(tmp_path / "eng2.idx")
    a.save_index(f1)
    assert os.path.getsize(f1) > 30_000_000  # table + Bloom words + postings + instance rows of eng
    b = model(A)
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # the file is accepted, only the upload is missing here
        b.load_index(f1)
    assert stats(b) == stats(a) and stats(a)[0] == 108802 and stats(a)[1] == 119773
    b.save_index(f2)
    assert open(f1, "rb").read() == open(f2, "rb").read()  # loaded index == built index, bit for bit


def test_index_file_is_validated(tmp_path):
    import analiticcl_b200 as A
    a = model(A, "eng")
    host_build(a)
    good = str(tmp_path / "eng.idx")
    a.save_index(good)
    raw = open(good, "rb").read()
    # another vocabulary (one more entry, or another lexicon): refused by the fingerprint
    for other in (model(A, "eng", extra=["zzyzx"]), model(A, "nld")):
        with pytest.raises(RuntimeError, match="different vocabulary"):
            other.load_index(good)
        assert other.index_size() == 0
    fresh = lambda: model(A, "eng")  # noqa: E731
    cases = {
        "missing.idx": (None, "cannot open"),
        "foreign.idx": (b"not an index" * 1000, "not an analiticcl_b200 index file"),
        "truncated.idx": (raw[: len(raw) // 2], "truncated or corrupt"),
        "layout.idx": (raw[:12] + b"\x11\x00\x00\x00" + raw[16:], "different data layout"),
    }
    for name, (content, msg) in cases.items():
        path = str(tmp_path / name)
        if content is not None:
            open(path, "wb").write(content)
        with pytest.raises(RuntimeError, match=msg):
            fresh().load_index(path)
    # a flipped posting (anagram rank out of range) is caught by the structural checks
    bad = bytearray(raw)
    m = fresh()
    host_build(m)
    n_ana = m.index_size()
    import struct
    hit = raw.rfind(struct.pack("<I", n_ana - 1))  # some u32 holding the last anagram rank (postings / offsets)
    assert hit > 0
    bad[hit:hit + 4] = struct.pack("<I", 0xFFFFFFF0)
    open(str(tmp_path / "flipped.idx"), "wb").write(bytes(bad))
    with pytest.raises(RuntimeError, match="inconsistent|truncated"):
        fresh().load_index(str(tmp_path / "flipped.idx"))


def test_tampered_index_files_are_refused(tmp_path):
    """Bit rot or a hostile writer must give an error, never a silently wrong (or spinning) lookup: header fields
    the kernels size loops and shared arrays from are range-checked, derived tables (primes, charcount mask) are
    recomputed instead of trusted, and every array byte is covered by the content checksum in the header."""
    import struct
    import analiticcl_b200 as A
    a = model(A, "eng")
    host_build(a)
    good = str(tmp_path / "eng.idx")
    a.save_index(good)
    raw = open(good, "rb").read()
    # header layout: magic 8 | header_bytes, slot_bytes, key_bytes, max_k | fingerprint 8 | shard, n_shards, norm_stride,
    # max_charcount, max_len, max_key_bits | sd | reserved | table_keys 8 | checksum 8 | charcount_mask 32 | prime_of 1024
    OFF = dict(shard=32, n_shards=36, norm_stride=40, max_charcount=44, max_len=48, max_key_bits=52, sd=56, table_keys=64,
               checksum=72, charcount_mask=80, prime_of=112, arrays=112 + 1024)
    assert struct.unpack_from("<i", raw, OFF["sd"])[0] == 1 and struct.unpack_from("<I", raw, OFF["max_len"])[0] == 24
    assert struct.unpack_from("<I", raw, OFF["prime_of"])[0] == 2 and struct.unpack_from("<I", raw, OFF["prime_of"] + 4)[0] == 3

    def patched(off, data):
        b = bytearray(raw)
        b[off:off + len(data)] = data
        return bytes(b)
    first_keys = OFF["arrays"] + 16  # ana_key array: count, element size, then the keys
    cases = {
        "sd7": patched(OFF["sd"], struct.pack("<i", 7)),
        "maxcc": patched(OFF["max_charcount"], struct.pack("<I", 100000)),
        "maxlen": patched(OFF["max_len"], struct.pack("<I", 250)),
        "stride": patched(OFF["norm_stride"], struct.pack("<I", 4096)),
        "shard": patched(OFF["n_shards"], struct.pack("<I", 0)),
        "primes": patched(OFF["prime_of"], b"\0" * 1024),
        "ccmask": patched(OFF["charcount_mask"], struct.pack("<Q", 0xFFFF)),
        "tablekeys": patched(OFF["table_keys"], struct.pack("<Q", 1)),
        "garbage_keys": patched(first_keys, os.urandom(100 * 24)),
        "one_bit": patched(len(raw) - 4096, bytes([raw[len(raw) - 4096] ^ 0x10])),
        "checksum": patched(OFF["checksum"], struct.pack("<Q", 12345)),
    }
    for name, content in cases.items():
        path = str(tmp_path / (name + ".idx"))
        open(path, "wb").write(content)
        m = model(A, "eng")
        with pytest.raises(RuntimeError, match="inconsistent|corrupt"):
            m.load_index(path)
        assert m.index_size() == 0, name


def test_frequency_change_invalidates_the_built_index():
    """The device holds its own copy of the frequencies; the reference reads decoder[].frequency live
    (src/lib.rs:1456).  Changing an indexed entry's frequency after build() must not leave a stale index in use."""
    import analiticcl_b200 as A
    m = model(A, "eng")
    host_build(m)
    assert m.index_size() == 108802
    m.add_to_vocabulary("separate", 1, A.VocabParams())      # Max(1, 1): nothing changes, the index stays valid
    assert m.index_size() == 108802
    m.add_to_vocabulary("separate", 500, A.VocabParams())    # frequency 1 -> 500
    assert m.index_size() == 0
    with pytest.raises(RuntimeError, match="not been built"):
        m.find_variants("seperate", A.SearchParameters())


def test_sharded_index_round_trip(tmp_path):
    """A lexicon shard (its own anagram subset + the global gather ids) survives the file as well, and a shard's
    file is not mistaken for another shard's: the shard coordinates are part of the file."""
    import analiticcl_b200 as A
    from analiticcl_b200 import sharded

    def shard_model(s, n):
        m = sharded.ShardedVariantModel(workloads.ALPHABET, A.Weights())
        m.read_lexicon(workloads.lexicon_path("eng"))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.build(shard=s, n_shards=n)
        return m
    parts = [shard_model(s, 3) for s in range(3)]
    assert sum(p.index_size() for p in parts) == 108802 and sum(p.instance_count() for p in parts) == 119773
    files = []
    for s, p in enumerate(parts):
        files.append(str(tmp_path / f"eng.{s}of3.idx"))
        p.save_index(files[-1])
    assert len({open(f, "rb").read() for f in files}) == 3
    again = model(A)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        again.load_index(files[1])
    assert (again.index_size(), again.instance_count()) == (parts[1].index_size(), parts[1].instance_count())
    other = sharded.ShardedVariantModel(workloads.ALPHABET, A.Weights())
    other.read_lexicon(workloads.lexicon_path("eng"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        other.load_index(files[2])
    assert (other.shard, other.n_shards) == (2, 3)
    again.save_index(str(tmp_path / "copy.idx"))
    assert open(str(tmp_path / "copy.idx"), "rb").read() == open(files[1], "rb").read()


def test_parallel_build_is_deterministic(tmp_path):
    """The index build runs on all cores (keys, instance sort, rows, posting generation and sort); its output must
    not depend on the number of threads: one thread (a plain serial pass) and three threads (uneven ranges) give
    the file the default thread count gives."""
    import subprocess
    import sys
    import analiticcl_b200 as A
    m = model(A, "nld")
    host_build(m)
    ref = str(tmp_path / "nld.idx")
    m.save_index(ref)
    code = ("import sys; sys.path.insert(0, %r); import workloads, analiticcl_b200 as A\n"
            "m = A.VariantModel(workloads.ALPHABET, A.Weights()); m.read_lexicon(workloads.lexicon_path('nld'))\n"
            "try:\n    m.build()\nexcept RuntimeError:\n    pass\n"
            "m.save_index(sys.argv[1])\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for threads in ("1", "3"):
        out = str(tmp_path / f"nld.t{threads}.idx")
        env = dict(os.environ, ANL_HOST_THREADS=threads)
        subprocess.run([sys.executable, "-c", code, out], check=True, env=env, timeout=300)
        assert open(out, "rb").read() == open(ref, "rb").read(), f"{threads} thread(s)"
