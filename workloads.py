"""Input data and synthetic workloads for tests and bench.py.

All generators are seeded (NumPy default_rng) and deterministic.  Generated files live under
data/_cache/ (git-ignored; travels to the GPU box with the snapshot if already generated, and is
regenerated there otherwise -- generation takes seconds).
"""
import gzip
import os
import shutil

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(ROOT, "data")
CACHE = os.path.join(DATA, "_cache")

ALPHABET = os.path.join(DATA, "simple.alphabet.tsv")
AMPHIBIANS = os.path.join(DATA, "amphibians.tsv")
REPTILES = os.path.join(DATA, "reptiles.tsv")


def _ensure_cache():
    os.makedirs(CACHE, exist_ok=True)


def lexicon_path(name):
    """name in {'eng', 'nld'} -> path of the decompressed aspell lexicon."""
    _ensure_cache()
    dst = os.path.join(CACHE, f"{name}.aspell.lexicon")
    if not os.path.exists(dst):
        tmp = dst + f".tmp{os.getpid()}"
        with gzip.open(os.path.join(DATA, f"{name}.aspell.lexicon.gz"), "rb") as fi, open(tmp, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        os.replace(tmp, dst)
    return dst


def read_words(name):
    with open(lexicon_path(name), encoding="utf-8") as f:
        return [line.rstrip("\n").split("\t")[0] for line in f if line.strip()]


_LETTERS = "abcdefghijklmnopqrstuvwxyz"


def _edit(word, rng):
    """One random edit: insert / delete / substitute / transpose-adjacent, letters uniform a-z."""
    op = rng.integers(0, 4)
    n = len(word)
    if op == 0 or n == 0:  # insert
        i = int(rng.integers(0, n + 1))
        return word[:i] + _LETTERS[int(rng.integers(0, 26))] + word[i:]
    if op == 1 and n > 1:  # delete
        i = int(rng.integers(0, n))
        return word[:i] + word[i + 1:]
    if op == 2:  # substitute
        i = int(rng.integers(0, n))
        return word[:i] + _LETTERS[int(rng.integers(0, 26))] + word[i + 1:]
    if n > 1:  # transpose adjacent
        i = int(rng.integers(0, n - 1))
        return word[:i] + word[i + 1] + word[i] + word[i + 2:]
    return word


def misspellings(words, n, seed, min_len=4, max_len=14, edit_probs=((1, 0.6), (2, 0.4))):
    """cfg-1 style generator: uniform source word with min_len <= len <= max_len, e random edits."""
    rng = np.random.default_rng(seed)
    pool = [w for w in words if min_len <= len(w) <= max_len]
    es, ps = zip(*edit_probs)
    out = []
    src = rng.integers(0, len(pool), size=n)
    ne = rng.choice(es, size=n, p=ps)
    for k in range(n):
        w = pool[int(src[k])]
        for _ in range(int(ne[k])):
            w = _edit(w, rng)
        if not w:
            w = pool[int(src[k])]
        out.append(w)
    return out


_OCR_CONF = [("f", "s"), ("s", "f"), ("i", "l"), ("l", "i"), ("c", "e"), ("e", "c"), ("u", "n"), ("n", "u"),
             ("rn", "m"), ("m", "rn"), ("y", "ij"), ("ij", "y"), ("h", "b"), ("b", "h")]


def ocr_noise(words, n, seed, p_conf=0.08, p_sub=0.02, p_del=0.01, p_ins=0.01):
    """cfg-2 style generator: per-character OCR confusions + random sub/del/ins."""
    rng = np.random.default_rng(seed)
    pool = [w for w in words if len(w) >= 2]
    conf = {}
    for a, b in _OCR_CONF:
        conf.setdefault(a, []).append(b)
    out = []
    src = rng.integers(0, len(pool), size=n)
    for k in range(n):
        w = pool[int(src[k])]
        res = []
        i = 0
        r = rng.random(size=len(w) + 1)
        while i < len(w):
            x = r[i]
            two = w[i:i + 2]
            if x < p_conf and (two in conf or w[i] in conf):
                if two in conf and len(two) == 2:
                    res.append(conf[two][0])
                    i += 2
                    continue
                opts = conf[w[i]]
                res.append(opts[int(rng.integers(0, len(opts)))])
            elif x < p_conf + p_sub:
                res.append(_LETTERS[int(rng.integers(0, 26))])
            elif x < p_conf + p_sub + p_del:
                pass
            elif x < p_conf + p_sub + p_del + p_ins:
                res.append(w[i])
                res.append(_LETTERS[int(rng.integers(0, 26))])
            else:
                res.append(w[i])
            i += 1
        s = "".join(res)
        out.append(s if s else w)
    return out


def zipf_frequencies(n, seed, s=1.07, top=1_000_000):
    """Zipf(s) frequencies assigned to entries by a seeded permutation (cfg 2)."""
    rng = np.random.default_rng(seed)
    ranks = rng.permutation(n) + 1
    f = np.maximum(1, (top / np.power(ranks.astype(np.float64), s)).astype(np.int64))
    return f


def nld_freq_lexicon(seed=2002):
    """cfg 2 lexicon: nld aspell words + synthetic Zipf frequency column -> path of `word\\tfreq` TSV."""
    _ensure_cache()
    dst = os.path.join(CACHE, f"nld.freq{seed}.lexicon")
    if not os.path.exists(dst):
        words = read_words("nld")
        freqs = zipf_frequencies(len(words), seed)
        tmp = dst + f".tmp{os.getpid()}"
        with open(tmp, "w", encoding="utf-8") as f:
            for w, q in zip(words, freqs):
                f.write(f"{w}\t{int(q)}\n")
        os.replace(tmp, dst)
    return dst


CFG2_CONFUSABLES = [("-[f]+[s]", 1.1), ("-[y]+[i]", 1.1), ("-[c]+[e]", 0.95), ("-[l]+[i]", 1.05), ("-[u]+[n]", 0.95)]


def pack(queries):
    """list[str] -> (utf-8 blob bytes, uint64 offsets[n+1])."""
    enc = [q.encode("utf-8") for q in queries]
    offs = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        offs[1:] = np.cumsum([len(e) for e in enc], dtype=np.uint64)
    return b"".join(enc), offs


def cfg1_queries(n=10_000, seed=1001):
    return misspellings(read_words("eng"), n, seed)


def cfg2_queries(n=1_000_000, seed=2003):
    _ensure_cache()
    dst = os.path.join(CACHE, f"cfg2.q{n}.s{seed}.txt")
    if os.path.exists(dst):
        with open(dst, encoding="utf-8") as f:
            return f.read().split("\n")[:-1]
    qs = ocr_noise(read_words("nld"), n, seed)
    tmp = dst + f".tmp{os.getpid()}"
    with open(tmp, "w", encoding="utf-8") as f:
        f.write("\n".join(qs) + "\n")
    os.replace(tmp, dst)
    return qs


def cfg4_queries(n=1_000_000, seed=4001):
    return misspellings(read_words("eng"), n, seed, min_len=8, max_len=24, edit_probs=((2, 1 / 3), (3, 1 / 3), (4, 1 / 3)))


def cfg5_lexicon(n_entries=10_000_000, seed=5001, max_len=24):
    """cfg 5: synthetic corpus lexicon of `n_entries` unique entries built by concatenating two eng
    entries (total length <= max_len symbols so keys stay <= 169 bits), with a Zipf frequency column.
    -> path of the `word\tfreq` TSV (generated on first use; ~160 MB for 10 M entries)."""
    _ensure_cache()
    dst = os.path.join(CACHE, f"cfg5.n{n_entries}.s{seed}.lexicon")
    if os.path.exists(dst):
        return dst
    rng = np.random.default_rng(seed)
    words = [w for w in read_words("eng") if w.isascii() and w.isalpha() and 2 <= len(w) <= max_len - 2]
    warr = np.array(words)
    wlen = np.array([len(w) for w in words], dtype=np.int64)
    nw = len(words)
    pairs = np.empty(0, dtype=np.int64)
    while len(pairs) < n_entries:
        m = int((n_entries - len(pairs)) * 1.6) + 1000
        a = rng.integers(0, nw, size=m)
        b = rng.integers(0, nw, size=m)
        ok = wlen[a] + wlen[b] <= max_len
        pairs = np.unique(np.concatenate([pairs, a[ok] * nw + b[ok]]))
    pairs = rng.permutation(pairs)[:n_entries]
    a, b = pairs // nw, pairs % nw
    entries = np.char.add(warr[a], warr[b])
    entries = np.unique(entries)  # different pairs can spell the same string
    freqs = zipf_frequencies(len(entries), seed + 1)
    tmp = dst + f".tmp{os.getpid()}"
    with open(tmp, "w", encoding="ascii") as f:
        chunk = 500_000
        for i in range(0, len(entries), chunk):
            f.write("\n".join(f"{w}\t{q}" for w, q in zip(entries[i:i + chunk].tolist(), freqs[i:i + chunk].tolist())))
            f.write("\n")
    os.replace(tmp, dst)
    return dst


def cfg5_queries(n=1_000_000, seed=5002, n_entries=10_000_000, lex_seed=5001):
    """Misspellings (cfg-1 generator) of entries of the cfg-5 lexicon."""
    _ensure_cache()
    dst = os.path.join(CACHE, f"cfg5.q{n}.s{seed}.n{n_entries}.txt")
    if os.path.exists(dst):
        with open(dst, encoding="utf-8") as f:
            return f.read().split("\n")[:-1]
    rng = np.random.default_rng(seed)
    path = cfg5_lexicon(n_entries, lex_seed)
    # sample source entries by line without loading 10 M Python strings
    with open(path, "rb") as f:
        data = f.read()
    starts = np.flatnonzero(np.frombuffer(data, dtype=np.uint8) == 10) + 1
    starts = np.concatenate([[0], starts[:-1]])
    pick = rng.integers(0, len(starts), size=n)
    src = []
    for p in starts[pick]:
        e = data.index(b"\t", p)
        src.append(data[p:e].decode("ascii"))
    qs = misspellings(src, n, seed + 1, min_len=4, max_len=24)
    tmp = dst + f".tmp{os.getpid()}"
    with open(tmp, "w", encoding="utf-8") as f:
        f.write("\n".join(qs) + "\n")
    os.replace(tmp, dst)
    return qs


def cfg3_text(n_tokens=1_000_000, seed=3001, misspell=0.10):
    """cfg 3: running text of `n_tokens` tokens drawn Zipf-like over the eng lexicon, 10 % misspelled
    (cfg-1 generator), sentences of 8-20 tokens terminated by ". " (a multi-char boundary is Hard,
    src/search.rs:245-247), single spaces otherwise."""
    rng = np.random.default_rng(seed)
    words = [w for w in read_words("eng") if w.isascii() and w.isalpha()]
    nw = len(words)
    # Zipf over a seeded permutation of the lexicon
    ranks = np.minimum(nw - 1, (rng.zipf(1.15, size=n_tokens) - 1)).astype(np.int64)
    perm = rng.permutation(nw)
    idx = perm[ranks]
    noisy = rng.random(n_tokens) < misspell
    toks = []
    for i in range(n_tokens):
        w = words[int(idx[i])]
        if noisy[i]:
            w = _edit(w, rng) or w
        toks.append(w)
    out = []
    i = 0
    while i < n_tokens:
        k = int(rng.integers(8, 21))
        out.append(" ".join(toks[i:i + k]))
        i += k
    return ". ".join(out) + "."
