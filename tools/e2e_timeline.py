#!/usr/bin/env python3
"""usage: e2e_timeline.py [n_queries [n_devices]]
One anl_find_variants_batch call of cfg 2 (1 M queries) with ANL_TIMELINE=1: per chunk, when its stages ended on the
device (CUDA events against one reference) and when the host launched / placed / finished downloading it."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
import analiticcl_b200 as A  # noqa: E402
from analiticcl_b200 import _capi  # noqa: E402

L = _capi.lib()
m = A.VariantModel(workloads.ALPHABET, A.Weights())
m.read_lexicon(workloads.nld_freq_lexicon())
for pat, w in workloads.CFG2_CONFUSABLES:
    m.add_to_confusables(pat, w)
n_dev = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if n_dev > 1:
    m.build(devices=list(range(n_dev)))  # one process, a replica on every GPU
else:
    m.build(device=0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
qs = workloads.cfg2_queries(1_000_000, 2003)
qs = (qs * ((n + len(qs) - 1) // len(qs)))[:n]
sp = A.SearchParameters(max_anagram_distance=3, max_edit_distance=3, freq_weight=0.25)
blob, offs = _capi.pack(qs)
offs_p = _capi.u64ptr(offs)
for it in range(4):
    if it == 3:
        os.environ["ANL_TIMELINE"] = "1"
    rs = C.c_void_p()
    t0 = time.perf_counter()
    assert L.anl_find_variants_batch(m._h, blob, offs_p, n, C.byref(sp.data), C.byref(rs)) == 0, L.anl_last_error()
    dt = time.perf_counter() - t0
    L.anl_result_set_free(rs)
    print("call %d: %.2f ms = %.2f M q/s" % (it, dt * 1e3, n / dt / 1e6), file=sys.stderr, flush=True)
