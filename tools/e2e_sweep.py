#!/usr/bin/env python3
"""e2e throughput of anl_find_variants_batch (cfg2) for a few chunk sizes / pipeline depths; one process per setting
(the knobs are read once per process).  Usage: python tools/e2e_sweep.py [chunk:inflight[:host_threads] ...]"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import workloads
    import analiticcl_b200 as A
    from analiticcl_b200 import _capi
    L = _capi.lib()
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(workloads.nld_freq_lexicon())
    for pat, w in workloads.CFG2_CONFUSABLES:
        m.add_to_confusables(pat, w)
    m.build(device=0)
    n = 1_000_000
    qs = workloads.cfg2_queries(n, 2003)
    sp = A.SearchParameters(max_anagram_distance=3, max_edit_distance=3, freq_weight=0.25)
    blob, offs = _capi.pack(qs)
    offs_p = _capi.u64ptr(offs)
    times = []
    for it in range(8):
        rs = C.c_void_p()
        t0 = time.perf_counter()
        st = L.anl_find_variants_batch(m._h, blob, offs_p, n, C.byref(sp.data), C.byref(rs))
        dt = time.perf_counter() - t0
        assert st == 0, L.anl_last_error()
        L.anl_result_set_free(rs)
        if it >= 2:
            times.append(dt)
    print("chunk=%s inflight=%s threads=%s  e2e %.2f M q/s (best %.2f)" % (
        os.environ.get("ANL_CHUNK", "65536"), os.environ.get("ANL_INFLIGHT", "4"), os.environ.get("ANL_HOST_THREADS", "all"),
        n / (sum(times) / len(times)) / 1e6, n / min(times) / 1e6), flush=True)


if __name__ == "__main__":
    if os.environ.get("ANL_SWEEP_CHILD"):
        child()
    else:
        for spec in (sys.argv[1:] or ["65536:4", "131072:4", "262144:4", "65536:8", "131072:2"]):
            chunk, inflight, *threads = spec.split(":")
            env = dict(os.environ, ANL_SWEEP_CHILD="1", ANL_CHUNK=chunk, ANL_INFLIGHT=inflight)
            if threads:
                env["ANL_HOST_THREADS"] = threads[0]
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
