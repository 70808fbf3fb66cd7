#!/usr/bin/env python3
"""e2e throughput of anl_find_variants_batch for several workloads x pipeline settings (one process per cell: the model
is rebuilt, the knobs are read per call).  usage: e2e_matrix.py workload[,workload..] "ENV=V ENV=V" ["ENV=V .."] ..."""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(workload):
    import bench
    import analiticcl_b200 as A
    from analiticcl_b200 import _capi
    import workloads
    L = _capi.lib()
    spec = bench.workload_spec(workload)
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    m.read_lexicon(spec["lexicon"])
    for pat, w in spec["confusables"]:
        m.add_to_confusables(pat, w)
    m.build(device=0)
    n = spec["n"]
    qs = spec["queries"](n)
    sp = A.SearchParameters(**spec["params"])
    blob, offs = _capi.pack(qs)
    offs_p = _capi.u64ptr(offs)
    settings = sys.argv[3:]
    for setting in settings:
        for kv in setting.split():
            k, v = kv.split("=")
            os.environ[k] = v
        times = []
        for it in range(7):
            rs = C.c_void_p()
            t0 = time.perf_counter()
            assert L.anl_find_variants_batch(m._h, blob, offs_p, n, C.byref(sp.data), C.byref(rs)) == 0, L.anl_last_error()
            dt = time.perf_counter() - t0
            L.anl_result_set_free(rs)
            if it >= 2:
                times.append(dt)
        print("%-16s %-48s e2e %.2f M q/s (best %.2f)" % (workload, setting, n / (sum(times) / len(times)) / 1e6, n / min(times) / 1e6), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        for w in sys.argv[1].split(","):
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", w] + sys.argv[2:], check=False)
