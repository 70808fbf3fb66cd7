"""Times the e2e C-ABI call with ANL_PROFILE phase output (developer tool)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import analiticcl_b200 as A, workloads
from analiticcl_b200 import _capi
conf = "--noconf" not in sys.argv
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1_000_000
m = A.VariantModel(workloads.ALPHABET, A.Weights()); m.read_lexicon(workloads.nld_freq_lexicon())
if conf:
    for p, w in workloads.CFG2_CONFUSABLES: m.add_to_confusables(p, w)
m.build()
qs = workloads.cfg2_queries(1_000_000, 2003)[:n]
sp = A.SearchParameters(freq_weight=0.25)
blob, offs = _capi.pack(qs); L = _capi.lib()
for it in range(int(os.environ.get("ITERS", "3"))):
    rs = C.c_void_p(); t = time.perf_counter()
    st = L.anl_find_variants_batch(m._h, blob, _capi.u64ptr(offs), n, C.byref(sp.data), C.byref(rs))
    dt = time.perf_counter() - t
    assert st == 0, L.anl_last_error()
    print(f"iter {it}: {dt*1e3:.1f} ms  {n/dt:,.0f} q/s  results {L.anl_result_set_offsets(rs)[n]}", file=sys.stderr)
    L.anl_result_set_free(rs)
