#!/bin/bash
# GPU index build: tests, then build times host vs device on the 2 M and 10 M cfg-5 lexicons, cfg5:10M bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_index_build.py tests/test_gpu_cfg5.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_build.log; cat gpurun_out/pytest_build.log
cat > /tmp/buildtime.py <<'PY'
import sys, time, os
sys.path.insert(0, os.getcwd())
import workloads, analiticcl_b200 as A
n = int(sys.argv[1])
lex = workloads.cfg5_lexicon(n)
for where in (True, False):
    m = A.VariantModel(workloads.ALPHABET, A.Weights())
    t0 = time.perf_counter(); m.read_lexicon(lex); t1 = time.perf_counter()
    m.build(gpu_build=where); t2 = time.perf_counter()
    print("cfg5 %d entries: read %.2f s, build+upload on %s %.2f s, anagrams %d" % (n, t1 - t0, "device" if where else "host", t2 - t1, m.index_size()), flush=True)
    del m
PY
ANL_PROFILE=1 timeout 900 python /tmp/buildtime.py 2000000 2>&1 | grep -E "cfg5|gpu build|build:" | tee gpurun_out/r02g_buildtime_2M.txt
ANL_PROFILE=1 timeout 1500 python /tmp/buildtime.py 10000000 2>&1 | grep -E "cfg5|gpu build|build:" | tee gpurun_out/r02g_buildtime_10M.txt
timeout 900 python bench.py --workload cfg5:10000000 --steps 3 --warmup 3 --e2e-steps 3 --cpu-sample 100 > gpurun_out/r02g_cfg5_10M.json 2> gpurun_out/r02g_cfg5_10M.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02g_cfg5_10M.json").read().strip().splitlines()[-1]); k=d["kernels"]
    print("cfg5:10M value %.2fM | stages %s | e2e %.2fM | build %.1fs | frac %.3f" % (d["value"]/1e6, {a: round(b,2) for a,b in k["stages_ms"].items()}, d["e2e"]["value"]/1e6, d["config"]["build_seconds"], d["roofline"]["frac"]))
except Exception as e: print("failed", e); print(open("gpurun_out/r02g_cfg5_10M.err").read()[-1500:])
PY
