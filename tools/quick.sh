#!/bin/bash
# developer loop on the GPU box: parity tests, then a short bench of the given workloads (default cfg2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
for w in ${@:-cfg2}; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --e2e-steps 3 --cpu-sample 2000 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json")); k=d["kernels"]
    print("$w value %.2fM q/s | probe %.2f score %.2f rescore %.2f ms | e2e %.2fM | cpu %.0f | frac %.3f | gcups %.0f" % (d["value"]/1e6, k["probe_ms"], k["score_ms"], k.get("rescore_ms",0), d["e2e"]["value"]/1e6, d["cpu_baseline"]["value"], d["roofline"]["frac"], d["dp_gcups"]))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-2000:])
PY
done
