#!/bin/bash
# A/B of developer knobs without the parity tests: usage quick2.sh "<ENV=..>" ... ; each arg = one env setting for a cfg2 bench
mkdir -p gpurun_out
W=${W:-cfg2}
for envs in "$@"; do
  env $envs timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --e2e-steps ${E2E:-1} --cpu-sample 64 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json")); k=d["kernels"]
    print("$envs $W value %.2fM q/s | probe %.2f score %.2f rescore %.2f ms | e2e %.2fM" % (d["value"]/1e6, k["probe_ms"], k["score_ms"], k.get("rescore_ms",0), d["e2e"]["value"]/1e6))
except Exception as e:
    print("$envs failed", e); print(open("gpurun_out/ab.err").read()[-1500:])
PY
done
