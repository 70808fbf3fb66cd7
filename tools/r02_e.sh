#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/e2e_sweep.py 65536:4 131072:4 262144:4 65536:8 32768:8 > gpurun_out/r02e_sweep.txt 2>&1; cat gpurun_out/r02e_sweep.txt
timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --e2e-steps 5 --cpu-sample 200 > gpurun_out/r02e_cfg2.json 2> gpurun_out/r02e_cfg2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02e_cfg2.json").read().strip().splitlines()[-1]); k=d["kernels"]
print("cfg2 value %.2fM | stages %s | e2e %.2fM" % (d["value"]/1e6, {a: round(b,2) for a,b in k["stages_ms"].items()}, d["e2e"]["value"]/1e6))
PY
