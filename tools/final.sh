#!/bin/bash
# round-end measurement pass on one B200 (under gpurun): parity tests, ncu evidence, the bench lines of every config
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
tools/profile.sh > gpurun_out/profile.log 2>&1
python bench.py > gpurun_out/final_cfg2.json 2> gpurun_out/final_cfg2.err
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py --workload cfg1 --steps 20 --warmup 5 --e2e-steps 5 --cpu-sample 10000 > gpurun_out/final_cfg1.json 2> gpurun_out/final_cfg1.err
python bench.py --workload cfg4 --steps 3 --warmup 3 --e2e-steps 3 --cpu-sample 1500 > gpurun_out/final_cfg4.json 2> gpurun_out/final_cfg4.err
python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/final_cfg3.json 2> gpurun_out/final_cfg3.err
[ -n "$SKIP_CFG5" ] || python bench.py --workload cfg5:10000000 --steps 3 --warmup 3 --e2e-steps 2 --cpu-sample 200 > gpurun_out/final_cfg5_10M.json 2> gpurun_out/final_cfg5_10M.err
for f in cfg2 ref cfg1 cfg4 cfg3 cfg5_10M; do
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/final_$f.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("$f", "value %.4g %s" % (d["value"], d["unit"]), "| e2e %.4g" % e.get("value", float("nan")), "| ms/step %.3f" % d.get("ms_per_step", float("nan")),
          "| stages", {k: round(v, 2) for k, v in (d.get("kernels", {}).get("stages_ms") or {}).items()})
except Exception as ex:
    print("$f failed", ex); print(open("gpurun_out/final_$f.err").read()[-1500:])
PY
done
