#!/bin/bash
# final state of the round: whole GPU suite, smoke, the driver's default bench line and the reference arm, the other configs
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/r02p_default.json 2> gpurun_out/r02p_default.err; tail -2 gpurun_out/r02p_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02p_reference.json 2> /dev/null
for w in eng3 cfg4 cfg1 cfg5:10000000; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --cpu-sample 400 > gpurun_out/r02p_${w%%:*}.json 2> gpurun_out/r02p_${w%%:*}.err
done
python - <<PY
import json
for w in ("default", "eng3", "cfg4", "cfg1", "cfg5"):
    try:
        d=json.loads(open("gpurun_out/r02p_%s.json" % w).read().strip().splitlines()[-1]); k=d["kernels"]["stages_ms"]
        print("%s value %.2fM e2e %.2fM cpu %.1f | %s | frac %.3f dom %s %.3f" % (w, d["value"]/1e6, d["e2e"]["value"]/1e6, d["cpu_baseline"]["value"], {a: round(b,2) for a,b in k.items()}, d["roofline"]["frac"], d["roofline"]["time_dominant_kernel"], d["roofline"]["time_dominant_frac"] or 0))
    except Exception as e: print(w, "failed", e)
d=json.loads(open("gpurun_out/r02p_reference.json").read().strip().splitlines()[-1]); print("reference arm %.1f q/s on %d cores" % (d["value"], d["cpu_baseline"]["cores"]))
PY
