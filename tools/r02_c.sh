#!/bin/bash
# pair-list score stage: GPU suite with it on (default), the sensitive parity files again with it off, then cfg2 / cfg4
# benches both ways and a launch list
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
ANL_PAIRS=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_nopairs.log; cat gpurun_out/pytest_nopairs.log
for mode in 1 0; do
for w in cfg2 cfg4; do
  ANL_PAIRS=$mode timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --e2e-steps 5 --cpu-sample 200 > gpurun_out/r02c_${w}_p$mode.json 2> gpurun_out/r02c_${w}_p$mode.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_${w}_p$mode.json").read().strip().splitlines()[-1]); k=d["kernels"]; c=d["counters"]
    print("$w pairs=$mode value %.2fM | stages %s | e2e %.2fM | dp_pairs %d dp_cells %.3g dl_cells %.3g" % (d["value"]/1e6, {a: round(b,2) for a,b in k["stages_ms"].items()}, d["e2e"]["value"]/1e6, c["dp_pairs"], c["dp_cells"], c["dl_cells"]))
except Exception as e: print("$w $mode failed", e); print(open("gpurun_out/r02c_${w}_p$mode.err").read()[-1500:])
PY
done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02c_launches.csv \
    python bench.py --workload cfg2 --queries 262144 --steps 1 --warmup 3 --e2e-steps 0 --cpu-sample 64 > /dev/null 2> gpurun_out/launches.err
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02c_launches.csv")))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); k = d["Kernel Name"].split("(")[0]; v = float(d["Metric Value"].replace(",", ""))
        u = d.get("Metric Unit", "")
        v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
        agg.setdefault(k, []).append(v)
for k, v in agg.items(): print("%-28s n=%3d last=%.3f ms" % (k[:28], len(v), v[-1]))
PY
