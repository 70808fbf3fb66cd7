#!/bin/bash
# TMA-staged DP, two-queue confusable stage, block-grab rank kernel: suite, A/B benches, launch list, random-read microbench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
for tma in 1 0; do
for w in cfg2 cfg4; do
  ANL_DP_TMA=$tma timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --e2e-steps 5 --cpu-sample 200 > gpurun_out/r02d_${w}_t$tma.json 2> gpurun_out/r02d_${w}_t$tma.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02d_${w}_t$tma.json").read().strip().splitlines()[-1]); k=d["kernels"]; c=d["counters"]
    print("$w tma=$tma value %.2fM | stages %s | e2e %.2fM" % (d["value"]/1e6, {a: round(b,2) for a,b in k["stages_ms"].items()}, d["e2e"]["value"]/1e6))
except Exception as e: print("$w $tma failed", e); print(open("gpurun_out/r02d_${w}_t$tma.err").read()[-1500:])
PY
done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02d_launches.csv \
    python bench.py --workload cfg2 --queries 262144 --steps 1 --warmup 3 --e2e-steps 0 --cpu-sample 64 > /dev/null 2> gpurun_out/launches.err
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02d_launches.csv")))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); k = d["Kernel Name"].split("(")[0]; v = float(d["Metric Value"].replace(",", ""))
        u = d.get("Metric Unit", "")
        v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
        agg.setdefault(k, []).append(round(v, 3))
for k, v in agg.items(): print("%-28s %s" % (k[:28], v[-4:]))
PY
timeout 300 tools/micro/randread > gpurun_out/r02d_randread.txt 2>&1; cat gpurun_out/r02d_randread.txt
