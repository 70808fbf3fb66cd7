#!/bin/bash
# 2 GPUs: whole suite (device-only index forms; NCCL exchange inside the library), sharded bench, cfg2 / cfg5 benches
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
for w in cfg5:2000000 cfg2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --sharded --workload $w --steps 5 --warmup 3 > gpurun_out/r02f_sharded_${w%%:*}_n2.json 2> gpurun_out/r02f_sharded_${w%%:*}_n2.err
tail -c 1200 gpurun_out/r02f_sharded_${w%%:*}_n2.json; tail -3 gpurun_out/r02f_sharded_${w%%:*}_n2.err
done
for w in cfg2 cfg5:2000000; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --e2e-steps 5 --cpu-sample 200 > gpurun_out/r02f_${w%%:*}.json 2> gpurun_out/r02f_${w%%:*}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02f_${w%%:*}.json").read().strip().splitlines()[-1]); k=d["kernels"]
    print("$w value %.2fM | stages %s | e2e %.2fM" % (d["value"]/1e6, {a: round(b,2) for a,b in k["stages_ms"].items()}, d["e2e"]["value"]/1e6))
except Exception as e: print("$w failed", e); print(open("gpurun_out/r02f_${w%%:*}.err").read()[-1500:])
PY
done
