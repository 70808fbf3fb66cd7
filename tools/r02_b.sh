#!/bin/bash
# round 2, after the export-stage / multi-device rework: whole GPU suite, then cfg2 bench (1 GPU, and 2 GPUs in one process)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 10 --cpu-sample 2000 > gpurun_out/r02b_cfg2.json 2> gpurun_out/r02b_cfg2.err; tail -3 gpurun_out/r02b_cfg2.err
python - <<PY
import json
for f in ("r02b_cfg2",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); k=d["kernels"]
        print(f, "value %.2fM | stages %s | e2e %.2fM" % (d["value"]/1e6, {a: round(b,2) for a,b in k["stages_ms"].items()}, d["e2e"]["value"]/1e6))
    except Exception as e: print(f, "failed", e)
PY
ANL_HOST_THREADS=4 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 10 --cpu-sample 200 > gpurun_out/r02b_cfg2_t4.json 2> gpurun_out/r02b_cfg2_t4.err
timeout 600 python bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 10 --cpu-sample 200 > gpurun_out/r02b_cfg2_sp2.json 2> gpurun_out/r02b_cfg2_sp2.err; tail -3 gpurun_out/r02b_cfg2_sp2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 10 > gpurun_out/r02b_cfg2_n2.json 2> gpurun_out/r02b_cfg2_n2.err
python - <<PY
import json
for f in ("r02b_cfg2_t4","r02b_cfg2_sp2","r02b_cfg2_n2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.2fM | e2e %.2fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6), d["e2e"])
    except Exception as e: print(f, "failed", e)
PY
