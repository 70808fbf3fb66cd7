#!/bin/bash
# 2 GPUs: the lexicon-sharded mode after the device-side overflow check and the one-load header gather in merge_kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_multi_device.py -m gpu -q 2>&1 | tail -5
for w in cfg5:2000000 cfg2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --sharded --workload $w --steps 5 --warmup 3 > gpurun_out/r02l_sharded_${w%%:*}_n2.json 2> gpurun_out/r02l_sharded_${w%%:*}_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02l_sharded_${w%%:*}_n2.json").read().strip().splitlines()[-1]); print("$w", round(d["value"]/1e6,2), "M q/s", {k: (round(v,2) if isinstance(v,float) else v) for k,v in d["kernels"].items()})
PY
done
