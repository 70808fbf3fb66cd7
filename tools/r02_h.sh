#!/bin/bash
# device-side batch producer (gpu_segment.cu): whole suite, cfg3 bench with either producer, cfg2 bench, ncu evidence for cfg2
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
for seg in host device; do
  ANL_PROFILE=1 ANL_SEGMENT=$seg timeout 600 python bench.py --workload cfg3 --queries 4000000 --steps 3 --warmup 3 > gpurun_out/r02h_cfg3_$seg.json 2> gpurun_out/r02h_cfg3_$seg.err
  tail -c 900 gpurun_out/r02h_cfg3_$seg.json; grep -E "segmentation" gpurun_out/r02h_cfg3_$seg.err | tail -4
done
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/r02h_cfg2.json 2> gpurun_out/r02h_cfg2.err
tail -c 600 gpurun_out/r02h_cfg2.json
Q=1000000 W=cfg2 timeout 1500 bash tools/profile.sh > gpurun_out/profile.log 2>&1
tail -5 gpurun_out/profile.log
