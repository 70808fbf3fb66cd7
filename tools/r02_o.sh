#!/bin/bash
# find_all_matches after the shared variant lists / recycled segmentation buffers / pooled device scratch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_segment.py tests/test_gpu_zz_consolidation.py tests/test_gpu_learn.py tests/test_gpu_parity.py tests/test_gpu_query_failures.py tests/test_gpu_multi_device.py -m gpu -q 2>&1 | tail -3
ANL_PROFILE=1 timeout 600 python bench.py --workload cfg3 --queries 4000000 --steps 3 --warmup 3 > gpurun_out/r02o_cfg3_4M.json 2> gpurun_out/r02o_cfg3_4M.err
grep -E "search:|device segmentation|consolidate:" gpurun_out/r02o_cfg3_4M.err | tail -26
timeout 900 python bench.py --workload cfg3 --queries 20000000 --steps 2 --warmup 3 > gpurun_out/r02o_cfg3_20M.json 2>/dev/null
python - <<PY
import json
for f in ("4M", "20M"):
    d=json.loads(open("gpurun_out/r02o_cfg3_%s.json" % f).read().strip().splitlines()[-1]); print("cfg3 %s: %.2f M lookups/s %.2f M tokens/s consolidate %.0f ms" % (f, d["value"]/1e6, d["tokens_per_s"]/1e6, d["consolidate_ms"]))
PY
