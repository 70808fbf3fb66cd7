#!/bin/bash
# random-read microbenchmark with the L2 fetch granularity at its default and at 32 B; e2e with few host threads (the
# per-rank share at 8 GPUs); eng3 bench; cfg5 (10 M entries) bench + ncu evidence of the probe kernels in the HBM regime
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/randread tools/micro/randread.cu
timeout 300 /tmp/randread > gpurun_out/r02i_randread_default.txt; cat gpurun_out/r02i_randread_default.txt
timeout 300 /tmp/randread 32 > gpurun_out/r02i_randread_g32.txt; cat gpurun_out/r02i_randread_g32.txt
nproc
timeout 900 python tools/e2e_sweep.py 131072:4:2 131072:4:4 131072:4:8 131072:4 > gpurun_out/r02i_sweep.txt 2>&1; cat gpurun_out/r02i_sweep.txt
timeout 600 python bench.py --workload eng3 --steps 5 --warmup 3 > gpurun_out/r02i_eng3.json 2> gpurun_out/r02i_eng3.err; tail -c 700 gpurun_out/r02i_eng3.json
timeout 900 python bench.py --workload cfg5:10000000 --steps 3 --warmup 3 --e2e-steps 5 --cpu-sample 100 > gpurun_out/r02i_cfg5_10M.json 2> gpurun_out/r02i_cfg5_10M.err; tail -c 700 gpurun_out/r02i_cfg5_10M.json
Q=1000000 W=cfg5:10000000 KERNELS="bloom exact" timeout 1500 bash tools/profile.sh > gpurun_out/profile_cfg5.log 2>&1
tail -4 gpurun_out/profile_cfg5.log
