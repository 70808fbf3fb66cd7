#!/bin/bash
# 8 GPUs of one box: query-partitioned replicas (torchrun, one process per GPU; and one process driving all GPUs through
# anl_model_build_multi), the lexicon-sharded mode with its NCCL exchange, and the multi-GPU tests the 1-GPU box skips
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02j_cfg2_n$N.json 2> gpurun_out/r02j_cfg2_n$N.err
tail -c 500 gpurun_out/r02j_cfg2_n$N.json; tail -2 gpurun_out/r02j_cfg2_n$N.err
timeout 600 python bench.py --gpus $N --steps 5 --warmup 3 --cpu-sample 2000 > gpurun_out/r02j_cfg2_single_process_n$N.json 2> gpurun_out/r02j_cfg2_single_process_n$N.err
tail -c 500 gpurun_out/r02j_cfg2_single_process_n$N.json; tail -2 gpurun_out/r02j_cfg2_single_process_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --sharded --workload cfg5:2000000 --steps 5 --warmup 3 > gpurun_out/r02j_sharded_cfg5_n$N.json 2> gpurun_out/r02j_sharded_cfg5_n$N.err
tail -c 900 gpurun_out/r02j_sharded_cfg5_n$N.json; tail -2 gpurun_out/r02j_sharded_cfg5_n$N.err
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_multi_device.py -m gpu -q 2>&1 | tail -5
