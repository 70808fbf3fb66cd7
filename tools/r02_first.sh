#!/bin/bash
# round 2, first GPU call (2 GPUs): the whole GPU suite incl. the new cfg-5 / per-query-failure tests and the real-NCCL
# sharded test; the lexicon-sharded bench on the cfg-5 lexicon; an ncu capture of the probe kernels in the HBM regime
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nproc >> gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --sharded --workload cfg5:2000000 --steps 3 --warmup 3 > gpurun_out/sharded_cfg5_2M_n2.json 2> gpurun_out/sharded_cfg5_2M_n2.err
tail -c 1500 gpurun_out/sharded_cfg5_2M_n2.json; tail -5 gpurun_out/sharded_cfg5_2M_n2.err
timeout 600 python bench.py --workload cfg5:2000000 --steps 3 --warmup 3 --e2e-steps 3 --cpu-sample 200 > gpurun_out/bench_cfg5_2M.json 2> gpurun_out/bench_cfg5_2M.err
tail -c 2500 gpurun_out/bench_cfg5_2M.json
for K in bloom exact; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 3 -c 1 -f -o gpurun_out/r02a_cfg5_$K \
      python bench.py --workload cfg5:2000000 --queries 262144 --steps 1 --warmup 3 --e2e-steps 0 --cpu-sample 64 > /dev/null 2> gpurun_out/prof_$K.err
done
ls -la gpurun_out
