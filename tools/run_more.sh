set -x
mkdir -p gpurun_out
python bench.py --workload cfg1 --steps 20 --warmup 5 --e2e-steps 5 --cpu-sample 10000 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
python bench.py --workload cfg4 --steps 3 --warmup 3 --e2e-steps 2 --cpu-sample 1500 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
( time python bench.py --workload cfg5:2000000 --steps 3 --warmup 3 --e2e-steps 2 --cpu-sample 300 ) > gpurun_out/bench_cfg5_2m.json 2> gpurun_out/bench_cfg5_2m.err
tail -4 gpurun_out/*.err
