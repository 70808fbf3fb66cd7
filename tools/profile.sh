#!/bin/bash
# ncu evidence for the round (run under gpurun on one B200):
#   1. launch list with per-launch device time (cold-cache, serialised -> compare SHARES)
#   2. one full capture of each of the three main kernels (Bloom stage, exact stage, score/rank), at the
#      bench's own launch size
# Outputs land in gpurun_out/ ; summaries are copied into profiles/ by tools/summarize_ncu.py.
set -x
mkdir -p gpurun_out
Q=${Q:-1000000}
W=${W:-cfg2}
echo "$W $Q" > gpurun_out/profile_launch.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(encode|probe|bloom|exact|pairfilter|pairscan|pairscatter|dp|dp_tma|rank|triage|confusable|confusable_wide|finish|count|export|offsets|patch|merge)_kernel' -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --workload $W --queries $Q --steps 2 --warmup 3 --e2e-steps 0 --cpu-sample 64 > gpurun_out/launches_bench.json 2> gpurun_out/launches.err
KERNELS=${KERNELS:-"bloom exact pairfilter dp rank confusable"}
for K in $KERNELS; do
  S=3  # three warm-up passes skipped: the capture is the launch of the timed pass
  if [ "$K" = dp ]; then S=6; fi  # (two launches per pass: the short class -- the main one -- then the long class)
  ncu --set full --clock-control none --import-source on -k regex:^${K}_kernel -s $S -c 1 -f -o gpurun_out/prof_$K \
      python bench.py --workload $W --queries $Q --steps 1 --warmup 3 --e2e-steps 0 --cpu-sample 64 > /dev/null 2> gpurun_out/prof_$K.err
done
ls -la gpurun_out
tail -n 2 gpurun_out/launches.err gpurun_out/prof_*.err
