#!/bin/bash
# whole GPU suite on the current tree, the driver's default bench line, the DP kernel's main launch under ncu
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/r02k_default.json 2> gpurun_out/r02k_default.err; tail -c 1500 gpurun_out/r02k_default.json; tail -3 gpurun_out/r02k_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02k_reference.json 2> gpurun_out/r02k_reference.err; tail -c 400 gpurun_out/r02k_reference.json
rm -f gpurun_out/prof_*.ncu-rep
Q=1000000 W=cfg2 KERNELS="dp" timeout 900 bash tools/profile.sh > gpurun_out/profile_dp.log 2>&1; tail -3 gpurun_out/profile_dp.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
