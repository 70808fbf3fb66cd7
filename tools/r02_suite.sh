#!/bin/bash
# the whole GPU suite (no -x: every failure is reported)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
