#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_widening_gpu.py tests/test_x_oab.py tests/test_y_reference_suites_gpu.py -m gpu -q --durations=8 ) > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -25 gpurun_out/pytest_gpu2.log
( time timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
